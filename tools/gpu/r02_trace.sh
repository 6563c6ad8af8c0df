#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CSF_BUILD_DEFINES=-DCSF_TILED_PROF python -m cyclistsocialforce_b200.build --force > /dev/null 2>&1
for G in 0 1 4; do
echo "== shard groups=$G"
CSF_TILED_GROUPS=$G CSF_BENCH_EMULATE_WORLD=8 CSF_BENCH_ROLES=1 timeout 300 python bench.py --steps 5 --warmup 5 --no-cpu-baseline 2>&1 >/dev/null | grep -iE "roles|item trace"
cp gpurun_out/k1_item_trace.npy gpurun_out/k1_item_trace_shard_g$G.npy
done
echo "== full"
CSF_BENCH_ROLES=1 timeout 300 python bench.py --steps 5 --warmup 5 --no-cpu-baseline 2>&1 >/dev/null | grep -iE "roles|item trace"
cp gpurun_out/k1_item_trace.npy gpurun_out/k1_item_trace_full.npy
python -m cyclistsocialforce_b200.build --force > /dev/null 2>&1
