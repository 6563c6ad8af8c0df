#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export CSF_BENCH_EMULATE_WORLD=8 CSF_BENCH_GRAPH=0
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pair_tiled -s 8 -c 1 -f -o gpurun_out/shard_k1 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/shard_ncu.log 2>&1
echo "ncu rc=$?"
CSF_BUILD_DEFINES=-DCSF_TILED_PROF python -m cyclistsocialforce_b200.build --force > /dev/null 2>&1
for G in 0 1; do
CSF_TILED_GROUPS=$G CSF_BENCH_ROLES=1 timeout 300 python bench.py --steps 5 --warmup 5 --no-cpu-baseline 2>&1 >/dev/null | grep -i "roles"
done
python -m cyclistsocialforce_b200.build --force > /dev/null 2>&1
