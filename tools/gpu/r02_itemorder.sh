#!/bin/bash
# pair kernel on a shard before / after the spatial re-sort at pair call 64 (emulated 1/W shard on one GPU):
# visiting order of the shard's targets (partition / curve), item-order policy (refresh / stale)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for W in ${@:-2}; do
for CFG in "partition refresh" "curve refresh"; do
set -- $CFG
CSF_SHARD_TARGET_ORDER=$1 CSF_ITEM_ORDER=$2 CSF_BENCH_DEBUG=1 CSF_BENCH_EMULATE_WORLD=$W timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/io_$1_$2_w$W.json 2> gpurun_out/io_$1_$2_w$W.err
echo "== world $W targets $1 items $2 rc=$?"; grep "second pass" gpurun_out/io_$1_$2_w$W.err | cut -c1-330
done
done
