#!/bin/bash
# one rank's shard of an 8-way split of the 65,536 crowd, emulated on one GPU: sweep the item split
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for G in 1 2 4 8; do
  CSF_TILED_GROUPS=$G CSF_BENCH_EMULATE_WORLD=8 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/shard_g$G.json 2> gpurun_out/shard_g$G.err
  python - <<PY
import json
d=json.load(open("gpurun_out/shard_g$G.json")); r=d["roofline"]; a=d["roofline_agent_kernel"]
print("groups=$G ms/step %.4f K1 ms %.4f frac %.3f agent ms %.4f pairs %.3g" % (d["ms_per_step"], r["kernel_ms"], r["frac"], a["kernel_ms"], r["executed_pairs_per_launch"]))
PY
done
CSF_BUILD_DEFINES=-DCSF_TILED_PROF python -m cyclistsocialforce_b200.build --force > /dev/null 2>&1
for G in 1 8; do
CSF_TILED_GROUPS=$G CSF_BENCH_EMULATE_WORLD=8 CSF_BENCH_ROLES=1 timeout 300 python bench.py --steps 5 --warmup 5 --no-cpu-baseline 2>&1 >/dev/null | grep -i "roles" 
done
python -m cyclistsocialforce_b200.build --force > /dev/null 2>&1
