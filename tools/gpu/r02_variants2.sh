#!/bin/bash
# build variants x {full crowd, emulated 1/8 shard}: bash tools/gpu/r02_variants2.sh "<defines A>" ...
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
show() { python - <<PY
import json
d=json.load(open("$1")); r=d["roofline"]; a=d["roofline_agent_kernel"]
print("$2: ms/step %.4f | prep %.4f K1 %.4f agent %.4f ms | frac %.3f" % (d["ms_per_step"], r.get("prepare_kernel_ms", 0), r["kernel_ms"], a["kernel_ms"], r["frac"]))
PY
}
i=0
for DEF in "$@"; do
  i=$((i+1))
  echo "=== variant $i: $DEF"
  CSF_BUILD_DEFINES="$DEF" python -m cyclistsocialforce_b200.build --force > /dev/null 2>&1 || { echo "build failed"; continue; }
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/v${i}_full.json 2> gpurun_out/v${i}_full.err; show gpurun_out/v${i}_full.json "full "
  CSF_BENCH_EMULATE_WORLD=8 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/v${i}_shard.json 2> gpurun_out/v${i}_shard.err; show gpurun_out/v${i}_shard.json "shard"
  CSF_TILED_GROUPS=4 CSF_BENCH_EMULATE_WORLD=8 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/v${i}_shard4.json 2> gpurun_out/v${i}_shard4.err; show gpurun_out/v${i}_shard4.json "shard g=4"
done
python -m cyclistsocialforce_b200.build --force > /dev/null 2>&1
