#!/bin/bash
# compare build variants of the tiled pair kernel: bash tools/gpu/r02_variants.sh "<defines A>" "<defines B>" ...
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
i=0
for DEF in "$@"; do
  i=$((i+1))
  echo "=== variant $i: $DEF"
  CSF_BUILD_DEFINES="$DEF" python -m cyclistsocialforce_b200.build --force > /dev/null 2>&1 || { echo "build failed"; continue; }
  timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pair_forces and tiled and 5000" 2>&1 | tail -1
  for N in 65536 8192; do
    CSF_BENCH_N=$N timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/var${i}_$N.json 2> gpurun_out/var${i}_$N.err
    python - <<PY
import json
d=json.load(open("gpurun_out/var${i}_$N.json")); r=d["roofline"]
print("N=$N ms/step %.4f K1 ms %.4f frac %.3f" % (d["ms_per_step"], r["kernel_ms"], r["frac"]))
PY
  done
done
python -m cyclistsocialforce_b200.build --force > /dev/null 2>&1
