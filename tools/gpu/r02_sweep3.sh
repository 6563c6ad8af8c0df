#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
show() { python - <<PY
import json
d=json.load(open("$1")); r=d["roofline"]; a=d["roofline_agent_kernel"]
print("$2: ms/step %.4f | prep %.4f K1 %.4f agent %.4f ms | frac %.3f" % (d["ms_per_step"], r.get("prepare_kernel_ms", 0), r["kernel_ms"], a["kernel_ms"], r["frac"]))
PY
}
for W in 8 4 2; do
for G in 2 3 6; do
  CSF_TILED_GROUPS=$G CSF_BENCH_EMULATE_WORLD=$W timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s3.json 2> gpurun_out/s3.err
  show gpurun_out/s3.json "1/$W shard groups=$G"
done
done
