#!/bin/bash
# K1 CTA-shape variants prebuilt under variants/ (CSF_B200_LIB selects the library): full crowd + emulated 1/8 shard
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
show() { python - <<PY
import json
d=json.load(open("$1")); r=d["roofline"]; a=d["roofline_agent_kernel"]
print("$2: ms/step %.4f | prep %.4f K1 %.4f agent %.4f ms | frac %.3f" % (d["ms_per_step"], r.get("prepare_kernel_ms", 0), r["kernel_ms"], a["kernel_ms"], r["frac"]))
PY
}
for V in base "$@"; do
  if [ "$V" = "base" ]; then unset CSF_B200_LIB; else export CSF_B200_LIB=$PWD/variants/lib$V.so; fi
  echo "=== $V"
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/shape_${V}_full.json 2> gpurun_out/shape_${V}_full.err; show gpurun_out/shape_${V}_full.json "full "
  CSF_BENCH_EMULATE_WORLD=8 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/shape_${V}_shard.json 2> gpurun_out/shape_${V}_shard.err; show gpurun_out/shape_${V}_shard.json "shard"
  CSF_BENCH_EMULATE_WORLD=2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/shape_${V}_half.json 2> gpurun_out/shape_${V}_half.err; show gpurun_out/shape_${V}_half.json "half "
done
