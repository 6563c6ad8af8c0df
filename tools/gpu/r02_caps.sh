#!/bin/bash
# K1 survivor-buffer capacity variants (prebuilt under variants/) x items per block: full crowd and shards
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
show() { python - <<PY
import json
d=json.load(open("$1")); r=d["roofline"]
print("$2: ms/step %.4f | K1 %.4f ms | frac %.3f" % (d["ms_per_step"], r["kernel_ms"], r["frac"]))
PY
}
for V in base "$@"; do
  if [ "$V" = "base" ]; then unset CSF_B200_LIB; else export CSF_B200_LIB=$PWD/variants/lib$V.so; fi
  for G in 0 1 3; do
  export CSF_TILED_GROUPS=$G
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/cap_${V}_g${G}.json 2> gpurun_out/cap_${V}_g${G}.err; show gpurun_out/cap_${V}_g${G}.json "$V groups=$G full "
  done
  unset CSF_TILED_GROUPS
  CSF_BENCH_EMULATE_WORLD=2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/cap_${V}_half.json 2> gpurun_out/cap_${V}_half.err; show gpurun_out/cap_${V}_half.json "$V half "
  CSF_BENCH_EMULATE_WORLD=8 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/cap_${V}_shard.json 2> gpurun_out/cap_${V}_shard.err; show gpurun_out/cap_${V}_shard.json "$V shard"
done
