#!/bin/bash
# survivor buffers of 32 dynamic tiles: parity (pair tests) with the default build, then A/B against 16 and slow-start variants
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -q -m gpu -x > gpurun_out/dt_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/dt_pytest.log | cut -c1-300
show() { python - <<PY
import json
d=json.load(open("$1")); r=d["roofline"]
print("$2: ms/step %.4f | K1 %.4f ms | frac %.3f | pairs %.4g" % (d["ms_per_step"], r["kernel_ms"], r["frac"], r["executed_pairs_per_launch"]))
PY
}
for V in base "$@"; do
  if [ "$V" = "base" ]; then unset CSF_B200_LIB; else export CSF_B200_LIB=$PWD/variants/lib$V.so; fi
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/dt_${V}_full.json 2> gpurun_out/dt_${V}_full.err; show gpurun_out/dt_${V}_full.json "$V full "
  CSF_BENCH_EMULATE_WORLD=2 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/dt_${V}_half.json 2> gpurun_out/dt_${V}_half.err; show gpurun_out/dt_${V}_half.json "$V half "
  CSF_BENCH_EMULATE_WORLD=8 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/dt_${V}_shard.json 2> gpurun_out/dt_${V}_shard.err; show gpurun_out/dt_${V}_shard.json "$V shard"
  CSF_BENCH_N=1048576 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/dt_${V}_1M.json 2> gpurun_out/dt_${V}_1M.err; show gpurun_out/dt_${V}_1M.json "$V 1M   "
done
