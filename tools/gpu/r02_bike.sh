#!/bin/bash
# tiled v0.1 Bicycle field: the GPU tests that touch it, then a 16,384-agent Bicycle crowd tiled vs dense
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -q -m gpu -k "bicycle or v01 or mixed" > gpurun_out/bike_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/bike_pytest.log | cut -c1-400
for MODE in tiled dense; do
timeout 300 python tools/bench_models.py --models bicycle --steps 20 --pair-mode $MODE --count-pairs > gpurun_out/bike_bench_$MODE.json 2> gpurun_out/bike_bench_$MODE.err
echo "bench $MODE rc=$?"; tail -2 gpurun_out/bike_bench_$MODE.err; cut -c1-600 gpurun_out/bike_bench_$MODE.json
done
