#!/bin/bash
# parity of the pair kernels + bench (+ optional ncu of K1): bash tools/gpu/r02_k1_check.sh <tag> [ncu]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-k1}
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pair_forces or tiled" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python -m pytest tests/test_gpu_api.py -x -q -m gpu -k "graph_step or f32_per_step" >> gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest2 rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json | head -c 3000; tail -3 gpurun_out/${TAG}_bench.err
if [ "$2" = "ncu" ]; then
  export CSF_BENCH_GRAPH=0
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:pair_tiled -s 8 -c 1 -f -o gpurun_out/${TAG}_k1 \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
  echo "ncu rc=$?"
fi
