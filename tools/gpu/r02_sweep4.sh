#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
show() { python - <<PY
import json
d=json.load(open("$1")); r=d["roofline"]; a=d["roofline_agent_kernel"]
print("$2: ms/step %.4f | prep %.4f K1 %.4f agent %.4f ms | frac %.3f" % (d["ms_per_step"], r.get("prepare_kernel_ms", 0), r["kernel_ms"], a["kernel_ms"], r["frac"]))
PY
}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q -m gpu -k "pair_forces or tiled or graph_step or f32_per_step or stop_dest" 2>&1 | tail -3
for WIDE in 1 0; do
for W in 8 4 2 0; do
  CSF_TILED_WIDE=$WIDE CSF_BENCH_EMULATE_WORLD=$W timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s4.json 2> gpurun_out/s4.err
  show gpurun_out/s4.json "wide=$WIDE 1/$W shard"
done
done
for N in 4096 16384; do
  CSF_BENCH_N=$N timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s4.json 2> gpurun_out/s4.err
  show gpurun_out/s4.json "N=$N (auto)"
done
