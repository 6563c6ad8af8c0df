#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
show() { python - <<PY
import json
d=json.load(open("$1")); r=d["roofline"]; a=d["roofline_agent_kernel"]
print("$2: ms/step %.4f value %.4g e2e %.4g | prep %.4f K1 %.4f agent %.4f ms | frac %.3f pairs %.3g" % (d["ms_per_step"], d["value"], d["e2e"]["value"], r.get("prepare_kernel_ms", 0), r["kernel_ms"], a["kernel_ms"], r["frac"], r["executed_pairs_per_launch"]))
PY
}
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q -m gpu -k "pair_forces or tiled or graph_step" 2>&1 | tail -2
for G in 0 1 3; do
  CSF_TILED_GROUPS=$G timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/full_g$G.json 2> gpurun_out/full_g$G.err
  show gpurun_out/full_g$G.json "full groups=$G"
done
for G in 0 4 8; do
  CSF_TILED_GROUPS=$G CSF_BENCH_EMULATE_WORLD=8 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/shard_g$G.json 2> gpurun_out/shard_g$G.err
  show gpurun_out/shard_g$G.json "shard groups=$G"
done
for N in 4096 16384; do
  CSF_BENCH_N=$N timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/n$N.json 2> gpurun_out/n$N.err
  show gpurun_out/n$N.json "N=$N"
done
