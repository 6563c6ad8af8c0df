#!/bin/bash
# full GPU test suite + bench + emulated shard: bash tools/gpu/r02_full.sh <tag>
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-full}
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/${TAG}_pytest.log | cut -c1-600
show() { python - <<PY
import json
d=json.load(open("$1")); r=d["roofline"]; a=d["roofline_agent_kernel"]
print("$1: ms/step %.4f value %.4g e2e %.4g | prep %.4f K1 %.4f agent %.4f ms | frac %.3f pairs %.3g" % (d["ms_per_step"], d["value"], d["e2e"]["value"], r.get("prepare_kernel_ms", 0), r["kernel_ms"], a["kernel_ms"], r["frac"], r["executed_pairs_per_launch"]))
PY
}
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err; show gpurun_out/${TAG}_bench.json
for G in 1 2 8; do
CSF_TILED_GROUPS=$G CSF_BENCH_EMULATE_WORLD=8 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_shard_g$G.json 2> gpurun_out/${TAG}_shard_g$G.err
show gpurun_out/${TAG}_shard_g$G.json
done
CSF_BENCH_N=1048576 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_1M.json 2> gpurun_out/${TAG}_bench_1M.err
show gpurun_out/${TAG}_bench_1M.json
