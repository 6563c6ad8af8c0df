#!/bin/bash
# full GPU test suite + bench (+ ncu of the agent kernel): bash tools/gpu/r02_full.sh <tag> [ncu]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-full}
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest.log | cut -c1-300
show() { python - <<PY
import json
d=json.load(open("$1")); r=d["roofline"]; a=d["roofline_agent_kernel"]
print("$1: ms/step %.4f value %.4g e2e %.4g | prep %.4f K1 %.4f agent %.4f ms | frac %.3f pairs %.3g" % (d["ms_per_step"], d["value"], d["e2e"]["value"], r.get("prepare_kernel_ms", 0), r["kernel_ms"], a["kernel_ms"], r["frac"], r["executed_pairs_per_launch"]))
PY
}
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err; show gpurun_out/${TAG}_bench.json
CSF_BENCH_EMULATE_WORLD=8 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_shard8.json 2> gpurun_out/${TAG}_shard8.err
show gpurun_out/${TAG}_shard8.json
if [ "$2" = "ncu" ]; then
  export CSF_BENCH_GRAPH=0
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:agent_kernel -s 8 -c 1 -f -o gpurun_out/${TAG}_agent \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
  echo "ncu rc=$?"
fi
