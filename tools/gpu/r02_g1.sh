#!/bin/bash
# GPU run: full -m gpu suite, bench (1 GPU), launch list + one ncu --set full capture of the pair kernel and of the
# agent kernel:  bash tools/gpu/r02_g1.sh <tag>
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-g1}
rm -f gpurun_out/parity_report.jsonl
timeout 1700 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log | cut -c1-400
show() { python - <<PY
import json
d=json.load(open("$1")); r=d["roofline"]; a=d["roofline_agent_kernel"]
print("$1: ms/step %.4f value %.4g e2e %.4g | prep %.4f K1 %.4f agent %.4f ms | frac %.3f pairs %.3g" % (d["ms_per_step"], d["value"], d["e2e"]["value"], r.get("prepare_kernel_ms", 0), r["kernel_ms"], a["kernel_ms"], r["frac"], r["executed_pairs_per_launch"]))
PY
}
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err; show gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err
echo "ref rc=$?"; cut -c1-600 gpurun_out/${TAG}_ref.json
export CSF_BENCH_GRAPH=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_tiled_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_pair \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_pair.log 2>&1
echo "ncu pair rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:agent_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_agent \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_agent.log 2>&1
echo "ncu agent rc=$?"
