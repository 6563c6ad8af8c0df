#!/bin/bash
# final single-GPU pass: suite, default bench (+ reference arm), 1M / 4,096 benches, item-group check for a half crowd,
# launch list + ncu --set full captures of the pair kernel and the agent kernel with the final build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-fin}
rm -f gpurun_out/parity_report.jsonl
timeout 1700 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log | cut -c1-300
show() { python - <<PY
import json
d=json.load(open("$1")); r=d["roofline"]; a=d["roofline_agent_kernel"]
print("$1: ms/step %.4f value %.4g e2e %.4g | prep %.4f K1 %.4f agent %.4f ms | frac %.3f pairs %.3g" % (d["ms_per_step"], d["value"], d["e2e"]["value"], r.get("prepare_kernel_ms", 0), r["kernel_ms"], a["kernel_ms"], r["frac"], r["executed_pairs_per_launch"]))
PY
}
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err; show gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/${TAG}_ref.json
CSF_BENCH_N=1048576 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_1M.json 2> gpurun_out/${TAG}_bench_1M.err
echo "bench 1M rc=$?"; show gpurun_out/${TAG}_bench_1M.json
CSF_BENCH_N=4096 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_4096.json 2> gpurun_out/${TAG}_bench_4096.err
echo "bench 4096 rc=$?"; show gpurun_out/${TAG}_bench_4096.json
for G in 2 3; do
CSF_TILED_GROUPS=$G CSF_BENCH_EMULATE_WORLD=2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_half_g$G.json 2> gpurun_out/${TAG}_half_g$G.err; show gpurun_out/${TAG}_half_g$G.json
done
export CSF_BENCH_GRAPH=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_tiled_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_pair \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu_pair.log 2>&1
echo "ncu pair rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:agent_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_agent \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu_agent.log 2>&1
echo "ncu agent rc=$?"
unset CSF_BENCH_GRAPH
timeout 600 ncu --set full --clock-control none --import-source on -k regex:agent_kernel -s 6 -c 1 -f -o gpurun_out/${TAG}_agent_ip \
    python tools/bench_scenarios.py --steps 4 --warmup 4 --no-graph > gpurun_out/${TAG}_ncu_ip.log 2>&1
echo "ncu ip rc=$?"
