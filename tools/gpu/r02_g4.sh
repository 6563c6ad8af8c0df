#!/bin/bash
# GPU run 4: full -m gpu suite with the 1+2+11 x2 pair-kernel shape and the new tests, default bench (with the
# config-4 extra), N = 1,048,576 on one GPU, ncu of the InvPendulum agent kernel (config 4) + launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-g4}
rm -f gpurun_out/parity_report.jsonl
timeout 1700 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/${TAG}_pytest.log | cut -c1-300
show() { python - <<PY
import json
d=json.load(open("$1")); r=d["roofline"]; a=d["roofline_agent_kernel"]
print("$1: ms/step %.4f value %.4g e2e %.4g | prep %.4f K1 %.4f agent %.4f ms | frac %.3f pairs %.3g" % (d["ms_per_step"], d["value"], d["e2e"]["value"], r.get("prepare_kernel_ms", 0), r["kernel_ms"], a["kernel_ms"], r["frac"], r["executed_pairs_per_launch"]))
print("   extra:", json.dumps(d.get("extra"))[:600])
PY
}
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err; show gpurun_out/${TAG}_bench.json
CSF_BENCH_N=1048576 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_1M.json 2> gpurun_out/${TAG}_bench_1M.err
echo "bench 1M rc=$?"; tail -3 gpurun_out/${TAG}_bench_1M.err; show gpurun_out/${TAG}_bench_1M.json
CSF_BENCH_N=4096 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_4096.json 2> gpurun_out/${TAG}_bench_4096.err
echo "bench 4096 rc=$?"; show gpurun_out/${TAG}_bench_4096.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_c4_launches.csv \
    python tools/bench_scenarios.py --steps 4 --warmup 4 --no-graph > gpurun_out/${TAG}_c4_list.log 2>&1
echo "ncu c4 list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:agent_kernel -s 6 -c 1 -f -o gpurun_out/${TAG}_agent_ip \
    python tools/bench_scenarios.py --steps 4 --warmup 4 --no-graph > gpurun_out/${TAG}_ncu_ip.log 2>&1
echo "ncu ip rc=$?"; tail -3 gpurun_out/${TAG}_ncu_ip.log
