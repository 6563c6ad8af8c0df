#!/bin/bash
# GPU run 2: new tests (trajectory stream, stochastic), InvPendulum / BalancingRider parity after the register
# rewrite, config 4 and per-model benches, ncu of the InvPendulum agent kernel:  bash tools/gpu/r02_g2.sh <tag>
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-g2}
timeout 1700 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log | cut -c1-400
timeout 600 python tools/bench_scenarios.py --steps 50 > gpurun_out/${TAG}_scen.json 2> gpurun_out/${TAG}_scen.err
echo "scen rc=$?"; cut -c1-400 gpurun_out/${TAG}_scen.json
timeout 600 python tools/bench_scenarios.py --steps 50 --model balancingrider > gpurun_out/${TAG}_scen_br.json 2> gpurun_out/${TAG}_scen_br.err
echo "scen br rc=$?"; cut -c1-400 gpurun_out/${TAG}_scen_br.json
timeout 900 python tools/bench_models.py > gpurun_out/${TAG}_models.jsonl 2> gpurun_out/${TAG}_models.err
echo "models rc=$?"; cut -c1-300 gpurun_out/${TAG}_models.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:agent_kernel -s 8 -c 1 -f -o gpurun_out/${TAG}_agent_ip \
    python tools/bench_scenarios.py --steps 4 --warmup 4 --no-graph > gpurun_out/${TAG}_ncu_ip.log 2>&1
echo "ncu ip rc=$?"
