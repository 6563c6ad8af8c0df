#!/bin/bash
# GPU run 5: suite after the agent-kernel changes, config 4 / per-model benches, wide-vs-narrow for quarter and eighth shards
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-g5}
timeout 1700 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log | cut -c1-300
for M in invpendulum balancingrider planarpoint twod; do
timeout 600 python tools/bench_scenarios.py --steps 50 --model $M > gpurun_out/${TAG}_scen_$M.json 2> gpurun_out/${TAG}_scen_$M.err
echo "scen $M rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_scen_$M.json")); print("   %s: %.4f ms/step  %.4g agent-steps/s  pair %.4f agent %.4f ms  hbm frac %.3f" % ("$M", d["ms_per_step"], d["value"], d["pair_kernel_ms"], d["agent_kernel_ms"], d["roofline"]["frac"]))
PY
done
timeout 900 python tools/bench_models.py > gpurun_out/${TAG}_models.jsonl 2> gpurun_out/${TAG}_models.err
echo "models rc=$?"; cut -c1-200 gpurun_out/${TAG}_models.jsonl
show() { python - <<PY
import json
d=json.load(open("$1")); r=d["roofline"]; a=d["roofline_agent_kernel"]
print("$2: ms/step %.4f | prep %.4f K1 %.4f agent %.4f ms | frac %.3f" % (d["ms_per_step"], r.get("prepare_kernel_ms", 0), r["kernel_ms"], a["kernel_ms"], r["frac"]))
PY
}
for W in 4 8; do for WIDE in 0 1; do
CSF_TILED_WIDE=$WIDE CSF_BENCH_EMULATE_WORLD=$W timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_w${W}_wide$WIDE.json 2> gpurun_out/${TAG}_w${W}_wide$WIDE.err; show gpurun_out/${TAG}_w${W}_wide$WIDE.json "1/$W shard wide=$WIDE"
done; done
