#!/bin/bash
# quick multi-GPU sanity of the final build: sharded == single-GPU through the peer exchange, then one bench line.  bash tools/gpu/r02_multi_quick.sh <ngpus>
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/check_sharded.py --peer > gpurun_out/quick_check_${N}.log 2>&1
echo "check_sharded --peer rc=$?"; grep -E "sharded_vs_single|Error|error" gpurun_out/quick_check_${N}.log | cut -c1-420 | tail -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/quick_bench_$N.json 2> gpurun_out/quick_bench_$N.err
echo "bench rc=$?"; tail -3 gpurun_out/quick_bench_$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/quick_bench_$N.json")); r=d["roofline"]; a=d["roofline_agent_kernel"]
print("n_gpus %d ms/step %.4f value %.4g e2e %.4g | prep %.4f K1 %.4f agent %.4f ms | frac %.3f" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], r.get("prepare_kernel_ms", 0), r["kernel_ms"], a["kernel_ms"], r["frac"]))
PY
