#!/bin/bash
# 8-GPU: fused peer exchange check at world 8 (incl. the 65,536 crowd), bench at 8 for N = 65,536 and 1,048,576, 4 GPUs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/check_sharded.py --peer --big > gpurun_out/multi_check_${N}--peer.log 2>&1
echo "check_sharded --peer rc=$?"; grep -E "sharded_vs_single|Error|error" gpurun_out/multi_check_${N}--peer.log | cut -c1-520 | tail -6
show() { python - <<PY
import json
d=json.load(open("$1")); r=d["roofline"]; a=d["roofline_agent_kernel"]
print("$1: n_gpus %d ms/step %.4f value %.4g e2e %.4g | prep %.4f K1 %.4f agent %.4f ms | frac %.3f" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], r.get("prepare_kernel_ms", 0), r["kernel_ms"], a["kernel_ms"], r["frac"]))
PY
}
for G in $N 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $G --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/multi_bench_$G.json 2> gpurun_out/multi_bench_$G.err
echo "bench $G rc=$?"; tail -2 gpurun_out/multi_bench_$G.err | cut -c1-300; show gpurun_out/multi_bench_$G.json
done
CSF_BENCH_N=1048576 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/multi_bench_1M_$N.json 2> gpurun_out/multi_bench_1M_$N.err
echo "bench 1M rc=$?"; tail -2 gpurun_out/multi_bench_1M_$N.err | cut -c1-300; show gpurun_out/multi_bench_1M_$N.json
