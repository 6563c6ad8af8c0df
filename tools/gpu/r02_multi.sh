#!/bin/bash
# multi-GPU: sharded == single-GPU checks (NCCL and fused peer exchange), then the bench: bash tools/gpu/r02_multi.sh <ngpus> [big]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
BIG=""
[ "$2" = "big" ] && BIG="--big"
for MODE in "" "--peer"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/check_sharded.py $MODE $BIG > gpurun_out/multi_check_${N}${MODE}.log 2>&1
  echo "check_sharded $MODE rc=$?"; grep -E "sharded_vs_single|Error|error" gpurun_out/multi_check_${N}${MODE}.log | cut -c1-420 | tail -6
done
show() { python - <<PY
import json
d=json.load(open("$1")); r=d["roofline"]; a=d["roofline_agent_kernel"]
print("$1: n_gpus %d ms/step %.4f value %.4g e2e %.4g | prep %.4f K1 %.4f agent %.4f ms | frac %.3f" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], r.get("prepare_kernel_ms", 0), r["kernel_ms"], a["kernel_ms"], r["frac"]))
PY
}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/multi_bench_$N.json 2> gpurun_out/multi_bench_$N.err
echo "bench rc=$?"; tail -3 gpurun_out/multi_bench_$N.err; show gpurun_out/multi_bench_$N.json
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/multi_bench_1.json 2> gpurun_out/multi_bench_1.err; show gpurun_out/multi_bench_1.json
