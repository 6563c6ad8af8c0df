#!/bin/bash
# 8-GPU box: scaling 1/2/4/8 at N = 65,536, N = 1,048,576 at 8 (and 1), config 4 at 8, sharded parity at world 8
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/sharded_check.jsonl
nvidia-smi -L | wc -l
show() { python - <<PY
import json
d=json.load(open("$1")); r=d["roofline"]; a=d["roofline_agent_kernel"]
print("$1: n_gpus %d ms/step %.4f value %.4g e2e %.4g | prep %.4f K1 %.4f agent %.4f ms | frac %.3f" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], r.get("prepare_kernel_ms", 0), r["kernel_ms"], a["kernel_ms"], r["frac"]))
PY
}
P=29600
for G in 8 4 2; do
P=$((P+1))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $P bench.py --gpus $G --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/mf_bench_$G.json 2> gpurun_out/mf_bench_$G.err
echo "bench $G rc=$?"; tail -2 gpurun_out/mf_bench_$G.err | cut -c1-300; show gpurun_out/mf_bench_$G.json
done
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/mf_bench_1.json 2> gpurun_out/mf_bench_1.err; show gpurun_out/mf_bench_1.json
CSF_BENCH_N=1048576 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/mf_bench_1M_8.json 2> gpurun_out/mf_bench_1M_8.err
echo "bench 1M x8 rc=$?"; tail -2 gpurun_out/mf_bench_1M_8.err | cut -c1-300; show gpurun_out/mf_bench_1M_8.json
if [ "$1" = "long" ]; then
CSF_BENCH_N=1048576 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/mf_bench_1M_1.json 2> gpurun_out/mf_bench_1M_1.err; show gpurun_out/mf_bench_1M_1.json
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 tools/bench_scenarios.py --scenarios 524288 --steps 50 > gpurun_out/mf_scen_8.json 2> gpurun_out/mf_scen_8.err
echo "scenarios x8 (weak: 65,536 per GPU) rc=$?"; cut -c1-330 gpurun_out/mf_scen_8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29615 tools/bench_scenarios.py --scenarios 65536 --steps 50 > gpurun_out/mf_scen_8_c4.json 2> gpurun_out/mf_scen_8_c4.err
echo "config 4 as named (65,536 scenarios over 8 GPUs) rc=$?"; cut -c1-330 gpurun_out/mf_scen_8_c4.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 tools/check_sharded.py --peer --big > gpurun_out/mf_check_8.log 2>&1
echo "check_sharded --peer --big rc=$?"; grep -E "sharded_vs_single" gpurun_out/mf_check_8.log | cut -c1-420 | tail -5
if [ "$1" = "long" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 tools/check_sharded.py > gpurun_out/mf_check_2_nccl.log 2>&1
echo "check_sharded nccl x2 rc=$?"; grep -E "sharded_vs_single" gpurun_out/mf_check_2_nccl.log | cut -c1-300 | tail -3
fi
