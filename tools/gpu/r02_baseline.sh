#!/bin/bash
# round-2 baseline: bench line + ncu full captures (K1 and the agent kernel) of the round-1 kernels
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_base.json 2> gpurun_out/r02_base.err
export CSF_BENCH_GRAPH=0
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pair_tiled -s 8 -c 1 -f -o gpurun_out/r02_k1_base \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:agent_kernel -s 8 -c 1 -f -o gpurun_out/r02_agent_base \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_agent.log 2>&1
ls -la gpurun_out | tail -8
