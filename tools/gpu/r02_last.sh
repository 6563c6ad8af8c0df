#!/bin/bash
# last pass of the round: one knob check (items in flight per CTA), then the final records
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
one() { python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['value'], d['roofline']['kernel_ms'])"; }
one base; CSF_B200_LIB=$PWD/variants/libinflight3.so one inflight3; one base; CSF_B200_LIB=$PWD/variants/libinflight3.so one inflight3
bash tools/gpu/r02_final1.sh fin3
