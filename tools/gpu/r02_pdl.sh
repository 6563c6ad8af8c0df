#!/bin/bash
# programmatic dependent launch of the agent kernel: suite + A/B bench (CSF_PDL=0/1) at three crowd sizes and for shards
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu -x > gpurun_out/pdl_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pdl_pytest.log | cut -c1-300
show() { python - <<PY
import json
d=json.load(open("$1")); r=d["roofline"]; a=d["roofline_agent_kernel"]
print("$2: ms/step %.4f value %.4g e2e %.4g | K1 %.4f agent %.4f ms" % (d["ms_per_step"], d["value"], d["e2e"]["value"], r["kernel_ms"], a["kernel_ms"]))
PY
}
for PDL in 0 1 0 1; do
export CSF_PDL=$PDL
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/pdl${PDL}_full.json 2> gpurun_out/pdl${PDL}_full.err; show gpurun_out/pdl${PDL}_full.json "pdl=$PDL N=65536"
done
for PDL in 0 1; do
export CSF_PDL=$PDL
CSF_BENCH_N=4096 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/pdl${PDL}_4096.json 2> gpurun_out/pdl${PDL}_4096.err; show gpurun_out/pdl${PDL}_4096.json "pdl=$PDL N=4096"
CSF_BENCH_EMULATE_WORLD=8 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/pdl${PDL}_shard.json 2> gpurun_out/pdl${PDL}_shard.err; show gpurun_out/pdl${PDL}_shard.json "pdl=$PDL 1/8 shard"
CSF_BENCH_EMULATE_WORLD=2 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/pdl${PDL}_half.json 2> gpurun_out/pdl${PDL}_half.err; show gpurun_out/pdl${PDL}_half.json "pdl=$PDL half"
done
