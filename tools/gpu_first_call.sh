#!/bin/bash
# first GPU call of a round: settle the open bf16 staged-vs-direct item, then the full parity suite, smoke and bench
TAG=${1:-r02a}
mkdir -p gpurun_out
timeout 300 python tools/wf_diag.py 3 > gpurun_out/wf_diag_$TAG.log 2>&1; echo "wf_diag exit $?" >> gpurun_out/wf_diag_$TAG.log
cat gpurun_out/wf_diag_$TAG.log
timeout 600 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
grep -E "passed|failed|FAILED|xfail|^\[|Error" gpurun_out/pytest_gpu_$TAG.log | tail -20
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
cut -c1-600 gpurun_out/bench_$TAG.json
