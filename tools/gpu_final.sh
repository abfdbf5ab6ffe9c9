#!/bin/bash
# round-end evidence: full GPU suite, then tools/gpu_profiles.sh (bench, launch list, ncu captures, sweeps, trace)
TAG=${1:-r02z}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
grep -E "passed|failed|FAILED|^ERROR|exit" gpurun_out/pytest_gpu_$TAG.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
bash tools/gpu_profiles.sh $TAG
