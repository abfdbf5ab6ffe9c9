#!/bin/bash
# warp-fuse: parity subset, then ncu --set full of the staged kernel (bf16 and fp32), after a plain run of the same command
TAG=${1:-r02y}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -s -k "warp_fus or staged" > gpurun_out/pytest_wf_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_wf_$TAG.log
grep -E "passed|failed|FAILED|staged vs|Error|exit" gpurun_out/pytest_wf_$TAG.log | tail -12
for DT in bf16 fp32; do
timeout 300 python tools/bench_warp_fuse.py --once --dtypes $DT > gpurun_out/wf_plain_$DT.log 2>&1 &&
timeout 600 ncu --set full --import-source on --clock-control none -k regex:warp_fuse -s 1 -c 1 -o gpurun_out/prof_wf_${DT}_$TAG -f python tools/bench_warp_fuse.py --once --dtypes $DT > gpurun_out/pf_ncu_wf_$DT.log 2>&1
echo "wf full $DT exit $?"
done
