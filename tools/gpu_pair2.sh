#!/bin/bash
TAG=${1:-r02r}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -s -k "cta_pair or tcgen05 or dense_video" > gpurun_out/pytest_pair_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_pair_$TAG.log
grep -E "passed|failed|FAILED|Error|exit|timed out|trap" gpurun_out/pytest_pair_$TAG.log | tail -8
timeout 120 python tools/pair_diag.py 2>&1 | cut -c1-120 | tail -6
timeout 200 python tools/conv_trace.py 200 bf16 > gpurun_out/conv_trace_pair_$TAG.txt 2>&1
CLASFV_UMMA_NO_PAIR=1 timeout 200 python tools/conv_trace.py 200 bf16 > gpurun_out/conv_trace_nopair_$TAG.txt 2>&1
head -3 gpurun_out/conv_trace_pair_$TAG.txt | cut -c1-150; head -3 gpurun_out/conv_trace_nopair_$TAG.txt | cut -c1-150
