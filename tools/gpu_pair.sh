#!/bin/bash
# CTA-pair convolution bring-up: conv parity (pair vs single vs F.conv3d), schedule bit-identity, then the bench stage split
TAG=${1:-r02p}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "cta_pair or tcgen05" > gpurun_out/pytest_pair_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_pair_$TAG.log
grep -E "passed|failed|FAILED|Error|exit|timed out|trap" gpurun_out/pytest_pair_$TAG.log | tail -12
if grep -q "pytest exit 0" gpurun_out/pytest_pair_$TAG.log; then
timeout 300 python -m pytest tests -m gpu -q -x -k "dense_video or bf16 or fp16" 2>&1 | grep -E "passed|failed|FAILED" | tail -5
for V in A CLASFV_UMMA_NO_PAIR; do
env $V=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary 2>&1 | sed -e 's/.*"value": \([0-9.]*\).*"stage_ms_per_step": \({[^}]*}\).*"frac": \([0-9.]*\), "traffic.*/'$V' fps \1 stage \2 frac \3/' | tail -1
done
fi
