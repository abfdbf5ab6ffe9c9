#!/bin/bash
# conv + schedule parity, then the bench stage split
timeout 600 python -m pytest tests -m gpu -q -k "tcgen05 or dense_video or bf16" 2>&1 | grep -E "passed|failed|FAILED" | tail -12
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | sed -e 's/.*"value": \([0-9.]*\).*"stage_ms_per_step": \({[^}]*}\).*"e2e": {"value": \([0-9.]*\).*"frac": \([0-9.]*\), "traffic.*/fps \1 stage \2 e2e \3 frac \4/' | tail -1
