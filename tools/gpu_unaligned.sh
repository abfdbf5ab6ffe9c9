#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -k "tcgen05 or dense_video" 2>&1 | grep -E "passed|failed|FAILED" | tail -12
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | sed -e 's/.*"value": \([0-9.]*\).*"stage_ms_per_step": \({[^}]*}\).*/fps \1 stage \2/' | tail -1
CLASFV_UMMA_ALIGNED_TAPS=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | sed -e 's/.*"value": \([0-9.]*\).*"stage_ms_per_step": \({[^}]*}\).*/fps \1 stage \2/' | tail -1
