#!/bin/bash
for d in 0 1 17 49 51; do
  echo "== CLASFV_HEAD_DBG=$d"
  CLASFV_HEAD_DBG=$d timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | sed -e 's/.*"stage_ms_per_step": \({[^}]*}\).*/stage \1/' | tail -1
done
