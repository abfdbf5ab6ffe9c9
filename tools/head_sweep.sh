#!/bin/bash
for pf in 0 2 4 8; do
  echo "== CLASFV_HEAD_PF=$pf"
  CLASFV_HEAD_PF=$pf timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | sed -e 's/.*"stage_ms_per_step": \({[^}]*}\).*/stage \1/' | tail -1
done
