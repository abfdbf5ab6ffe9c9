#!/bin/bash
for bc in 64 128 169; do
  echo "== batch_clips=$bc"
  timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --batch-clips $bc 2>&1 | sed -e 's/.*"value": \([0-9.]*\).*"stage_ms_per_step": \({[^}]*}\).*/fps \1 stage \2/' | tail -1
done
