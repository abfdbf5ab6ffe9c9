#!/bin/bash
# repeatability of the bench line: N short runs, prints value / e2e / per-step fusion-stage ms of each
N=${1:-6}
for i in $(seq 1 $N); do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | sed -e 's/.*"value": \([0-9.]*\).*"warp_fuse_ms_each_step": \(\[[^]]*\]\).*"e2e": {"value": \([0-9.]*\).*/value \1 fuse_ms \2 e2e \3/' | tail -1
done
