#!/bin/bash
# flow-staged warp-fuse bring-up: parity first, then the config-3 point in every ring variant
TAG=${1:-r02w}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -s -k "warp_fus or staged or long_video or fusion" > gpurun_out/pytest_wf_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_wf_$TAG.log
grep -E "passed|failed|FAILED|staged vs|Error|exit" gpurun_out/pytest_wf_$TAG.log | tail -12
run() { echo "== $1"; env $2 timeout 300 python tools/bench_warp_fuse.py --clips 256 --flow-px 0 4 --dtypes fp32 bf16 ${3} 2>&1 | tee -a gpurun_out/warp_fuse_variants_$TAG.jsonl | cut -c1-150; }
run default "A=1"
run wide1024 "CLASFV_WARP_FUSE_1024=1"
run slices4 "CLASFV_WARP_FUSE_SLICES=4"
run slices3 "CLASFV_WARP_FUSE_SLICES=3"
run old_noflows "CLASFV_WARP_FUSE_NO_FLOW_STAGING=1"
run default224 "A=1" "--size 224 --clips 64"
