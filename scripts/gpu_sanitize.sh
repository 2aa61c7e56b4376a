#!/bin/bash
# compute-sanitizer over the parity tests of the kernels added in the second half of round 2.  Usage: bash scripts/gpu_sanitize.sh <tag>
TAG=${1:-san}; OUT=gpurun_out/$TAG; mkdir -p $OUT
SEL='ray_grouped or coarse_row_reuse or per_ray_view_term or sampler_block or composite_random or fused_composite_huber or coarse_reuse_leaves or render_image_tiles or fused_render_entry or shipped_shape'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_hash.py tests/test_gpu_mlp.py tests/test_gpu_render.py tests/test_gpu_pipeline.py \
      -q --no-header -x --timeout 1400 -k "$SEL" > $OUT/$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|Invalid|hazard" $OUT/$tool.log | head -12
done
