#!/bin/bash
# Round-2 GPU call: evaluation reuse (importance-only fine forward).  Usage: bash scripts/gpu_r2q.sh <tag> [full] [bench]
TAG=${1:-r2q}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_hash.py tests/test_gpu_pipeline.py tests/test_gpu_render.py tests/test_gpu_host.py -q --no-header -rf -x --timeout 300 > $OUT/pytest_reuse.log 2>&1
echo "reuse pytest exit $?"; tail -25 $OUT/pytest_reuse.log
for env in "" "NRF_RENDER_REUSE=0" "NRF_RENDER_REUSE=0 NRF_HASH_SPLIT=0" "NRF_HASH_SPLIT=0"; do
  env $env timeout 300 python scripts/exp/render_ab.py 2>&1 | tail -1 | tee -a $OUT/render_ab.jsonl
done
for what in "$@"; do
  case $what in
    full) timeout 1800 python -m pytest tests -m gpu -q --no-header -rf --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 $OUT/pytest_gpu.log;;
    smoke) timeout 600 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/smoke.log;;
    bench) timeout 900 python bench.py --steps 100 --warmup 10 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json; tail -5 $OUT/bench.err;;
  esac
done
