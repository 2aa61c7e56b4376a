#!/bin/bash
# Round-2 GPU call: LeRF training kernels first run (+ optional full suite / bench).  Usage: bash scripts/gpu_r2b.sh <tag> [full] [bench] [smoke]
TAG=${1:-r2b}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_lerf_train.py -q --no-header -rf -x -s --timeout 300 > $OUT/pytest_lerf_train.log 2>&1; echo "lerf train pytest exit $?"; tail -40 $OUT/pytest_lerf_train.log
for what in "$@"; do
  case $what in
    full) timeout 1800 python -m pytest tests -m gpu -q --no-header -rf --timeout 300 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 $OUT/pytest_gpu.log;;
    smoke) timeout 600 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/smoke.log;;
    bench) timeout 900 python bench.py --steps 100 --warmup 10 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json; tail -5 $OUT/bench.err;;
  esac
done
