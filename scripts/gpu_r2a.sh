#!/bin/bash
# Round-2 first GPU call: hash launch-plan A/B, GPU parity tests, smoke, bench.
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1; echo "nproc=$(nproc)" >> $OUT/gpu.txt
for cfg in "NRF_HASH_PLAN=legacy NRF_HASH_SPLIT=0" "NRF_HASH_SPLIT=0" "NRF_HASH_PLAN=legacy" ""; do
  env $cfg timeout 300 python scripts/exp/hash_plan_ab.py >> $OUT/hash_plan_ab.jsonl 2>> $OUT/hash_plan_ab.err
done
cat $OUT/hash_plan_ab.jsonl
timeout 600 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/smoke.log
timeout 1800 python -m pytest tests -m gpu -q --no-header -rf --timeout 300 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -30 $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 100 --warmup 10 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
