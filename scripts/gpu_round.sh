#!/bin/bash
# One gpurun call: smoke, GPU parity tests, CUDA goldens from the reference kernels, a short bench, launch list.
# Usage (from the repo root on the GPU box): bash scripts/gpu_round.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "nproc=$(nproc)" >> $OUT/gpu.txt
timeout 600 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 300 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
timeout 600 python tests/golden/make_golden_cuda.py $OUT/golden > $OUT/golden.log 2>&1; echo "golden exit $?" | tee -a $OUT/golden.log
timeout 900 python bench.py --steps 100 --warmup 10 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?" | tee -a $OUT/bench.err
tail -5 $OUT/smoke.log; tail -30 $OUT/pytest_gpu.log; tail -3 $OUT/golden.log; cat $OUT/bench.json; tail -5 $OUT/bench.err
