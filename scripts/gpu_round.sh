#!/bin/bash
# One gpurun call: smoke, GPU parity tests, CUDA goldens from the reference kernels, a short bench, launch list, ncu capture.
# Usage (from the repo root on the GPU box): bash scripts/gpu_round.sh [tag] [ncu-kernel-regex]
TAG=${1:-r1}
KREGEX=${2:-hash_fwd_kernel}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "nproc=$(nproc)" >> $OUT/gpu.txt
timeout 600 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 300 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
timeout 600 python tests/golden/make_golden_cuda.py $OUT/golden > $OUT/golden.log 2>&1; echo "golden exit $?" | tee -a $OUT/golden.log
timeout 900 python bench.py --steps 100 --warmup 10 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?" | tee -a $OUT/bench.err
timeout 300 python scripts/ref_cuda_bench.py 4096 30 > $OUT/ref_cuda_bench.json 2> $OUT/ref_cuda_bench.err; echo "ref cuda bench exit $?"
# every launch of 2 steady-state steps with its device time (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 4 -c 2 -f -o $OUT/prof_$KREGEX \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
tail -5 $OUT/smoke.log; tail -30 $OUT/pytest_gpu.log; tail -3 $OUT/golden.log; cat $OUT/bench.json; tail -5 $OUT/bench.err; cat $OUT/ref_cuda_bench.json
