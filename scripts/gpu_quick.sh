#!/bin/bash
# Short gpurun call: GPU parity tests (+ optional extras).  Usage: bash scripts/gpu_quick.sh <tag> [goldens] [refbench] [bench] [ncu:<regex>]
TAG=${1:-q}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1; echo "nproc=$(nproc)" >> $OUT/gpu.txt
if [[ " $* " != *" nopytest "* ]]; then
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 300 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
fi
for what in "$@"; do
  case $what in
    goldens) timeout 600 python tests/golden/make_golden_cuda.py $OUT/golden > $OUT/golden.log 2>&1; echo "golden exit $?" | tee -a $OUT/golden.log;;
    refbench) timeout 600 python scripts/ref_cuda_bench.py 4096 30 > $OUT/ref_cuda_bench.json 2> $OUT/ref_cuda_bench.err; echo "ref cuda bench exit $?"; cat $OUT/ref_cuda_bench.json;;
    bench) timeout 900 python bench.py --steps 100 --warmup 10 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json; tail -3 $OUT/bench.err;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
             python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?";;
    refarm) timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref arm exit $?"; cat $OUT/bench_reference.json;;
    nopytest) ;;
    smoke) timeout 600 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke.log;;
    ncu:*) K=${what#ncu:}; timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/prof_$K \
             python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_$K.log 2>&1; echo "ncu $K exit $?";;
  esac
done
if [ -f $OUT/pytest_gpu.log ]; then tail -25 $OUT/pytest_gpu.log; fi
