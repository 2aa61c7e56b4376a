OUT=gpurun_out/$1; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 300 --tb=short > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 10 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
tail -12 $OUT/pytest_gpu.log
