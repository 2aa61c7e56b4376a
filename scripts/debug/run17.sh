OUT=gpurun_out/$1; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 300 --tb=short > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
timeout 300 python scripts/classic_train_bench.py 1024 20 > $OUT/classic_train.json 2> $OUT/classic_train.err; echo "classic exit $?"; cat $OUT/classic_train.json; tail -3 $OUT/classic_train.err
timeout 600 python bench.py --steps 100 --warmup 10 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -3 $OUT/bench.err
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke.log
