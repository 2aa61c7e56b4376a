OUT=gpurun_out/nb4; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_mlp_nerf.py tests/test_gpu_host.py -q --no-header -rf --timeout 300 --tb=short > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
timeout 300 python scripts/classic_train_bench.py 1024 20 > $OUT/classic_train.json 2> $OUT/classic_train.err; echo "classic exit $?"; cat $OUT/classic_train.json; tail -3 $OUT/classic_train.err
timeout 600 python bench.py --steps 100 --warmup 10 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
tail -30 $OUT/pytest.log
