OUT=gpurun_out/$1; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_host.py tests/test_gpu_mlp_nerf.py -q --no-header -rf --timeout 300 --tb=short > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
timeout 300 python scripts/classic_train_bench.py 1024 20 > $OUT/classic_train.json 2> $OUT/classic_train.err; echo "classic exit $?"; cat $OUT/classic_train.json; tail -3 $OUT/classic_train.err
tail -25 $OUT/pytest.log
