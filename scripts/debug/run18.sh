OUT=gpurun_out/$1; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_pipeline.py -q --no-header -rf --timeout 120 --tb=short > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
tail -12 $OUT/pytest.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/bench_tc.json 2> $OUT/bench_tc.err; echo "bench tc exit $?"; tail -2 $OUT/bench_tc.err
NRF_MLP_BWD_DW=mma timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/bench_mma.json 2> $OUT/bench_mma.err; echo "bench mma exit $?"
python scripts/debug/ab_print.py $OUT/bench_tc.json $OUT/bench_mma.json
