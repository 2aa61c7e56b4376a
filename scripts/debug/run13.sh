OUT=gpurun_out/$1; mkdir -p $OUT
NRF_NERF_CLUSTER=1 timeout 120 python scripts/debug/nerf_leg.py 10 > $OUT/leg_cl1.json 2>&1; echo "leg cl1 exit $?"; cat $OUT/leg_cl1.json
timeout 120 python scripts/debug/nerf_leg.py 10 > $OUT/leg_cl2.json 2>&1; echo "leg cl2 exit $?"; cat $OUT/leg_cl2.json
timeout 600 python -m pytest tests/test_gpu_mlp_nerf.py tests/test_gpu_host.py -q --no-header -rf --timeout 120 --tb=short > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
tail -8 $OUT/pytest.log
