OUT=gpurun_out/$1; mkdir -p $OUT
timeout 120 python scripts/debug/nerf_leg.py 10 > $OUT/leg.json 2>&1; echo "leg exit $?"; cat $OUT/leg.json
NRF_NERF_CLUSTER=2 timeout 120 python scripts/debug/nerf_leg.py 10 > $OUT/leg_cl2.json 2>&1; echo "leg cl2 exit $?"; cat $OUT/leg_cl2.json
timeout 600 python -m pytest tests/test_gpu_mlp_nerf.py tests/test_gpu_host.py -q --no-header -rf --timeout 120 --tb=short > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
tail -5 $OUT/pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mlp_nerf --csv --log-file $OUT/nerf_launches.csv python scripts/debug/nerf_leg.py 2 > $OUT/ncu_l.log 2>&1; echo "ncu launches exit $?"
grep mlp_nerf $OUT/nerf_launches.csv | cut -d, -f5,15- | tail -4
