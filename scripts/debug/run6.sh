OUT=gpurun_out/$1; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_mlp_nerf.py -q --no-header -rf --timeout 300 --tb=short > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mlp_nerf --csv --log-file $OUT/nerf_launches.csv python scripts/debug/nerf_leg.py 2 > $OUT/ncu_l.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_nerf_bwd -s 2 -c 2 -f -o $OUT/prof_mlp_nerf_bwd python scripts/debug/nerf_leg.py 2 > $OUT/ncu_f.log 2>&1; echo "ncu full exit $?"
timeout 300 python scripts/debug/nerf_leg.py 10 > $OUT/leg.json 2>&1; cat $OUT/leg.json
tail -8 $OUT/pytest.log; grep mlp_nerf $OUT/nerf_launches.csv | cut -d, -f5,15- | tail -6
