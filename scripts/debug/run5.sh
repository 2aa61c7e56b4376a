OUT=gpurun_out/nb5; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_mlp_nerf.py tests/test_gpu_host.py -q --no-header -rf --timeout 300 --tb=short > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
timeout 300 python scripts/classic_train_bench.py 1024 20 > $OUT/classic_train.json 2> $OUT/classic_train.err; echo "classic exit $?"; cat $OUT/classic_train.json; tail -3 $OUT/classic_train.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mlp_nerf --csv --log-file $OUT/nerf_launches.csv python scripts/debug/nerf_leg.py 2 > $OUT/ncu_l.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_nerf_bwd -s 2 -c 2 -f -o $OUT/prof_mlp_nerf_bwd python scripts/debug/nerf_leg.py 2 > $OUT/ncu_f.log 2>&1; echo "ncu full exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_nerf_fwd_tc_kernel.*1 -s 1 -c 1 -f -o $OUT/prof_mlp_nerf_fwd_train python scripts/debug/nerf_leg.py 2 > $OUT/ncu_f2.log 2>&1; echo "ncu full2 exit $?"
tail -30 $OUT/pytest.log; grep mlp_nerf $OUT/nerf_launches.csv | cut -d, -f5,15- | tail -12
