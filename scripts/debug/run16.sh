OUT=gpurun_out/$1; mkdir -p $OUT
timeout 300 python scripts/classic_train_bench.py 1024 20 > $OUT/classic_train.json 2> $OUT/classic_train.err; echo "classic exit $?"; cat $OUT/classic_train.json; tail -3 $OUT/classic_train.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/classic_launches.csv python scripts/classic_train_bench.py 1024 2 fused_tcgen05 > $OUT/ncu_l.log 2>&1; echo "ncu launches exit $?"
python scripts/launch_summary.py $OUT/classic_launches.csv 2>/dev/null | tail -40
