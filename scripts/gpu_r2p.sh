#!/bin/bash
# ncu evidence of the round-2 step: launch list of the default bench command + --set full of the hash kernels and the fused compositing tail
TAG=${1:-r2p}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python bench.py --steps 100 --warmup 10 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --quick > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hash_fwd_kernel|hash_bwd_kernel|composite_bwd_kernel" -s 30 -c 8 -f -o $OUT/prof_hash python bench.py --steps 2 --warmup 3 --no-cpu-baseline --quick > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
python scripts/launch_summary.py $OUT/launches.csv | head -30
