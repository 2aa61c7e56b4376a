OUT=gpurun_out/$1; mkdir -p $OUT
timeout 900 python bench.py --steps 200 --warmup 20 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -2 $OUT/bench.err | cut -c1-200
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref exit $?"; head -c 400 $OUT/bench_reference.json
python scripts/debug/ab_print.py $OUT/bench.json
