OUT=gpurun_out/$1; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 100 --warmup 10 > $OUT/bench_8gpu.json 2> $OUT/bench_8gpu.err; echo "bench8 exit $?"; grep "nerfpp_b200\]" $OUT/bench_8gpu.err | head -3
python scripts/debug/ab_print.py $OUT/bench_8gpu.json
