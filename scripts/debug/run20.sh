OUT=gpurun_out/$1; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_multi.py -q --no-header -rf --timeout 300 --tb=short > $OUT/pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -5 $OUT/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err; echo "bench2 exit $?"; tail -3 $OUT/bench_2gpu.err | cut -c1-300
python scripts/debug/ab_print.py $OUT/bench_2gpu.json
