#!/bin/bash
# Multi-GPU call (gpurun --gpus N): 2-rank pytest of the fused data-parallel optimiser + the bench at N GPUs (short).
# Usage: bash scripts/gpu_multi.sh <tag> <N> [steps]
TAG=${1:-multi}; N=${2:-2}; STEPS=${3:-50}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_lerf_host.py -q --no-header -rf -s --timeout 600 > $OUT/pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -12 $OUT/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $STEPS --warmup 5 > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err; echo "bench $N exit $?"; cat $OUT/bench_${N}gpu.json; tail -5 $OUT/bench_${N}gpu.err
