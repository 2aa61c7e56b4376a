#!/bin/bash
# Multi-GPU call (gpurun --gpus N): 2-rank pytest of the fused data-parallel optimiser + the bench at N GPUs (short), multicast on / off.
# Usage: bash scripts/gpu_multi.sh <tag> <N> [steps] [ab]
TAG=${1:-multi}; N=${2:-2}; STEPS=${3:-50}; AB=${4:-}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -q --no-header -rf -s --timeout 600 > $OUT/pytest_multi.log 2>&1; echo "pytest multi exit $?"; grep -E "MULTI_GPU_WORKER_OK|passed|failed|Error" $OUT/pytest_multi.log | cut -c1-1200
run_bench() {  # $1 = suffix, rest = env
  local sfx=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $STEPS --warmup 5 > $OUT/bench_${N}gpu$sfx.json 2> $OUT/bench_${N}gpu$sfx.err; echo "bench $N $sfx exit $?"
  python - <<PY
import json
d=json.load(open("$OUT/bench_${N}gpu$sfx.json"))
print("N=$N$sfx value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "dp", d["dp_check"] and {k:d["dp_check"][k] for k in ("ok","multicast","max_abs_shadow_diff","entries_differing_by_more_than_5pct_of_lr")}, "strong", d["strong"] and (round(d["strong"]["value"]), round(d["strong"]["ms_per_step"],3)), "timeout", d["flags_timeout_after_timed_regions"])
t=d.get("train_lerf") or {}
print("   train_lerf", {k:t.get(k) for k in ("value","ms_per_step","error")}, (t.get("dp_check") or {}).get("ok"))
PY
}
run_bench "" NRF_DP_MULTICAST=auto
if [ -n "$AB" ]; then run_bench "_multicast" NRF_DP_MULTICAST=1; run_bench "_peerloads" NRF_DP_MULTICAST=0; fi
tail -3 $OUT/bench_${N}gpu.err
