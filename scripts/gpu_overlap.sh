#!/bin/bash
# Overlapped data-parallel exchange (NRF_DP_OVERLAP=1): 2-rank correctness + bench A/B.  Usage: bash scripts/gpu_overlap.sh <tag> <N> [steps]
TAG=${1:-ov}; N=${2:-2}; STEPS=${3:-50}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_hash.py -q --no-header -k "level_split" --timeout 300 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_multi.py -q --no-header -rf -s --timeout 600 > $OUT/pytest_multi.log 2>&1; echo "pytest multi exit $?"; grep -E "MULTI_GPU_WORKER_OK|passed|failed|Error|assert" $OUT/pytest_multi.log | cut -c1-600 | head -12
for ov in 0 1; do
  NRF_DP_OVERLAP=$ov timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $STEPS --warmup 5 --quick > $OUT/bench_${N}gpu_ov$ov.json 2> $OUT/bench_${N}gpu_ov$ov.err; echo "bench ov=$ov exit $?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${N}gpu_ov$ov.json"))
    print("N=$N overlap=$ov value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "dp", d["dp_check"] and {k:d["dp_check"].get(k) for k in ("ok","multicast","overlap","max_abs_shadow_diff")}, "timeout", d["flags_timeout_after_timed_regions"], "loss", d["final_loss"])
except Exception as e:
    print("no bench line:", e); print(open("$OUT/bench_${N}gpu_ov$ov.err").read()[-1500:])
PY
done
