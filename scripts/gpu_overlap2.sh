#!/bin/bash
# Overlapped exchange: width of the side kernel and split level.  Usage: bash scripts/gpu_overlap2.sh <tag> <N> "<ctas list>" "<level list>"
TAG=${1:-ov}; N=${2:-2}; CT=${3:-"148 64"}; LV=${4:-"14"}
OUT=gpurun_out/$TAG; mkdir -p $OUT
run() {  # env...
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 --quick --global-rays 0 --no-cpu-baseline > $OUT/b.json 2> $OUT/b.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/b.json"))
    print("$*", "ms", round(d["ms_per_step"],4), "value", round(d["value"]), "dp_ok", d["dp_check"] and d["dp_check"]["ok"], "timeout", d["flags_timeout_after_timed_regions"], {k:v for k,v in d["kernels_ms_per_step"].items() if "adam" in k or "bwd" in k})
except Exception as e:
    print("$*", "FAILED", e); print(open("$OUT/b.err").read()[-800:])
PY
}
run NRF_DP_OVERLAP=0
for c in $CT; do for l in $LV; do run NRF_DP_OVERLAP=1 NRF_DP_OVERLAP_CTAS=$c NRF_DP_OVERLAP_LEVEL=$l; done; done
