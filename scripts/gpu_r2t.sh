#!/bin/bash
# Round-2 GPU call: full suite + smoke + bench.  Usage: bash scripts/gpu_r2t.sh <tag>
TAG=${1:-r2t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q --no-header -rf --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -12 $OUT/pytest_gpu.log
timeout 600 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/smoke.log
timeout 900 python bench.py --steps 100 --warmup 10 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -5 $OUT/bench.err
python - <<PY
import json
d = json.load(open("$OUT/bench.json"))
for k in ("value", "ms_per_step"): print(k, d[k])
print("e2e", d["e2e"]["value"])
print("kernels", d["kernels_ms_per_step"])
for k in ("train_shipped_shape", "train_as_shipped", "render", "render_lerf", "dropin_cpp", "reference_cuda"):
    v = d.get(k)
    print(k, {kk: vv for kk, vv in v.items() if kk in ("ms_per_step", "value", "ms_per_frame", "error", "loss", "kernels_per_step")} if v else v)
t = d.get("train_lerf"); print("train_lerf", t and {k: t[k] for k in ("value", "ms_per_step")})
PY
