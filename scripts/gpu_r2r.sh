#!/bin/bash
# Round-2 GPU call: ray-grouped hash forward for the render path.  Usage: bash scripts/gpu_r2r.sh <tag>
TAG=${1:-r2r}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_hash.py tests/test_gpu_pipeline.py tests/test_gpu_render.py tests/test_gpu_lerf.py -q --no-header -rf --timeout 300 > $OUT/pytest.log 2>&1
echo "pytest exit $?"; tail -8 $OUT/pytest.log
for g in 1 4 8 16 32 64 128 256; do
  NRF_RENDER_RAY_GROUP=$g timeout 300 python scripts/exp/render_ab.py 2>&1 | tail -1 | tee -a $OUT/render_group.jsonl
done
NRF_RENDER_RAY_GROUP=32 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/render_launches.csv python scripts/exp/render_ab.py > /dev/null 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open("$OUT/render_launches.csv") if l.startswith('"')))
h = rows[0]; k = h.index("Kernel Name"); v = h.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    agg[r[k][:60]].append(float(r[v].replace(",", "")))
tot = sum(sum(x) for x in agg.values())
for n, x in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{n:60s} n={len(x):4d} mean_us={sum(x)/len(x)/1e3:9.1f} share={sum(x)/tot:.3f}")
PY
