#!/bin/bash
# One gpurun call that refreshes the ncu evidence under gpurun_out/<tag>/ (summaries are copied into profiles/ by hand afterwards).
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
# launch list of the default bench command (shares of the step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
# full captures: rendering ops inside the training step, the classic-NeRF training kernels
for K in composite_fwd_nb_kernel composite_bwd_kernel sample_pdf_merge_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/prof_$K \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_$K.log 2>&1; echo "ncu $K exit $?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_nerf -s 3 -c 4 -f -o $OUT/prof_mlp_nerf \
    python scripts/debug/nerf_leg.py 2 > $OUT/ncu_mlp_nerf.log 2>&1; echo "ncu mlp_nerf exit $?"
ls -la $OUT
