#!/bin/bash
# gpurun call for the classic-NeRF training kernels: layer-by-layer diagnosis, their parity tests, then (optionally) the whole GPU suite.
TAG=${1:-nb}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for n in 1000 20011; do
  timeout 300 python scripts/debug/nerf_bwd_check.py $n > $OUT/check_$n.log 2>&1; echo "check $n exit $?" | tee -a $OUT/check_$n.log
done
timeout 300 python scripts/debug/nerf_bwd_check.py 4096 0.1 > $OUT/check_xavier.log 2>&1; echo "check xavier exit $?" | tee -a $OUT/check_xavier.log
timeout 900 python -m pytest tests/test_gpu_mlp_nerf.py -q --no-header -rf --timeout 300 --tb=line > $OUT/pytest_nerf.log 2>&1; echo "pytest nerf exit $?" | tee -a $OUT/pytest_nerf.log
cat $OUT/check_1000.log; tail -12 $OUT/check_20011.log; tail -30 $OUT/check_xavier.log; tail -15 $OUT/pytest_nerf.log
if [[ " $* " == *" all "* ]]; then bash scripts/gpu_quick.sh $TAG "${@/all/}"; fi
