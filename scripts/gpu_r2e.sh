#!/bin/bash
TAG=${1:-r2e}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_lerf_train.py tests/test_gpu_lerf.py tests/test_gpu_hash.py -q --no-header -rf -x --timeout 300 > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -8 $OUT/pytest.log
python scripts/exp/lerf_train_prof.py 5 > $OUT/eager.json 2>$OUT/eager.err; cat $OUT/eager.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/exp/lerf_train_prof.py 2 > $OUT/ncu.log 2>&1
python scripts/launch_summary.py $OUT/launches.csv | head -30
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lerf_bwd_chain_kernel|mlp_nerf_bwd_dw_kernel|lerf_fwd_tc_kernel" -s 6 -c 4 -f -o $OUT/prof_lerf_bwd python scripts/exp/lerf_train_prof.py 2 > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
