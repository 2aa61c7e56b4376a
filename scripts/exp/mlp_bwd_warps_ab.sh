#!/bin/bash
# NeRFSmall backward: 12 warps / 192-row tiles (default) vs 8 warps / 128-row tiles
for e in "NRF_MLP_BWD_WARPS=12" "NRF_MLP_BWD_WARPS=8" "NRF_MLP_BWD_WARPS=12" "NRF_MLP_BWD_WARPS=8"; do
  echo "== $e"
  env $e python bench.py --quick --steps 100 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step', round(d['ms_per_step'],4), 'mlp_small_bwd', d['kernels_ms_per_step']['mlp_small_bwd'], 'loss', d['final_loss'])"
done
