#!/bin/bash
# A/B of programmatic dependent launch along the step's kernel chain (NRF_PDL=0: plain stream order).
for e in "NRF_PDL=1" "NRF_PDL=0" "NRF_PDL=1" "NRF_PDL=0"; do
  echo "== $e"
  env $e python bench.py --quick --steps 200 --warmup 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'loss', d['final_loss'])"
done
for e in "NRF_PDL=1" "NRF_PDL=0"; do env $e python scripts/exp/render_ab.py | tail -1; done
