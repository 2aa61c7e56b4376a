#!/bin/bash
# headline numbers of a quick bench run: bash scripts/exp/quick_line.sh [bench args]
python bench.py --quick "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), '| e2e', round(d['e2e']['value']), 'ms', round(d['e2e']['ms_per_step'],4), '| loss', d['final_loss'], '| kernels', d['kernels_per_step'])
"
