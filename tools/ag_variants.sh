#!/bin/bash
# A/B of library builds on BASELINE config 4 (tophat afterglow): tools/ag_variants.sh lib_a.so lib_b.so ...
for lib in "$@"; do
  REDBACK_B200_LIB=$(realpath $lib) python bench.py --samples 200000 --steps 1 --warmup 3 --paths on --only config4 \
    --parity-samples 100 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['paths'][0]
print('$lib', p['key'], '%.4g unit/s' % p['units_per_s'], p['parity'].get('max_rel'))"
done
