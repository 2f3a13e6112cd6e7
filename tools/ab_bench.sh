#!/bin/bash
# A/B of kernel variants: prints value + per-kernel ms for each environment setting given as arguments ("VAR=val VAR2=val" per argument)
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --no-cpu-baseline --no-gba --no-local-mapping --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$cfg', '| value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items() if v['ms_per_step']>0.02})"
done
