#!/bin/bash
# ncu --set full of the BA kernels of one short bench run; $1 = kernel regex, $2 = output stem, rest = env settings
PAT="$1"; OUT="$2"; shift 2
env "$@" timeout 500 ncu --set full --clock-control none --import-source on -k "regex:$PAT" -s 6 -c 4 -o gpurun_out/$OUT python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gba --no-local-mapping > gpurun_out/$OUT.log 2>&1
tail -1 gpurun_out/$OUT.log | cut -c1-200
