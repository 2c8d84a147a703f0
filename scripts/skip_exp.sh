#!/bin/bash
# timing experiment: in-pipeline cost of each phase = full step time - step time with that phase's kernels dropped
for m in ${MASKS:-0 1 2 4 8 3 12}; do
  DBMM_SKIP=$m python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null > /tmp/skip_$m.json
  python -c "
import json,sys; d=json.load(open('/tmp/skip_$m.json')); print('skip mask $m:', round(d['us_per_sgd_step'],2), 'us/step')"
done
