#!/bin/bash
for cfg in "MPE_K1B_POOL=2" "MPE_K1B_POOL=0" "MPE_K1B_POOL=4" "MPE_K1B_POOL=2" "MPE_K1B_POOL=0"; do
  env $cfg python bench.py --mode tracking --no-cpu --no-e2e --no-extras 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$cfg', 'value %.0f ms/step %.4f' % (d['value'], d['ms_per_step']))"
done
