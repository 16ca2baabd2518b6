#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/quick_tests.log 2>&1; tail -3 gpurun_out/quick_tests.log
python bench.py --no-extras --no-cpu --no-e2e --steps 10 > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/quick_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'])
for k in d['kernels']: print('  ', k['name'], round(k['ms'],4))
PY
python tests/probes/latency_trace.py 2>&1 | grep "graphs True"
