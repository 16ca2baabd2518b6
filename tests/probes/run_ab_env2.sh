#!/bin/bash
# usage: run_ab_env2.sh "ENV.. -- bench args" ...
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  envs="${cfg%%--*}"; args="${cfg#*--}"
  env $envs python bench.py --no-extras --no-cpu --no-e2e --steps 10 $args > gpurun_out/abenv_$i.json 2> gpurun_out/abenv_$i.err
  python - "$cfg" $i <<'PY'
import json,sys
d=json.loads(open('gpurun_out/abenv_%s.json'%sys.argv[2]).read().strip().splitlines()[-1])
print(sys.argv[1],'| value %.0f'%d['value'],'ms/step %.3f'%d['ms_per_step'],' '.join('%s %.4f'%(k['name'].split()[0],k['ms']) for k in d['kernels']))
PY
done
