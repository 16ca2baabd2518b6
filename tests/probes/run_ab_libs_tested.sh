#!/bin/bash
# usage: run_ab_libs_tested.sh lib ... : GPU test suite + headline stage times for each in-tree build
mkdir -p gpurun_out
for lib in "$@"; do
  export MPE_B200_LIB=$PWD/rpg_monocular_pose_estimator_b200/$lib
  python -m pytest tests -m gpu -x -q 2>&1 | tail -1
  python bench.py --no-extras --no-cpu --no-e2e --steps 10 > gpurun_out/ab_$lib.json 2> gpurun_out/ab_$lib.err
  python - "$lib" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/ab_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1],'value %.0f'%d['value'],'ms/step %.3f'%d['ms_per_step'],' '.join('%s %.4f'%(k['name'].split()[0],k['ms']) for k in d['kernels']))
PY
done
