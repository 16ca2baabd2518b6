for cfg in "256 2" "256 3" "128 5" "128 6" "128 8" "64 12"; do set -- $cfg; cd rpg_monocular_pose_estimator_b200/csrc; touch k2_p3p_sweep.cu; EXTRA_NVCC_FLAGS="-DMPE_K2_THREADS=$1 -DMPE_K2_MINBLOCKS=$2" ./build.sh > /dev/null 2>&1; cd ../..; echo -n "threads $1 minblocks $2: "; timeout 300 python bench.py --steps 6 --warmup 3 --batch 8192 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('p3p_sweep ms', round(d['kernels'][2]['ms'],3), 'value', round(d['value']))"; done
