#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; tail -2 gpurun_out/final_tests.log
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 300 gpurun_out/r02_bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_final.csv python bench.py --steps 1 --warmup 3 --batches-per-step 1 --no-extras --no-cpu --no-e2e > gpurun_out/r02_launches_bench_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'extract_blobs' -s 2 -c 1 -f -o gpurun_out/r02_final_k1b python profiles/profile_workload.py --batch 8192 --steps 1 --warmup 1 > gpurun_out/r02_final_k1b.log 2>&1
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_pose.py -q -x -k pooled_contour > gpurun_out/sanitizer_r02_memcheck_k1b_pooled.log 2>&1; echo "sanitizer rc $?"; tail -4 gpurun_out/sanitizer_r02_memcheck_k1b_pooled.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_final.json").read().strip().splitlines()[-1])
print("value",d["value"],d["ms_per_step"],"e2e",d["e2e"]["value"],"roof",d["roofline"]["frac"],d["roofline_fp64"]["frac"])
print([ (k["name"],round(k["ms"],4)) for k in d["kernels"]])
x=d["extra"]; print("tracking",x["tracking"]["value"],x["tracking"]["e2e"]["value"]); print({k:(v.get("value") if isinstance(v,dict) else v) for k,v in x.items()})
PY
