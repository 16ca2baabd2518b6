set -x
mkdir -p gpurun_out
# 1. every kernel of one cold batch at the bench launch shape, full sections (one launch each)
ncu --set full --clock-control none --import-source on -s 14 -c 14 -o gpurun_out/r02_final_all python profiles/profile_workload.py --batch 8192 --steps 1 --warmup 1 > gpurun_out/r02_final_all.log 2>&1
# 2. launch list of the bench command itself (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 3 --batches-per-step 1 --no-extras --no-cpu --no-e2e > gpurun_out/r02_launches_bench.log 2>&1
# 3. launch list of tracking steps
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/r02_launches_tracking.csv python profiles/profile_tracking.py > gpurun_out/r02_launches_tracking.log 2>&1
ls -la gpurun_out/ | tail -8
