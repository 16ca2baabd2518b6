# compute-sanitizer over the GPU parity tests (memcheck on everything that is not a long sweep, racecheck + synccheck on a cross-section)
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --leak-check no --error-exitcode 97 --log-file gpurun_out/r02_memcheck.log \
  python -m pytest tests/test_gpu_golden.py tests/test_gpu_tracking.py tests/test_gpu_find_leds.py tests/test_gpu_pose.py -m gpu -x -q \
  -k "not random_blobs and not reject_filter and not large_batch and not sigma" > gpurun_out/r02_memcheck_pytest.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_memcheck_pytest.txt
tail -3 gpurun_out/r02_memcheck_pytest.txt; tail -5 gpurun_out/r02_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 97 --log-file gpurun_out/r02_racecheck.log \
  python -m pytest tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/r02_racecheck_pytest.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_racecheck_pytest.txt
tail -3 gpurun_out/r02_racecheck_pytest.txt; tail -5 gpurun_out/r02_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 97 --log-file gpurun_out/r02_synccheck.log \
  python -m pytest tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/r02_synccheck_pytest.txt 2>&1
echo "synccheck rc=$?" >> gpurun_out/r02_synccheck_pytest.txt
tail -3 gpurun_out/r02_synccheck_pytest.txt; tail -5 gpurun_out/r02_synccheck.log
# launch list of one camera's tracking steps (plain launches)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_single_camera.csv python profiles/profile_latency.py 12 > /dev/null 2>&1
tail -30 gpurun_out/r02_launches_single_camera.csv | cut -c1-200
