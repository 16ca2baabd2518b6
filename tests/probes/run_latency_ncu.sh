#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 \
  -k regex:'extract_blobs|check_wide|refine_kernel|gauss_newton|track_begin|track_finish|track_after' -s 350 -c 7 \
  -f -o gpurun_out/lat_kernels python tests/probes/latency_trace.py > gpurun_out/lat_ncu.log 2>&1
tail -3 gpurun_out/lat_ncu.log
ls -la gpurun_out/lat_kernels.ncu-rep
