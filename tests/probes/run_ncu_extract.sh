#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'extract_blobs|blur_kernel' -s 2 -c 2 -f -o gpurun_out/r02_extract python profiles/profile_workload.py --batch 8192 --steps 1 --warmup 1 > gpurun_out/r02_extract.log 2>&1
tail -2 gpurun_out/r02_extract.log; ls -la gpurun_out/r02_extract.ncu-rep
