#!/bin/bash
mkdir -p gpurun_out
MPE_STEP_TRACE=1 python tests/probes/latency_trace.py > gpurun_out/lat_trace.out 2> gpurun_out/lat_trace.err
tail -4 gpurun_out/lat_trace.out
grep "step trace" gpurun_out/lat_trace.err | sed -n '60,64p'
