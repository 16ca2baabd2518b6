#!/bin/bash
# GPU tests, then the one-camera latency with the round's switches on and off
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/lat_tests.log 2>&1; tail -3 gpurun_out/lat_tests.log
for cfg in "MPE_PDL=1 MPE_MAPPED_IO=1" "MPE_PDL=0 MPE_MAPPED_IO=1" "MPE_PDL=1 MPE_MAPPED_IO=0" "MPE_PDL=0 MPE_MAPPED_IO=0"; do
  echo "== $cfg"
  env $cfg python tests/probes/latency_trace.py 2>&1 | grep "graphs True"
done
MPE_STEP_TRACE=1 python tests/probes/latency_trace.py 2>&1 | grep "step trace" | sed -n '60,61p'
