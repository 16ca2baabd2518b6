#!/bin/bash
for lib in libmpe_b200.so libmpe_b200_late.so; do
for cfg in "MPE_PDL=1" "MPE_PDL=0"; do
  echo "== $lib $cfg"
  env $cfg MPE_B200_LIB=$PWD/rpg_monocular_pose_estimator_b200/$lib python tests/probes/latency_trace.py 2>&1 | grep "graphs"
done; done
