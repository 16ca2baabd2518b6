#!/bin/bash
# Builds rpg_monocular_pose_estimator_b200/libmpe_b200.so (in-tree) for sm_100a.
# -fmad=false: every FP64 multiply and add rounds separately, like the reference's x86-64 -O3 build
# without -march=native (monocular_pose_estimator_lib/CMakeLists.txt:5-6).
set -e
cd "$(dirname "$0")"
OUT=../libmpe_b200.so
NVCC=${NVCC:-nvcc}
FLAGS="$EXTRA_NVCC_FLAGS -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Xcompiler -fPIC -Xcompiler -O2"
mkdir -p ../../build
pids=()
for f in k1_find_leds k2_p3p_sweep k3_validate_refine k4_tracking mpe_abi; do
  if [ ! -f ../../build/$f.o ] || [ $f.cu -nt ../../build/$f.o ] || [ mpe_internal.cuh -nt ../../build/$f.o ] || [ p3p_device.cuh -nt ../../build/$f.o ] || [ ../../include/mpe_b200.h -nt ../../build/$f.o ]; then
    rm -f ../../build/$f.o          # a failed compile must not leave a stale object for the link step
    $NVCC $FLAGS -c $f.cu -o ../../build/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait "$p" || { echo "nvcc failed (pid $p)" >&2; exit 1; }; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT ../../build/k1_find_leds.o ../../build/k2_p3p_sweep.o ../../build/k3_validate_refine.o ../../build/k4_tracking.o ../../build/mpe_abi.o -lcudart
echo "built $(readlink -f $OUT)"
