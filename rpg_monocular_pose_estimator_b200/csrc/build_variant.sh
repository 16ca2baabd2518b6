#!/bin/bash
# A/B builds: build_variant.sh NAME "<extra nvcc flags>" -> ../libmpe_b200_NAME.so (select with MPE_B200_LIB=<path>)
set -e
cd "$(dirname "$0")"
NAME=$1; shift
FLAGS="$* -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Xcompiler -fPIC -Xcompiler -O2"
D=../../build/variant_$NAME
mkdir -p $D
for f in k1_find_leds k2_p3p_sweep k3_validate_refine k4_tracking mpe_abi; do nvcc $FLAGS -c $f.cu -o $D/$f.o & done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libmpe_b200_$NAME.so $D/*.o -lcudart
echo "built libmpe_b200_$NAME.so"
