#!/bin/bash
# tools/build_variant.sh NAME "EXTRA_NVCC_FLAGS" — experimental build of the fast physics TU with other launch /
# math settings -> noahmp_b200/libnoahmp_b200_NAME.so (select with NOAHMP_B200_LIB=<path>). Used for the tuning
# sweeps recorded under profiles/.
set -e
cd "$(dirname "$0")/../noahmp_b200/csrc"
NAME=$1; shift
mkdir -p build
nvcc -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -I../../include \
  -Xptxas -v $@ -c nmp_kernels_fast.cu -o build/fast_$NAME.o 2> build/ptxas_$NAME.log
grep -E "registers|spill" build/ptxas_$NAME.log | grep -B1 -A0 "Used" | head -8
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../libnoahmp_b200_$NAME.so build/fast_$NAME.o \
  build/nmp_kernels_parity.o build/nmp_lib.o build/nmp_tables.o
echo built $NAME
