#!/bin/bash
# tools/build_split.sh NAME [EXTRA_NVCC_FLAGS] — experimental build of the WHOLE library with ENERGY and WATER as two kernels
# (NMP_SPLIT=1: the hand-off planes change the state layout, so nmp_lib.cu is rebuilt too) ->
# noahmp_b200/libnoahmp_b200_NAME.so (select with NOAHMP_B200_LIB=<path>).
set -e
cd "$(dirname "$0")/../noahmp_b200/csrc"
NAME=$1; shift
B=build_$NAME; mkdir -p $B
C="-std=c++17 -O3 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -I../../include -DNMP_SPLIT=1"
nvcc $C -DNMP_FASTMATH=1 --use_fast_math -DNMP_BLOCK=256 -DNMP_PHASE_SYNC=2 -DNMP_MINBLOCKS=2 $@ -Xptxas -v -c nmp_kernels_fast.cu -o $B/fast.o 2> $B/ptxas_fast.log &
nvcc $C -DNMP_BLOCK=256 -DNMP_PHASE_SYNC=1 -fmad=false -c nmp_kernels_parity.cu -o $B/parity.o &
nvcc $C -c nmp_lib.cu -o $B/lib.o &
wait
grep -A2 "land_kernel.*Li2ELi1ELi1ELi1E" $B/ptxas_fast.log | grep -E "Compiling|spill|Used" | cut -c1-200
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../libnoahmp_b200_$NAME.so $B/fast.o $B/parity.o $B/lib.o build/nmp_tables.o -ldl
echo built $NAME
