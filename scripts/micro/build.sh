#!/bin/sh
# Builds the standalone probes next to their sources (*.bin is git-ignored; the binaries travel with gpurun).
#   sh scripts/micro/build.sh && gpurun -- 'scripts/micro/i8_umma.bin; scripts/micro/ozaki_tile.bin; scripts/micro/ozaki_gemm.bin'
set -e
cd "$(dirname "$0")/../.."
for f in pivot_bench rot_bias i8_umma ozaki_tile ozaki_gemm ozaki_pipe; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I neo_ls_svm_b200/csrc \
       -o scripts/micro/$f.bin scripts/micro/$f.cu
done
ls -la scripts/micro/*.bin
