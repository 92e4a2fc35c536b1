#!/bin/bash
# Build an experimental variant of one translation unit into variants/lib_<tag>.so (git-ignored, travels with gpurun).
# usage: tools/build_variant.sh <tag> "<-D flags>" [source path = ullsam_b200/csrc/attention.cu] [replaced unit = attention]
set -e
TAG=$1; FLAGS=$2; SRC=${3:-/root/repo/ullsam_b200/csrc/attention.cu}; UNIT=${4:-attention}
C=/root/repo/ullsam_b200/csrc
mkdir -p $C/build_var /root/repo/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -I$C $FLAGS -c $SRC -o $C/build_var/${UNIT}_$TAG.o
OBJS=$(ls $C/build/*.o | grep -v "build/${UNIT}.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o /root/repo/variants/lib_$TAG.so $OBJS $C/build_var/${UNIT}_$TAG.o -cudart static
echo built variants/lib_$TAG.so
