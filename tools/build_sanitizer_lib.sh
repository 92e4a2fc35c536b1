#!/bin/bash
# Build variants/lib_sanitizer.so: the whole library with the mbarrier watchdog raised from 4 s to 20 min, for runs under
# compute-sanitizer (every memory access of an instrumented kernel is checked; the fused attention kernels then take seconds).
set -e
C=/root/repo/ullsam_b200/csrc
mkdir -p $C/build_san /root/repo/variants
for f in $C/*.cu; do
  b=$(basename $f .cu)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -I$C \
    -DULLSAM_MBAR_TIMEOUT_NS=1200000000000ull -c $f -o $C/build_san/$b.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o /root/repo/variants/lib_sanitizer.so $C/build_san/*.o -cudart static
echo built variants/lib_sanitizer.so
