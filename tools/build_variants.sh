#!/bin/bash
# tuning builds of the library (compile-time knobs of the kernel): hisparse_b200/libhsb_<tag>.so
# usage: tools/build_variants.sh "pf3:-DHSB_PREFETCH=3" "pf5:-DHSB_PREFETCH=5" ...
set -e
cd "$(dirname "$0")/../hisparse_b200/csrc"
for v in "$@"; do
  tag="${v%%:*}"; flags="${v#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a $flags -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-O3,-Wall,-pthread -Xptxas -v \
    -shared -o ../libhsb_$tag.so capi.cu host_api.cpp spmv_kernels.cu gpu_format.cu cpsr_decode_gpu.cu synth_gpu.cu tile_format.cpp cpsr_decode.cpp -lpthread 2>&1 \
    | grep -A1 "spmv_tiles_kernel" | grep -E "registers|spill" | sed "s/^/$tag: /" || true
done
