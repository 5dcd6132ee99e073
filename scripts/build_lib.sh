#!/bin/bash
# nvcc (sm_100a) -> vcfdist_b200/libvcfdist_b200.so; extra args are passed to nvcc (e.g. -Xptxas -v)
R=$(cd "$(dirname "$0")/.." && pwd)
exec /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
  -I "$R/include" "$R/vcfdist_b200/csrc/vd_api.cu" "$R/vcfdist_b200/csrc/vd_finalize.cpp" -o "$R/vcfdist_b200/libvcfdist_b200.so" "$@"
