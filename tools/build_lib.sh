#!/bin/bash
# nvcc -> libtic.so with ptxas statistics in /tmp/ptxas.log (what __graft_entry__.build() runs, plus -Xptxas -v)
cd "$(dirname "$0")/.." && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -Xptxas -v $TIC_EXTRA_FLAGS -o tf_image_compression_b200/libtic.so tf_image_compression_b200/csrc/tic_api.cu -lcuda > /tmp/ptxas.log 2>&1
rc=$?; grep -E "error|rror:" /tmp/ptxas.log | head -20; echo "build rc=$rc"; exit $rc
