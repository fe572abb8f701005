#!/bin/bash
# A/B measurement of alternative builds of libb2s.so (same ABI): kernel times by FFT size through tools/size_sweep.py.
# Usage on the GPU box: bash tools/ab_libs.sh <tag> <lib>...   ("main" = rtl-sdr-scanner-cpp_b200/lib/libb2s.so); output: gpurun_out/<tag>_ab.txt
tag=$1; shift
mkdir -p gpurun_out
out=gpurun_out/${tag}_ab.txt
: > $out
for lib in "$@"; do
  if [ "$lib" = main ]; then path=rtl-sdr-scanner-cpp_b200/lib/libb2s.so; else path=$lib; fi
  for rep in 1 2; do
    echo "== $lib (run $rep)" >> $out
    B2S_LIB=$PWD/$path timeout 300 python tools/size_sweep.py ${SIZES:-4096 8192 16384 32768} >> $out 2>&1
  done
done
cat $out
