#!/bin/bash
# Build libbtsdsp variants with a different -D set for resample.cu / kernels.cu into variants/<name>.so (git-ignored *.so).
# usage: tools/build_variants.sh name1 "-DX=1" name2 "-DX=2" ...   (run `python -m openbts_ttsou_b200.build` first: capi.o is reused)
set -e
cd "$(dirname "$0")/.."
C=openbts_ttsou_b200/csrc
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-fvisibility=hidden -cudart shared"
mkdir -p variants
while [ $# -ge 2 ]; do
  n=$1; d=$2; shift 2
  for s in resample kernels; do nvcc $F $d -c $C/$s.cu -o variants/${n}_$s.o & done; wait
  # same fence as build.py: the only packed FMA allowed is the multiply form (addend RZ)
  if cuobjdump -sass variants/${n}_resample.o variants/${n}_kernels.o | grep FFMA2 | grep -qv "RZ.F32 *;"; then echo "fused FFMA2 in $n"; exit 1; fi
  nvcc -shared -cudart shared -o variants/$n.so $C/capi.o variants/${n}_resample.o variants/${n}_kernels.o
  echo built variants/$n.so
done
