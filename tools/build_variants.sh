#!/bin/bash
# Builds local-kernel launch-shape variants into variants/*.so (git-ignored; they travel to the GPU box) for tools/ab_local.sh.
# usage: tools/build_variants.sh "320 2" "256 3" "256 2" ...      (HYPER_THREADS HYPER_MIN_BLOCKS [extra -D flags])
set -e
cd "$(dirname "$0")/../admm-elastic-sca_b200"
make -j8 >/dev/null
mkdir -p ../variants
rm -f ../variants/*.so
ARCH="-gencode arch=compute_100a,code=sm_100a"
for v in "$@"; do
  set -- $v
  T=$1; B=$2; shift 2; EXTRA="$*"
  name=t${T}_b${B}$(echo "$EXTRA" | tr -d ' =-' | tr 'A-Z' 'a-z')
  /usr/local/cuda/bin/nvcc -std=c++17 -O3 -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fopenmp $ARCH -fmad=false -Xptxas -v \
      -DHYPER_THREADS=$T -DHYPER_MIN_BLOCKS=$B $EXTRA -c csrc/kernels_local.cu -o build/kl_$name.o 2>&1 | grep -A2 "NHModelELi5" | grep -E "Used|spill" | tr '\n' ' '
  echo " <- $name"
  OBJS=$(ls build/*.o | grep -v "kl_" | grep -v kernels_local.o)
  /usr/local/cuda/bin/nvcc -shared $ARCH -ccbin /usr/bin/g++ -Xcompiler -fopenmp -o ../variants/$name.so $OBJS build/kl_$name.o -lcudart -ldl
done
ls -la ../variants
