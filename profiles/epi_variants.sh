#!/bin/bash
# Builds libddpm_b200.so variants that differ only in the epilogue arithmetic switches of conv_epilogue.cuh
# (profiles/scratch/epi_variants/lib_<name>.so; run here, they travel to the GPU box with the snapshot) and, with
# `run`, times the seven epilogue variants of the CTA-pair halo conv with each library in ONE gpurun call.
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/profiles/scratch/epi_variants
PKG=$ROOT/polyp_image_generator_b200
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -cudart shared -I $ROOT/include"
declare -A V
V[base]="-DEPI_PACKED_ADDS=0 -DEPI_PACKED_GNACC=0 -DEPI_CSUM_PAIRS=0 -DEPI_SILU_FORM2=0"
V[adds]="-DEPI_PACKED_ADDS=1 -DEPI_PACKED_GNACC=0 -DEPI_CSUM_PAIRS=0 -DEPI_SILU_FORM2=0"
V[gnacc]="-DEPI_PACKED_ADDS=0 -DEPI_PACKED_GNACC=1 -DEPI_CSUM_PAIRS=0 -DEPI_SILU_FORM2=0"
V[pairs]="-DEPI_PACKED_ADDS=0 -DEPI_PACKED_GNACC=0 -DEPI_CSUM_PAIRS=1 -DEPI_SILU_FORM2=0"
V[silu2]="-DEPI_PACKED_ADDS=0 -DEPI_PACKED_GNACC=0 -DEPI_CSUM_PAIRS=0 -DEPI_SILU_FORM2=1"
V[all]="-DEPI_PACKED_ADDS=1 -DEPI_PACKED_GNACC=1 -DEPI_CSUM_PAIRS=1 -DEPI_SILU_FORM2=1"
if [ "${1:-build}" = "build" ]; then
  mkdir -p $OUT
  for n in "${!V[@]}"; do
    ( nvcc $FLAGS ${V[$n]} -c $PKG/csrc/conv_halo.cu -o $OUT/halo_$n.o &
      nvcc $FLAGS ${V[$n]} -c $PKG/csrc/conv_igemm.cu -o $OUT/igemm_$n.o & wait
      OBJS=$(ls $PKG/build/*.o | grep -v -e conv_halo.cu.o -e conv_igemm.cu.o)
      nvcc -shared -cudart shared -o $OUT/lib_$n.so $OBJS $OUT/halo_$n.o $OUT/igemm_$n.o -gencode arch=compute_100a,code=sm_100a
      rm -f $OUT/halo_$n.o $OUT/igemm_$n.o; echo built $n ) &
  done
  wait
else
  for n in base adds gnacc pairs silu2 all base; do
    echo "== $n"
    DDPM_B200_LIB=$OUT/lib_$n.so python $ROOT/profiles/bench_kernels.py epi --first 1 --iters 30 2>&1 | grep "^epi"
  done
fi
