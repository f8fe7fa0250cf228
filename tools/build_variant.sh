#!/bin/bash
# build_variant.sh <name> <nvcc -D flags...>: an A/B build of libb200ldm.so under audioldm_with_lora_b200/variants/
# (git-ignored; travels to the GPU box), selected at run time with B200LDM_LIB=<path>.
set -e
name=$1; shift
cd "$(dirname "$0")/../audioldm_with_lora_b200"
mkdir -p variants/obj_$name
for f in conv_gemm attention attention_bwd norm sampler train; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c csrc/$f.cu -o variants/obj_$name/$f.o &
done
wait
nvcc -shared -o variants/libb200ldm_$name.so variants/obj_$name/*.o -lcudart -gencode arch=compute_100a,code=sm_100a
rm -rf variants/obj_$name
echo variants/libb200ldm_$name.so
