#!/bin/bash
# Build a VARIANT of libmsa_b200.so for an A/B run with scripts/ab_check.cu: one translation unit (msa_features.cu, or
# $SRC, e.g. SRC=msa_fusion_tc.cu) is recompiled with the given -D macros (or any nvcc flags) and linked with the other
# objects of the in-tree build; the result goes to scripts/ab/ (ignored
# by git, shipped to the GPU box by gpurun).  Prints the fp32 kernel's spills so a variant that spills is seen at once.
#   scripts/build_variant.sh base                      # the tree as it is
#   scripts/build_variant.sh try1 -DMSA_VAR_SOMETHING
#   scripts/build_variant.sh k1g3 -DMSA_K1_GROUPS=3                     # three wave-statistics load groups in flight
#   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/ab_check scripts/ab_check.cu -ldl
#   gpurun --timeout 60 -- 'timeout 45 ./scripts/ab_check scripts/ab/libmsa_base.so scripts/ab/libmsa_try1.so > gpurun_out/ab.json'
# (≈ 25 s of box time per call: no Python, no torch import.)
set -e
cd "$(dirname "$0")/../multimodal-sentiment-analyzer_b200/csrc"
name=$1; shift
src=${SRC:-msa_features.cu}
mkdir -p ../../scripts/ab build
tmp=$(mktemp -d)
nvcc -Wno-deprecated-gpu-targets -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Xptxas -v "$@" \
  -c "$src" -o "$tmp/var.o" 2> "$tmp/ptxas.txt" || { head -30 "$tmp/ptxas.txt"; exit 1; }
nvcc -Wno-deprecated-gpu-targets -shared -cudart static -o "../../scripts/ab/libmsa_$name.so" "$tmp/var.o" $(ls build/*.o | grep -v "${src%.cu}.o")
echo "$name ($src): $(grep -E 'spill' "$tmp/ptxas.txt" | sort | uniq -c | sed 's/^ *//' | tr '\n' ';')"
rm -rf "$tmp"
