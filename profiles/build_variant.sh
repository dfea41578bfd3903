#!/bin/bash
# Build a named variant of the native libraries into variants/<name>/ (git-ignored; travels to the GPU box).
# usage: profiles/build_variant.sh <name> "<extra nvcc flags>"   e.g.  build_variant.sh t512 "-DTPDCU_SORT_THREADS_WORDS=512"
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants/$name
make -C torpedo_b200/csrc -j4 OUT=../../variants/$name EXTRA="$*" 2>&1 | grep -E "error|warning: " | head -20
cuobjdump --dump-resource-usage variants/$name/libtpdcu.so 2>/dev/null | grep -A1 -E "onesweep_kernelILi1|onesweep_ws|blend_kernel|preprocess_kernel|emit_kernel" | grep -oE "Function [A-Za-z0-9_]+|REG:[0-9]+|STACK:[0-9]+" | paste -sd" " | sed "s/Function /\n/g" | cut -c1-140
