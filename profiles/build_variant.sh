#!/bin/bash
# Build a named variant of the native libraries into variants/<name>/ (git-ignored; travels to the GPU box).
# usage: profiles/build_variant.sh <name> "<extra nvcc flags>"   e.g.  build_variant.sh t512 "-DTPDCU_SORT_THREADS_WORDS=512"
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants/$name
make -C torpedo_b200/csrc -j4 OUT=../../variants/$name EXTRA="$*" 2>&1 | grep -E "onesweep_kernel<1>|error|warning: " -A2 | grep -E "registers|spill|error|warning" | head -20
