#!/bin/bash
# On the GPU box: for every variant directory given, install its libraries and run variant_bench.py.
# usage: profiles/run_variants.sh <outfile> <variant>...
out=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  cp variants/$v/*.so torpedo_b200/lib/
  timeout 300 python profiles/variant_bench.py $v 20 2>&1 | tail -1 >> gpurun_out/$out
done
cat gpurun_out/$out
