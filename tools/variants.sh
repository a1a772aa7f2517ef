#!/bin/bash
# Runs the scratch benches once per library variant in build_variants/ (MP2GPU_LIB override); GPU box only.
# usage: tools/variants.sh [quick]   -- "quick": Merkle-stage bench only
mkdir -p gpurun_out
for so in build_variants/*.so; do
  name=$(basename $so .so)
  echo "=== $name ==="
  if [ "$1" != "quick" ]; then MP2GPU_LIB=$PWD/$so timeout 300 python tools/quick_bench.py 2>&1 | tail -14; fi
  MP2GPU_LIB=$PWD/$so timeout 120 python tools/hash_bench.py 2>&1 | tail -6
done
