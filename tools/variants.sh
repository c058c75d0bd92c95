#!/bin/bash
# Build-and-measure loop on the GPU box: tools/variants.sh <file.cu> "<flags A>" "<flags B>" ...  (each variant recompiles
# <file.cu> with EXTRA=<flags> and runs tools/quick_probe.py); the last build left in place is the LAST variant.
f=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  touch upside-md_b200/csrc/$f
  make -s -C upside-md_b200/csrc EXTRA="$v" > gpurun_out/variant_build.log 2>&1 || { echo "BUILD FAILED: $v"; tail -5 gpurun_out/variant_build.log; continue; }
  python tools/quick_probe.py "[$v]" 2>&1 | tail -3
done
