#!/bin/bash
# Builds the native hardware checks (no Python at run time).  The binaries are git-ignored but travel with gpurun.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
CSRC="$ROOT/collaborative_distillation_b200/csrc"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
for name in validate_io profile_gram; do
  "$NVCC" -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I "$ROOT/include" -o "$HERE/$name" "$HERE/$name.cu" \
    -L "$CSRC" -lwctb -lwctb_io -lcudart -Xlinker -rpath -Xlinker '$ORIGIN/../../collaborative_distillation_b200/csrc'
done
echo built
