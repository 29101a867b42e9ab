#!/bin/sh
# Builds libmmc_b200.so in-tree for sm_100a (B200). Usage: sh build.sh [extra nvcc flags]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libmmc_b200.so"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
     -Xcompiler -fPIC,-O2,-Wall -shared "$@" \
     -o "$OUT" "$HERE/mmc_api.cu" -lcudart
echo "built $OUT"
