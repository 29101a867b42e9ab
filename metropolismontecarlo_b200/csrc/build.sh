#!/bin/sh
# Builds libmmc_b200.so in-tree for sm_100a (B200). Usage: sh build.sh [EXTRA="extra nvcc flags"]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
make -C "$HERE" -j"$(nproc)" "$@"
echo "built $HERE/../libmmc_b200.so"
