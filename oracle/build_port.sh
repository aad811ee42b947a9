#!/usr/bin/env bash
# Build the OpenMP C restatement used as the CPU baseline (test/baseline infrastructure, not product).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
mkdir -p "${HERE}/_build"
gcc -O3 -march=x86-64-v3 -fopenmp -fPIC -shared -o "${HERE}/_build/libref_port.so" "${HERE}/ref_port.c" -lm
echo "built ${HERE}/_build/libref_port.so"
