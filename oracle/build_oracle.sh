#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY — compiles the plain-C oracle (oracle/ss2d_oracle.c) with gcc.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$HERE/_build"
gcc -O2 -fPIC -shared -fopenmp -std=c11 -o "$HERE/_build/libss2d_oracle.so" "$HERE/ss2d_oracle.c" -lm
echo "[build_oracle] built $HERE/_build/libss2d_oracle.so"
