#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY — builds the *reference's own* selective_scan CUDA path
# (c95yang/FocalNet kernels/selective_scan, "oflex" variant) for sm_100a straight from the
# sources where they lie under /root/reference.  No reference source is copied into this
# repo; only the built binary lands in oracle/_ref/ (git-ignored, travels with gpurun).
#
# Why not the reference's own build system: kernels/selective_scan/setup.py:58-64 hard-codes
# -gencode sm_70/sm_80/sm_90 SASS with no PTX, so its output cannot launch on a B200.  We run
# nvcc/g++ on the same three source files (setup.py:96-100) with the same flags
# (setup.py:109-125) and only swap the -gencode list.
#
# Output: oracle/_ref/selective_scan_cuda_oflex_ref.so  (a torch pybind module exposing
# fwd/bwd, selective_scan_oflex.cpp:360-363).  It is the GPU parity target named by the north
# star; it is imported only by tests/ and by bench.py's comparison leg.
set -euo pipefail
REF=${REF:-/root/reference/kernels/selective_scan}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
NAME=selective_scan_cuda_oflex_ref
if [ ! -d "$REF" ]; then
  echo "[build_ref] $REF absent (GPU box?) - keeping prebuilt $OUT/$NAME.so if any"; exit 0
fi
if [ -f "$OUT/$NAME.so" ] && [ "${FORCE:-0}" != "1" ]; then
  echo "[build_ref] $OUT/$NAME.so already built"; exit 0
fi
mkdir -p "$OUT/obj"
PY=${PYTHON:-python}
TORCH_INC=$($PY - <<'EOF'
import torch.utils.cpp_extension as c, sysconfig
print(" ".join("-I"+p for p in c.include_paths(device_type="cuda")) + " -I" + sysconfig.get_paths()["include"])
EOF
)
TORCH_LIB=$($PY -c "import torch,os;print(os.path.join(os.path.dirname(torch.__file__),'lib'))")
SRC="$REF/csrc/selective_scan"
COMMON="-O3 -std=c++17 -DTORCH_EXTENSION_NAME=$NAME -DTORCH_API_INCLUDE_EXTENSION_H -I$SRC $TORCH_INC"
NVFLAGS="-U__CUDA_NO_HALF_OPERATORS__ -U__CUDA_NO_HALF_CONVERSIONS__ -U__CUDA_NO_BFLOAT16_OPERATORS__ \
 -U__CUDA_NO_BFLOAT16_CONVERSIONS__ -U__CUDA_NO_BFLOAT162_OPERATORS__ -U__CUDA_NO_BFLOAT162_CONVERSIONS__ \
 --expt-relaxed-constexpr --expt-extended-lambda --use_fast_math -lineinfo \
 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC"
(
  nvcc $COMMON $NVFLAGS -c "$SRC/cusoflex/selective_scan_core_fwd.cu" -o "$OUT/obj/fwd.o" &
  nvcc $COMMON $NVFLAGS -c "$SRC/cusoflex/selective_scan_core_bwd.cu" -o "$OUT/obj/bwd.o" &
  g++ $COMMON -fPIC -I/usr/local/cuda/include -c "$SRC/cusoflex/selective_scan_oflex.cpp" -o "$OUT/obj/host.o" &
  wait
)
g++ -shared -o "$OUT/$NAME.so" "$OUT/obj/fwd.o" "$OUT/obj/bwd.o" "$OUT/obj/host.o" \
  -L"$TORCH_LIB" -L/usr/local/cuda/lib64 -Wl,-rpath,"$TORCH_LIB" \
  -lc10 -ltorch -ltorch_cpu -ltorch_python -lc10_cuda -ltorch_cuda -lcudart
rm -rf "$OUT/obj"
echo "[build_ref] built $OUT/$NAME.so"
