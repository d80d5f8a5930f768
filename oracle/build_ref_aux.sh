#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference side extensions from the sources where
# they lie under /root/reference into oracle/_ref/ (git-ignored, travels to the GPU box with gpurun):
#   simple_knn                  (submodules/simple-knn: distCUDA2)
#   diff_gaussian_rasterization (submodules/diff-gaussian-rasterization: the depth/alpha 3DGS variant)
# Same recipe as build_ref.sh: scratch copy under /tmp (their setup.py writes into the tree), only the built package
# lands in oracle/_ref/.  Used for golden generation and parity tests on the GPU box; never imported by the product path.
set -euo pipefail
OUT="$(cd "$(dirname "$0")" && pwd)/_ref"
export TORCH_CUDA_ARCH_LIST="10.0"
export MAX_JOBS=${MAX_JOBS:-8}
build_one() {  # <source dir> <package name>
  local REF="$1" PKG="$2"
  if [ ! -d "$REF" ]; then echo "[build_ref_aux] $REF not present — using prebuilt $OUT/$PKG if any"; return 0; fi
  if ls "$OUT/$PKG"/_C*.so >/dev/null 2>&1; then echo "[build_ref_aux] already built: $OUT/$PKG"; return 0; fi
  local TMP; TMP=$(mktemp -d /tmp/ref_aux.XXXXXX)
  cp -r "$REF"/. "$TMP"/
  rm -rf "$TMP/build" "$TMP/dist" "$TMP"/*.egg-info
  ( cd "$TMP" && NVCC_APPEND_FLAGS="-include cstdint -include cfloat" python setup.py build_ext --inplace >"$TMP/build.log" 2>&1 ) \
    || { tail -50 "$TMP/build.log"; return 1; }
  mkdir -p "$OUT/$PKG"
  if [ -f "$TMP/$PKG/__init__.py" ]; then cp "$TMP/$PKG/__init__.py" "$OUT/$PKG/"; else : > "$OUT/$PKG/__init__.py"; fi
  cp "$TMP/$PKG"/_C*.so "$OUT/$PKG/"
  echo "[build_ref_aux] built $PKG into $OUT"
  rm -rf "$TMP"
}
build_one /root/reference/submodules/simple-knn simple_knn
build_one /root/reference/submodules/diff-gaussian-rasterization diff_gaussian_rasterization
