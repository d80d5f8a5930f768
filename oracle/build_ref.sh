#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference rasterizer (CUDA extension) from the
# sources where they lie under /root/reference into oracle/_ref/ (git-ignored, travels to the
# GPU box with gpurun).  Nothing is copied into the repo: the reference's setup.py insists on
# writing into its own tree, and /root/reference is read-only, so the build runs on a scratch
# copy under /tmp and only the built package (python shim + _C.so) lands in oracle/_ref/.
#
# Used by: tests (golden generation + parity on the GPU box), bench.py --impl reference.
# Never imported by the product path.
set -euo pipefail
REF=${REF:-/root/reference/submodules/diff-surfel-rasterization}
OUT="$(cd "$(dirname "$0")" && pwd)/_ref"
if [ ! -d "$REF" ]; then
  echo "[build_ref] $REF not present (GPU box?) — using prebuilt oracle/_ref if any"; exit 0
fi
if [ -f "$OUT/diff_surfel_rasterization/__init__.py" ] && ls "$OUT"/diff_surfel_rasterization/_C*.so >/dev/null 2>&1; then
  echo "[build_ref] already built: $OUT"; exit 0
fi
TMP=$(mktemp -d /tmp/dsr_ref.XXXXXX)
cp -r "$REF"/. "$TMP"/
cd "$TMP"
# rasterizer_impl.h misses <cstdint> under gcc 13 -> inject the include, source untouched.
export NVCC_APPEND_FLAGS="-include cstdint"
export TORCH_CUDA_ARCH_LIST="10.0"
export MAX_JOBS=${MAX_JOBS:-8}
python setup.py build_ext --inplace >"$TMP/build.log" 2>&1 || { tail -50 "$TMP/build.log"; exit 1; }
mkdir -p "$OUT/diff_surfel_rasterization"
cp diff_surfel_rasterization/__init__.py "$OUT/diff_surfel_rasterization/"
cp diff_surfel_rasterization/_C*.so "$OUT/diff_surfel_rasterization/"
echo "[build_ref] built into $OUT"
rm -rf "$TMP"
