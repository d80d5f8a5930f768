#!/bin/bash
# A/B builds of the blend kernels: tools/build_variant.sh NAME "-DD2GS_BWD_WARPS=1 -DD2GS_BWD_BATCH=32 ..." [files...]
# -> dynamic-2dgs_b200/build/variants/libd2gs_NAME.so (selected at run time with D2GS_LIB=...; see tests/gpu_ab.py)
set -e
NAME=$1; FLAGS=$2; shift 2
FILES=${@:-raster_forward.cu raster_backward.cu}
HERE=$(cd "$(dirname "$0")/.." && pwd)
CSRC=$HERE/dynamic-2dgs_b200/csrc
OUT=$HERE/dynamic-2dgs_b200/build/variants
mkdir -p $OUT/obj_$NAME
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
OBJS=""
for f in c_api raster_forward raster_backward deform epilogue mlp loss optim knn gs3d tile_binning; do
  if [[ " $FILES " == *" $f.cu "* ]]; then
    $NVCC -O3 -std=c++17 $ARCH -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -ccbin /usr/bin/g++ $FLAGS \
      -Xptxas -v -c $CSRC/$f.cu -o $OUT/obj_$NAME/$f.o 2> $OUT/obj_$NAME/$f.ptxas.log || (cat $OUT/obj_$NAME/$f.ptxas.log; exit 1)
    OBJS="$OBJS $OUT/obj_$NAME/$f.o"
  else
    OBJS="$OBJS $CSRC/$f.o"
  fi
done
$NVCC $ARCH -shared -o $OUT/libd2gs_$NAME.so $OBJS -lcudart
grep -h -A1 "blend_fwd_kernel\|blend_bwd_kernel" $OUT/obj_$NAME/*.ptxas.log | grep "Used" | sed "s/^/[$NAME] /"
