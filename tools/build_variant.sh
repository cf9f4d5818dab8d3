#!/bin/bash
# Tuning aid: build a variant of the library with other kernel constants, e.g.
#   bash tools/build_variant.sh tile4k -DR3D_SEL_TILE_PX=4096
# -> build_variants/libreal3d_b200_tile4k.so ; run with R3D_LIB_PATH=build_variants/libreal3d_b200_tile4k.so python bench.py
set -e
NAME=$1; shift
mkdir -p build_variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 --fmad=false -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC "$@" \
    -o build_variants/libreal3d_b200_${NAME}.so pcl_augmentation_b200/csrc/*.cu
echo build_variants/libreal3d_b200_${NAME}.so
