#!/bin/bash
# TEST / BASELINE INFRASTRUCTURE ONLY.  Compiles the reference's own CUDA source
# of bev_pool_v2 -- UNMODIFIED, from where it lies under /root/reference -- for
# sm_100a into oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).
# The file has no torch dependency: its host entry points
#   void bev_pool_v2(int c, int n_intervals, const float* depth, ...)      (:125)
#   void bev_pool_v2_grad(int c, int n_intervals, const float* out_grad, ...) (:133)
# are what mmdet3d/ops/bev_pool_v2/src/bev_pool.cpp:7-14 declares and calls.
# Used as (a) a second oracle for pw_bev_pool_v2 / pw_bev_pool_v2_grad on the
# GPU and (b) the GPU baseline "the kernel to beat" (BASELINE.md §2).
# The nerf extensions (models/nerf/cuda/*.cu) are torch extensions that need a
# source patch to build against torch 2.11 -- treated as unbuildable.
set -e
REF=${PREWORLD_REFERENCE_ROOT:-/root/reference}
SRC=$REF/mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu
HERE=$(cd "$(dirname "$0")" && pwd)
mkdir -p "$HERE/_ref"
${NVCC:-/usr/local/cuda/bin/nvcc} -O3 -gencode arch=compute_100a,code=sm_100a \
    -shared -Xcompiler -fPIC -o "$HERE/_ref/libbev_pool_v2_ref.so" "$SRC"
echo "built $HERE/_ref/libbev_pool_v2_ref.so"
