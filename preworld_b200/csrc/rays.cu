// Ray records for the rendering head (input formats):
// pts2ray / get_rays of mmdet3d/datasets/ray.py:34-56 -- one 16-float row per
// labelled pixel in the layout NerfHead reads (datasets/ray.py:49-56;
// nerf_head.py:361-407): [x, y, depth, semantic, origin(3), direction(3),
// unit view direction(3), rgb(3)].  The reference builds it on the loader's CPU
// from ~10 small tensor ops per camera; here the loader can hand over the pixel
// lists and get the rays on the device.
#include "common.cuh"
#include "../../include/preworld_b200.h"

namespace {

__global__ void __launch_bounds__(256)
pts2ray_kernel(const float* __restrict__ coor, const float* __restrict__ depth,
               const float* __restrict__ seg, const float* __restrict__ rgb,
               const float* __restrict__ c2w, const float* __restrict__ K, long long n,
               float* __restrict__ rays) {
  const float k00 = __ldg(K + 0), k02 = __ldg(K + 2), k11 = __ldg(K + 4), k12 = __ldg(K + 5);
  float m[3][4];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) m[r][c] = __ldg(c2w + r * 4 + c);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float x = __ldg(coor + 2 * i), y = __ldg(coor + 2 * i + 1);
    // get_rays(i = x + 0.5, j = y + 0.5, inverse_y=True): ray.py:34-38
    const float d0 = __fdiv_rn(__fsub_rn(__fadd_rn(x, 0.5f), k02), k00);
    const float d1 = __fdiv_rn(__fsub_rn(__fadd_rn(y, 0.5f), k12), k11);
    float rd[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)   // torch.sum(dirs[..., None, :] * c2w[:3,:3], -1): no FMA
      rd[r] = __fadd_rn(__fadd_rn(__fmul_rn(d0, m[r][0]), __fmul_rn(d1, m[r][1])), m[r][2]);
    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(rd[0], rd[0]), __fmul_rn(rd[1], rd[1])),
                                      __fmul_rn(rd[2], rd[2])));
    float* o = rays + 16 * i;
    o[0] = x; o[1] = y;
    o[2] = __ldg(depth + i);
    o[3] = __ldg(seg + i);
    o[4] = m[0][3]; o[5] = m[1][3]; o[6] = m[2][3];
    o[7] = rd[0]; o[8] = rd[1]; o[9] = rd[2];
    o[10] = __fdiv_rn(rd[0], nrm); o[11] = __fdiv_rn(rd[1], nrm); o[12] = __fdiv_rn(rd[2], nrm);
    o[13] = __ldg(rgb + 3 * i); o[14] = __ldg(rgb + 3 * i + 1); o[15] = __ldg(rgb + 3 * i + 2);
  }
}

}  // namespace

PW_API int pw_pts2ray(const float* coor, const float* label_depth, const float* label_seg,
                      const float* label_img, const float* c2w, const float* cam_intrinsic,
                      long long n, float* rays, void* stream) {
  PW_REQUIRE(n >= 0);
  if (n == 0) return 0;
  PW_REQUIRE(coor && label_depth && label_seg && label_img && c2w && cam_intrinsic && rays);
  const int blocks = (int)min((long long)148 * 8, (n + 255) / 256);
  pts2ray_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(coor, label_depth, label_seg, label_img,
                                                           c2w, cam_intrinsic, n, rays);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}
