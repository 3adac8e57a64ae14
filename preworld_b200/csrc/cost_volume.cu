// Stereo plane-sweep cost volume, fused: replaces DepthNet.gen_grid +
// calculate_cost_volumn (necks/view_transformer.py:546-604), i.e. the 64
// F.grid_sample launches over 4-channel groups, the [B*N,4,D*H,W] warped
// intermediates (~6 GB of traffic per frame), the |.|-sum, the bias mask and
// the softmax over D.  One warp per stereo pixel (cam, h4, w4): the current
// feature vector stays in registers, the D warped samples are gathered from
// the channels-last previous-frame feature (each bilinear corner is one
// contiguous C*4-byte row), costs are reduced with warp shuffles.
#include "common.cuh"
#include "../../include/preworld_b200.h"

namespace {

__device__ __forceinline__ float dot3(const float* m, float x, float y, float z) {
  float acc = m[0] * x;
  acc = fmaf(m[1], y, acc);
  return fmaf(m[2], z, acc);
}

__device__ __forceinline__ unsigned long long pack2(float x, float y) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& x, float& y) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b,
                                                   unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

template <int Q>   // float4 chunks per lane: C == 128*Q
__global__ void __launch_bounds__(256, Q <= 2 ? 2 : 1)
cost_volume_kernel(const float* __restrict__ curr, const float* __restrict__ prev,
                   const float* __restrict__ cam, const float* __restrict__ xs,
                   const float* __restrict__ ys, const float* __restrict__ ds,
                   float* __restrict__ out, int out_ld, int n, int H, int W, int C, int D,
                   float bias, float wi_m1, float hi_m1) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long total = (long long)n * H * W;
  if (warp >= total) return;
  const int w = (int)(warp % W);
  const int h = (int)((warp / W) % H);
  const int img = (int)(warp / ((long long)W * H));
  const float* cm = cam + (long long)img * PW_CV_CAM_FLOATS;

  float4 cur[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q)
    cur[q] = pw_ldg4(curr + warp * C + q * 128 + lane * 4);

  // frustum pixel (downsample-4 grid), image augmentation undone
  const float fx = __ldg(xs + w) - cm[9], fy = __ldg(ys + h) - cm[10];
  const int own_q = (C - 4) / 128, own_lane = ((C - 4) % 128) / 4;
  const float* pimg = prev + (long long)img * H * W * C;

  // ---- sampling geometry: lane l computes depth bins l, l+32, l+64, l+96 once
  // (cell + bilinear weights); the gather loop below fetches them by shuffle ----
  int gx0[4], gy0[4];
  float gw[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int d = min(k * 32 + lane, D - 1);
    float pz = __ldg(ds + d) - cm[11];
    float qx = dot3(cm + 0, fx, fy, pz), qy = dot3(cm + 3, fx, fy, pz), qz = dot3(cm + 6, fx, fy, pz);
    qx *= qz; qy *= qz;
    float sx = dot3(cm + 12, qx, qy, qz) + cm[21];
    float sy = dot3(cm + 15, qx, qy, qz) + cm[22];
    float sz = dot3(cm + 18, qx, qy, qz) + cm[23];
    bool neg = sz < 1e-3f;
    float ux = dot3(cm + 24, sx, sy, sz), uy = dot3(cm + 27, sx, sy, sz), uz = dot3(cm + 30, sx, sy, sz);
    ux = __fdiv_rn(ux, uz); uy = __fdiv_rn(uy, uz);
    float vx = fmaf(cm[34], uy, cm[33] * ux) + cm[37];
    float vy = fmaf(cm[36], uy, cm[35] * ux) + cm[38];
    float gx = __fdiv_rn(vx, wi_m1) * 2.f - 1.f, gy = __fdiv_rn(vy, hi_m1) * 2.f - 1.f;
    if (neg) { gx = -2.f; gy = -2.f; }
    // grid_sample(align_corners=True, padding zeros) on the [H,W] feature
    float ix = ((gx + 1.f) * 0.5f) * (float)(W - 1);
    float iy = ((gy + 1.f) * 0.5f) * (float)(H - 1);
    float x0f = floorf(ix), y0f = floorf(iy);
    gw[k][0] = (x0f + 1.f - ix) * (y0f + 1.f - iy);
    gw[k][1] = (ix - x0f) * (y0f + 1.f - iy);
    gw[k][2] = (x0f + 1.f - ix) * (iy - y0f);
    gw[k][3] = (ix - x0f) * (iy - y0f);
    // clamp before the int cast so far-away samples cannot overflow
    gx0[k] = (int)fminf(fmaxf(x0f, -2.f), (float)W + 1.f);
    gy0[k] = (int)fminf(fmaxf(y0f, -2.f), (float)H + 1.f);
  }

  // ---- gather.  The four corner rows of the current bilinear cell stay in
  // registers: consecutive depth bins mostly fall into the same or the
  // neighbouring cell of the epipolar line (out-of-image corners are zeros,
  // which is what grid_sample's zero padding contributes). ---------------------
  float4 c00[Q], c01[Q], c10[Q], c11[Q];
  int cx = -1000000, cy = -1000000;
  auto load_corner = [&](int x, int y, float4 (&dst)[Q]) {
    const bool ok = (unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H;
    const float* r = pimg + ((long long)y * W + x) * C + lane * 4;
#pragma unroll
    for (int q = 0; q < Q; ++q)
      dst[q] = ok ? pw_ldg4(r + q * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto copy = [&](float4 (&dst)[Q], const float4 (&src)[Q]) {
#pragma unroll
    for (int q = 0; q < Q; ++q) dst[q] = src[q];
  };

  float my_cost[4] = {0.f, 0.f, 0.f, 0.f};   // bins lane, lane+32, lane+64, lane+96
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k * 32 >= D) break;
    const int dn = min(32, D - k * 32);
    for (int dl = 0; dl < dn; ++dl) {
      const int x0 = __shfl_sync(0xffffffffu, gx0[k], dl);
      const int y0 = __shfl_sync(0xffffffffu, gy0[k], dl);
      const float wnw = __shfl_sync(0xffffffffu, gw[k][0], dl);
      const float wne = __shfl_sync(0xffffffffu, gw[k][1], dl);
      const float wsw = __shfl_sync(0xffffffffu, gw[k][2], dl);
      const float wse = __shfl_sync(0xffffffffu, gw[k][3], dl);
      if (x0 != cx || y0 != cy) {                      // warp-uniform
        if (y0 == cy && x0 == cx + 1) {
          copy(c00, c01); copy(c10, c11);
          load_corner(x0 + 1, y0, c01); load_corner(x0 + 1, y0 + 1, c11);
        } else if (y0 == cy && x0 == cx - 1) {
          copy(c01, c00); copy(c11, c10);
          load_corner(x0, y0, c00); load_corner(x0, y0 + 1, c10);
        } else if (x0 == cx && y0 == cy + 1) {
          copy(c00, c10); copy(c01, c11);
          load_corner(x0, y0 + 1, c10); load_corner(x0 + 1, y0 + 1, c11);
        } else if (x0 == cx && y0 == cy - 1) {
          copy(c10, c00); copy(c11, c01);
          load_corner(x0, y0, c00); load_corner(x0 + 1, y0, c01);
        } else {
          load_corner(x0, y0, c00); load_corner(x0 + 1, y0, c01);
          load_corner(x0, y0 + 1, c10); load_corner(x0 + 1, y0 + 1, c11);
        }
        cx = x0; cy = y0;
      }
      float part = 0.f;
      bool zero_flag = false;
      // packed fp32x2 arithmetic (FMUL2 / FFMA2 / FADD2): two channels per instruction
      const unsigned long long w2nw = pack2(wnw, wnw), w2ne = pack2(wne, wne),
                               w2sw = pack2(wsw, wsw), w2se = pack2(wse, wse);
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        unsigned long long a01 = mul2(pack2(c00[q].x, c00[q].y), w2nw);
        unsigned long long a23 = mul2(pack2(c00[q].z, c00[q].w), w2nw);
        a01 = fma2(pack2(c01[q].x, c01[q].y), w2ne, a01);
        a23 = fma2(pack2(c01[q].z, c01[q].w), w2ne, a23);
        a01 = fma2(pack2(c10[q].x, c10[q].y), w2sw, a01);
        a23 = fma2(pack2(c10[q].z, c10[q].w), w2sw, a23);
        a01 = fma2(pack2(c11[q].x, c11[q].y), w2se, a01);
        a23 = fma2(pack2(c11[q].z, c11[q].w), w2se, a23);
        float ax, ay, az, aw;
        unpack2(a01, ax, ay);
        unpack2(a23, az, aw);
        part += ((fabsf(cur[q].x - ax) + fabsf(cur[q].y - ay)) + fabsf(cur[q].z - az)) +
                fabsf(cur[q].w - aw);
        if (q == own_q && lane == own_lane) zero_flag = (ax == 0.f);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      zero_flag = __shfl_sync(0xffffffffu, (int)zero_flag, own_lane) != 0;
      if (bias != 0.f && zero_flag) part += bias;
      if (dl == lane) my_cost[k] = part;
    }
  }

  // softmax over D of -cost
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k * 32 + lane < D) m = fmaxf(m, -my_cost[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float e[4], s = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    e[k] = (k * 32 + lane < D) ? expf(-my_cost[k] - m) : 0.f;
    s += e[k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k * 32 + lane < out_ld)                      // channels [D, out_ld) are zero padding
      out[warp * out_ld + k * 32 + lane] = (k * 32 + lane < D) ? e[k] / s : 0.f;
}

}  // namespace

PW_API int pw_cost_volume(const float* curr, const float* prev, const float* cam, const float* xs,
                          const float* ys, const float* ds, float* out, int out_ld, int n, int h,
                          int w, int c, int d, float bias, int img_h, int img_w, void* stream) {
  PW_REQUIRE(curr && prev && cam && xs && ys && ds && out);
  PW_REQUIRE(n > 0 && h > 0 && w > 0 && d > 0 && d <= 128 && out_ld >= d && out_ld <= 128);
  PW_REQUIRE(c % 128 == 0 && c <= 512);
  long long warps = (long long)n * h * w;
  int blocks = pw_ceil_div(warps * 32, 256);
  cudaStream_t st = (cudaStream_t)stream;
  float wi_m1 = (float)img_w - 1.f, hi_m1 = (float)img_h - 1.f;
  switch (c / 128) {
    case 1: cost_volume_kernel<1><<<blocks, 256, 0, st>>>(curr, prev, cam, xs, ys, ds, out, out_ld, n, h, w, c, d, bias, wi_m1, hi_m1); break;
    case 2: cost_volume_kernel<2><<<blocks, 256, 0, st>>>(curr, prev, cam, xs, ys, ds, out, out_ld, n, h, w, c, d, bias, wi_m1, hi_m1); break;
    case 3: cost_volume_kernel<3><<<blocks, 256, 0, st>>>(curr, prev, cam, xs, ys, ds, out, out_ld, n, h, w, c, d, bias, wi_m1, hi_m1); break;
    default: cost_volume_kernel<4><<<blocks, 256, 0, st>>>(curr, prev, cam, xs, ys, ds, out, out_ld, n, h, w, c, d, bias, wi_m1, hi_m1); break;
  }
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}
