// Stereo plane-sweep cost volume, fused: replaces DepthNet.gen_grid +
// calculate_cost_volumn (necks/view_transformer.py:546-604), i.e. the 64
// F.grid_sample launches over 4-channel groups, the [B*N,4,D*H,W] warped
// intermediates (~6 GB of traffic per frame), the |.|-sum, the bias mask and
// the softmax over D.  One warp per stereo pixel (cam, h4, w4): the current
// feature vector stays in registers, the D warped samples are gathered from
// the channels-last previous-frame feature (each bilinear corner is one
// contiguous C*4-byte row), costs are reduced with warp shuffles.
#include <stdlib.h>

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
__device__ __forceinline__ unsigned long long sub2p(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b,
                                                   unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// Kernel layout (round 2).  A stereo pixel (cam, h4, w4) is handled by WP warps, each
// owning 128 * Q channels (Q float4 per lane); C = 128 * WP * Q.
//   * the sampling geometry of the D depth bins (cell + 4 bilinear weights) is
//     computed once per pixel by all lanes of its warps and parked in shared memory;
//     the gather loop reads it back with two broadcast LDS per bin (round 1: six
//     shuffles per bin);
//   * the four corner rows of the bilinear cell stay in registers in PARITY SLOTS:
//     corner (x, y) lives in slot (y & 1, x & 1), and the four weights are stored
//     already permuted to slot order.  Stepping to a neighbouring cell therefore
//     overwrites exactly the two stale rows and moves nothing (round 1 rotated the
//     register rows: ~35 predicated MOV / SEL per bin, executed even for the 82 % of
//     the bins that stay in their cell -- the kernel was issue bound on them);
//   * per-bin partial sums are not reduced with shuffles (10 instructions per bin)
//     but stored to a per-warp [32][33] shared-memory tile and summed by rows once
//     per 32 bins (2 instructions per bin, fixed summation order);
//   * warp 0 of the pixel adds the partial costs of the other warps, the bias mask
//     and does the softmax over D.
constexpr int CV_MAX_D = 128;

template <int WP>
struct CvCfg {
  static constexpr int PIX = (8 / WP) > 0 ? (8 / WP) : 1;      // pixels per block
  static constexpr int WARPS = PIX * WP;
  static constexpr int THREADS = WARPS * 32;
  static constexpr size_t SMEM = (size_t)PIX * CV_MAX_D * 16            // gw
                                 + (size_t)WARPS * 32 * 33 * 4          // reduction tiles (the
                                                                        // cells alias them first)
                                 + (size_t)WARPS * CV_MAX_D * 4         // partial costs
                                 + (size_t)PIX * 16                     // zero flags (bits)
                                 + (size_t)PIX * CV_MAX_D * 16;         // corner rows to (re)load
};

template <int WP, int Q, int MINB>
__global__ void __launch_bounds__(CvCfg<WP>::THREADS, MINB)
cost_volume_kernel(const float* __restrict__ curr, const float* __restrict__ prev,
                   const float* __restrict__ cam, const float* __restrict__ xs,
                   const float* __restrict__ ys, const float* __restrict__ ds,
                   float* __restrict__ out, int out_ld, int n, int H, int W, int D,
                   float bias, float wi_m1, float hi_m1) {
  using Cfg = CvCfg<WP>;
  constexpr int CW = 128 * Q;                            // channels of one warp
  constexpr int C = CW * WP;
  extern __shared__ __align__(16) unsigned char cv_smem[];
  float4* gw_all = reinterpret_cast<float4*>(cv_smem);
  float* red_all = reinterpret_cast<float*>(gw_all + Cfg::PIX * CV_MAX_D);
  float* pc_all = red_all + Cfg::WARPS * 32 * 33;
  unsigned char* fl_all = reinterpret_cast<unsigned char*>(pc_all + Cfg::WARPS * CV_MAX_D);

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int pl = warp / WP, part = warp - pl * WP;       // pixel of the block, channel block
  const long long pix = (long long)blockIdx.x * Cfg::PIX + pl;
  const long long total = (long long)n * H * W;
  if (pix >= total) return;                              // the whole warp group leaves together
  const int w = (int)(pix % W);
  const int h = (int)((pix / W) % H);
  const int img = (int)(pix / ((long long)W * H));
  const float* cm = cam + (long long)img * PW_CV_CAM_FLOATS;
  float4* gw = gw_all + pl * CV_MAX_D;
  float* red = red_all + warp * 32 * 33;
  // cells of the bins: only needed while the geometry is set up, in the (still
  // unused) reduction tile of the pixel's first warp
  int2* gxy = reinterpret_cast<int2*>(red_all + (pl * WP) * 32 * 33);
  float* pcost = pc_all + warp * CV_MAX_D;
  unsigned char* flags = fl_all + pl * 16;
  int4* offs = reinterpret_cast<int4*>(fl_all + Cfg::PIX * 16) + pl * CV_MAX_D;
  auto group_sync = [&]() {
    if (WP == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + pl), "n"(WP * 32) : "memory");
  };

  // ---- sampling geometry of every depth bin, once per pixel ---------------------
  {
    const float fx = __ldg(xs + w) - cm[9], fy = __ldg(ys + h) - cm[10];
    for (int d = part * 32 + lane; d < D; d += WP * 32) {
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
      const float wnw = (x0f + 1.f - ix) * (y0f + 1.f - iy), wne = (ix - x0f) * (y0f + 1.f - iy);
      const float wsw = (x0f + 1.f - ix) * (iy - y0f), wse = (ix - x0f) * (iy - y0f);
      // clamp before the int cast so far-away samples cannot overflow
      const int x0 = (int)fminf(fmaxf(x0f, -2.f), (float)W + 1.f);
      const int y0 = (int)fminf(fmaxf(y0f, -2.f), (float)H + 1.f);
      // weights in slot order: slot (py, px) holds corner (x0 + (px ^ x0 & 1), y0 + (py ^ y0 & 1))
      const bool ox = x0 & 1, oy = y0 & 1;
      const float t0 = oy ? wsw : wnw, t1 = oy ? wse : wne;      // row of slot py = 0
      const float b0 = oy ? wnw : wsw, b1 = oy ? wne : wse;      // row of slot py = 1
      gw[d] = make_float4(ox ? t1 : t0, ox ? t0 : t1, ox ? b1 : b0, ox ? b0 : b1);
      gxy[d] = make_int2(x0, y0);
    }
    group_sync();
    // which slots change between bin d - 1 and bin d (bin 0: all four): the gather
    // loop gets, per bin, the pixel index of the row each slot has to load (>= 0),
    // -1 = the corner lies outside the image (zeros), -2 = keep the row it holds
    for (int d = part * 32 + lane; d < D; d += WP * 32) {
      const int2 g = gxy[d];
      const int2 gp = gxy[d > 0 ? d - 1 : 0];
      const int nx0 = g.x + (g.x & 1), nx1 = g.x + 1 - (g.x & 1);      // even / odd column
      const int ny0 = g.y + (g.y & 1), ny1 = g.y + 1 - (g.y & 1);      // even / odd row
      const bool first = d == 0;
      const bool cx0 = first || nx0 != gp.x + (gp.x & 1);
      const bool cx1 = first || nx1 != gp.x + 1 - (gp.x & 1);
      const bool cy0 = first || ny0 != gp.y + (gp.y & 1);
      const bool cy1 = first || ny1 != gp.y + 1 - (gp.y & 1);
      auto row = [&](int x, int y, bool changed) {
        if (!changed) return -2;
        return ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) ? y * W + x : -1;
      };
      offs[d] = make_int4(row(nx0, ny0, cx0 || cy0), row(nx1, ny0, cx1 || cy0),
                          row(nx0, ny1, cx0 || cy1), row(nx1, ny1, cx1 || cy1));
    }
  }
  float4 cur[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) cur[q] = pw_ldg4(curr + pix * C + part * CW + q * 128 + lane * 4);
  group_sync();

  // ---- gather (out-of-image corners are zeros: grid_sample's zero padding) ------------
  const float* pimg = prev + (long long)img * H * W * C + part * CW + lane * 4;
  float4 s00[Q], s01[Q], s10[Q], s11[Q];                 // parity slots (y & 1, x & 1)
  auto reload = [&](int row, float4 (&dst)[Q]) {         // row: warp-uniform
    if (row == -2) return;
    const float* r = pimg + (long long)row * C;
#pragma unroll
    for (int q = 0; q < Q; ++q)
      dst[q] = row >= 0 ? pw_ldg4(r + q * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  // channel C - 4 decides the bias mask: last warp, last float4 of lane 31
  const bool flag_owner = (part == WP - 1) && lane == 31;
  unsigned zero_bits = 0;                                // bit dl: warped channel C - 4 == 0
  float mc0 = 0.f, mc1 = 0.f, mc2 = 0.f, mc3 = 0.f;      // bins lane, lane+32, +64, +96
  float* red_w = red + lane;

  // (A run-structured loop -- corner reloads hoisted to the start of each run of bins
  // that share a cell, straight arithmetic inside -- was measured 17 % slower: the runs
  // are 5.5 bins long on average and the unrolled flat loop overlaps the LDS of the
  // next bins with the arithmetic of the current one.)
#pragma unroll 4
  for (int d = 0; d < D; ++d) {
    const int4 of = offs[d];
    const float4 wt = gw[d];
    if ((of.x & of.y & of.z & of.w) != -2) {             // warp-uniform; ~18 % of the bins
      reload(of.x, s00); reload(of.y, s01); reload(of.z, s10); reload(of.w, s11);
    }
    // packed fp32x2 arithmetic (FMUL2 / FFMA2 / FADD2): two channels per instruction
    const unsigned long long w00 = pack2(wt.x, wt.x), w01 = pack2(wt.y, wt.y),
                             w10 = pack2(wt.z, wt.z), w11 = pack2(wt.w, wt.w);
    float psum = 0.f;
    float ax_last = 1.f;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      unsigned long long a01 = mul2(pack2(s00[q].x, s00[q].y), w00);
      unsigned long long a23 = mul2(pack2(s00[q].z, s00[q].w), w00);
      a01 = fma2(pack2(s01[q].x, s01[q].y), w01, a01);
      a23 = fma2(pack2(s01[q].z, s01[q].w), w01, a23);
      a01 = fma2(pack2(s10[q].x, s10[q].y), w10, a01);
      a23 = fma2(pack2(s10[q].z, s10[q].w), w10, a23);
      a01 = fma2(pack2(s11[q].x, s11[q].y), w11, a01);
      a23 = fma2(pack2(s11[q].z, s11[q].w), w11, a23);
      if (q == Q - 1) { float t; unpack2(a01, ax_last, t); }
      float dx, dy, dz, dw;
      unpack2(sub2p(pack2(cur[q].x, cur[q].y), a01), dx, dy);
      unpack2(sub2p(pack2(cur[q].z, cur[q].w), a23), dz, dw);
      psum += ((fabsf(dx) + fabsf(dy)) + fabsf(dz)) + fabsf(dw);
    }
    const int dl = d & 31;
    red_w[dl * 33] = psum;
    zero_bits |= (ax_last == 0.f ? 1u : 0u) << dl;
    if (dl == 31 || d == D - 1) {       // sum the rows of this 32-bin group (lane = bin)
      const int grp = d >> 5;
      if (flag_owner) reinterpret_cast<unsigned*>(flags)[grp] = zero_bits;
      zero_bits = 0;
      __syncwarp();
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) sum += red[lane * 33 + j];
      if (grp == 0) mc0 = sum; else if (grp == 1) mc1 = sum; else if (grp == 2) mc2 = sum; else mc3 = sum;
      __syncwarp();
    }
  }

  // ---- combine the channel blocks, bias mask, softmax over D of -cost ----------------
  float my_cost[4] = {mc0, mc1, mc2, mc3};
  if (WP > 1) {
#pragma unroll
    for (int k = 0; k < 4; ++k) pcost[k * 32 + lane] = my_cost[k];
    group_sync();
    if (part != 0) return;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      for (int o = 1; o < WP; ++o) my_cost[k] += pcost[o * CV_MAX_D + k * 32 + lane];
  } else {
    __syncwarp();
  }
  if (bias != 0.f) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k * 32 + lane < D && ((reinterpret_cast<const unsigned*>(flags)[k] >> lane) & 1u))
        my_cost[k] += bias;
  }
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
      out[pix * out_ld + k * 32 + lane] = (k * 32 + lane < D) ? e[k] / s : 0.f;
}

template <int WP, int Q, int MINB>
int launch_cost_volume(const float* curr, const float* prev, const float* cam, const float* xs,
                       const float* ys, const float* ds, float* out, int out_ld, int n, int h,
                       int w, int d, float bias, float wi_m1, float hi_m1, cudaStream_t st) {
  using Cfg = CvCfg<WP>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(cost_volume_kernel<WP, Q, MINB>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::SMEM);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const long long pixels = (long long)n * h * w;
  const int blocks = pw_ceil_div(pixels, Cfg::PIX);
  cost_volume_kernel<WP, Q, MINB><<<blocks, Cfg::THREADS, Cfg::SMEM, st>>>(
      curr, prev, cam, xs, ys, ds, out, out_ld, n, h, w, d, bias, wi_m1, hi_m1);
  PW_LAUNCH_CHECK();
  return 0;
}

}  // namespace

PW_API int pw_cost_volume(const float* curr, const float* prev, const float* cam, const float* xs,
                          const float* ys, const float* ds, float* out, int out_ld, int n, int h,
                          int w, int c, int d, float bias, int img_h, int img_w, void* stream) {
  PW_REQUIRE(curr && prev && cam && xs && ys && ds && out);
  PW_REQUIRE(n > 0 && h > 0 && w > 0 && d > 0 && d <= CV_MAX_D && out_ld >= d && out_ld <= 128);
  PW_REQUIRE(c % 128 == 0 && c <= 512);
  PW_REQUIRE((long long)h * w < (1ll << 31));
  PW_REQUIRE(((uintptr_t)curr & 15) == 0 && ((uintptr_t)prev & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  float wi_m1 = (float)img_w - 1.f, hi_m1 = (float)img_h - 1.f;
  int rc;
#define PW_CV_ARGS curr, prev, cam, xs, ys, ds, out, out_ld, n, h, w, d, bias, wi_m1, hi_m1, st
  // PW_CV_VARIANT (experiments): 0 = default; for C = 256: 1 = two warps x 128 channels
  static const int variant = [] { const char* e = getenv("PW_CV_VARIANT"); return e ? atoi(e) : 0; }();
  switch (c / 128) {
    case 1: rc = launch_cost_volume<1, 1, 4>(PW_CV_ARGS); break;
    case 2: rc = variant == 1 ? launch_cost_volume<2, 1, 4>(PW_CV_ARGS)
                              : launch_cost_volume<1, 2, 3>(PW_CV_ARGS); break;
    case 3: rc = launch_cost_volume<1, 3, 2>(PW_CV_ARGS); break;
    default: rc = launch_cost_volume<2, 2, 3>(PW_CV_ARGS); break;
  }
#undef PW_CV_ARGS
  if (rc != 0) return rc;
  pw_count_launch(1);
  return 0;
}
