// Generic fp32 implicit-GEMM convolution (1-D/2-D/3-D, channels-last) with a
// fused per-channel affine (folded BatchNorm / bias), residual add and
// activation epilogue.  SIMT FFMA path: bit-for-bit fp32 accumulate so the
// occupancy argmax parity (BASELINE.json north_star) holds; the tcgen05 path
// for the image backbone lives in conv_umma.cu.
//
// Replaces on the reference path: every torch.nn.Conv2d / Conv3d / Linear +
// BatchNorm + ReLU group (cuDNN/cuBLAS calls) of
//   mmdet ResNet, necks/fpn.py:154-203, necks/view_transformer.py:473-638,
//   backbones/resnet.py:88-184, necks/lss_fpn.py:120-148,
//   detectors/preworld.py:72-105, heads/occupancy_head.py:81-177,
//   detectors/preworld_temporal_traj.py:119-150.
//
// GEMM view: M = N*OD*OH*OW output positions, N = Cout, K = KD*KH*KW*Cin.
// A[m][k] is gathered on the fly (zero for padding), B = weights [K][Cout].
#include "common.cuh"
#include "../../include/preworld_b200.h"

namespace {

constexpr int BK = 16;
constexpr int NTHREADS = 256;

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(NTHREADS)
conv_igemm_kernel(const pw_conv_desc p, const float* __restrict__ x,
                  const float* __restrict__ w, const float* __restrict__ scale,
                  const float* __restrict__ bias, const float* __restrict__ res,
                  float* __restrict__ y) {
  static_assert((BM / TM) * (BN / TN) == NTHREADS, "tile/thread mismatch");
  static_assert(TM == 8 && (TN == 4 || TN == 8), "thread tile");
  constexpr int A_SLOTS = BM * (BK / 4) / NTHREADS;           // float4 per thread
  constexpr int B_CHUNKS = BK * BN / 4;                       // float4 per tile
  constexpr int B_SLOTS = (B_CHUNKS + NTHREADS - 1) / NTHREADS;
  static_assert(A_SLOTS >= 1, "BM too small");

  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int M = p.n * p.od * p.oh * p.ow;
  const int K = p.kd * p.kh * p.kw * p.cin;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // ---- per-thread gather rows (fixed over the K loop) ----
  // chunk c = tid + s*NTHREADS: row = c / 4, kc = c % 4 -> four consecutive
  // lanes read 64 contiguous bytes of one input pixel.
  const int kc = tid & 3;
  int a_base[A_SLOTS];      // pixel offset of (n, iz0, iy0, ix0) in elements / in_ld
  int a_iz0[A_SLOTS], a_iy0[A_SLOTS], a_ix0[A_SLOTS];
  bool a_ok[A_SLOTS];
#pragma unroll
  for (int s = 0; s < A_SLOTS; ++s) {
    int row = (tid >> 2) + s * (NTHREADS / 4);
    int m = m0 + row;
    a_ok[s] = m < M;
    int mm = a_ok[s] ? m : 0;
    int ox = mm % p.ow; mm /= p.ow;
    int oy = mm % p.oh; mm /= p.oh;
    int oz = mm % p.od; int n = mm / p.od;
    a_iz0[s] = oz * p.sd - p.pd;
    a_iy0[s] = oy * p.sh - p.ph;
    a_ix0[s] = ox * p.sw - p.pw;
    a_base[s] = n * p.d;
  }

  float4 a_reg[A_SLOTS];
  float4 b_reg[B_SLOTS];

  auto load_tiles = [&](int kt) {
    // A: gathered input
    int k = kt * BK + kc * 4;
    bool kok = k < K;
    int tap = kok ? k / p.cin : 0;
    int ci = k - tap * p.cin;
    int kx = tap % p.kw; int t2 = tap / p.kw;
    int ky = t2 % p.kh; int kz = t2 / p.kh;
    kz *= p.dd; ky *= p.dh; kx *= p.dw;
#pragma unroll
    for (int s = 0; s < A_SLOTS; ++s) {
      int iz = a_iz0[s] + kz, iy = a_iy0[s] + ky, ix = a_ix0[s] + kx;
      bool ok = kok && a_ok[s] && (unsigned)iz < (unsigned)p.d &&
                (unsigned)iy < (unsigned)p.h && (unsigned)ix < (unsigned)p.w;
      if (ok) {
        size_t pix = ((size_t)(a_base[s] + iz) * p.h + iy) * p.w + ix;
        a_reg[s] = pw_ldg4(x + pix * p.in_ld + ci);
      } else {
        a_reg[s] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    // B: weights [K][w_ld]
#pragma unroll
    for (int s = 0; s < B_SLOTS; ++s) {
      int c = tid + s * NTHREADS;
      int kr = c / (BN / 4), nc = (c % (BN / 4)) * 4;
      int kk = kt * BK + kr;
      bool ok = (c < B_CHUNKS) && kk < K && (n0 + nc) < p.w_ld;
      b_reg[s] = ok ? pw_ldg4(w + (size_t)kk * p.w_ld + n0 + nc)
                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };

  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int s = 0; s < A_SLOTS; ++s) {
      int row = (tid >> 2) + s * (NTHREADS / 4);
      // XOR swizzle of the 8-row group by kc keeps the four lanes that share
      // a row on different banks (see read side below).
      int col = row ^ (kc << 3);
      As[buf][kc * 4 + 0][col] = a_reg[s].x;
      As[buf][kc * 4 + 1][col] = a_reg[s].y;
      As[buf][kc * 4 + 2][col] = a_reg[s].z;
      As[buf][kc * 4 + 3][col] = a_reg[s].w;
    }
#pragma unroll
    for (int s = 0; s < B_SLOTS; ++s) {
      int c = tid + s * NTHREADS;
      if (c < B_CHUNKS) {
        int kr = c / (BN / 4), nc = (c % (BN / 4)) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][kr][nc]) = b_reg[s];
      }
    }
  };

  const int tx = tid % (BN / TN);
  const int ty = tid / (BN / TN);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();

  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const int grp = ty ^ (k >> 2);          // undo the store-side swizzle
      float a[TM], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][grp * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][grp * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN + j]);
        b[j] = bv.x; b[j + 1] = bv.y; b[j + 2] = bv.z; b[j + 3] = bv.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue: affine, residual, activation ----
  const int col0 = n0 + tx * TN;
  float sc[TN], bi[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    int c = col0 + j;
    sc[j] = (scale != nullptr && c < p.cout) ? __ldg(scale + c) : 1.f;
    bi[j] = (bias != nullptr && c < p.cout) ? __ldg(bias + c) : 0.f;
  }
  const int act_end = p.act_channels > 0 ? p.act_channels : p.cout;
  // float4 epilogue only when the (possibly channel-sliced) output and
  // residual rows are 16-byte aligned
  const bool vec_ok = ((p.out_ld & 3) == 0) && (col0 + TN <= p.cout) &&
                      ((reinterpret_cast<uintptr_t>(y) & 15) == 0) &&
                      (res == nullptr || ((p.res_ld & 3) == 0 &&
                                          (reinterpret_cast<uintptr_t>(res) & 15) == 0));
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= M) continue;
    float v[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) v[j] = fmaf(acc[i][j], sc[j], bi[j]);
    if (vec_ok) {
      if (res != nullptr) {
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          float4 r = pw_ldg4(res + (size_t)m * p.res_ld + col0 + j);
          v[j] += r.x; v[j + 1] += r.y; v[j + 2] += r.z; v[j + 3] += r.w;
        }
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const int a = (col0 + j < act_end) ? p.act : PW_ACT_NONE;   // act_end % 4 == 0
        float4 o = make_float4(pw_activate(v[j], a), pw_activate(v[j + 1], a),
                               pw_activate(v[j + 2], a), pw_activate(v[j + 3], a));
        *reinterpret_cast<float4*>(y + (size_t)m * p.out_ld + col0 + j) = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        int c = col0 + j;
        if (c < p.cout) {
          float t = v[j];
          if (res != nullptr) t += __ldg(res + (size_t)m * p.res_ld + c);
          y[(size_t)m * p.out_ld + c] = pw_activate(t, c < act_end ? p.act : PW_ACT_NONE);
        }
      }
    }
  }
}

template <int BM, int BN, int TM, int TN>
int launch(const pw_conv_desc& p, const float* x, const float* w, const float* scale,
           const float* bias, const float* res, float* y, cudaStream_t st) {
  long long M = (long long)p.n * p.od * p.oh * p.ow;
  dim3 grid(pw_ceil_div(M, BM), pw_ceil_div(p.cout, BN));
  conv_igemm_kernel<BM, BN, TM, TN><<<grid, NTHREADS, 0, st>>>(p, x, w, scale, bias, res, y);
  PW_LAUNCH_CHECK();
  pw_count_launch(1);
  return 0;
}

}  // namespace

PW_API int pw_conv_fwd(const pw_conv_desc* d, const float* x, const float* w,
                       const float* scale, const float* bias, const float* residual,
                       float* y, void* stream) {
  PW_REQUIRE(d && x && w && y);
  const pw_conv_desc& p = *d;
  PW_REQUIRE(p.n > 0 && p.d > 0 && p.h > 0 && p.w > 0 && p.cin > 0 && p.cout > 0);
  PW_REQUIRE((p.cin & 3) == 0 && (p.in_ld & 3) == 0 && p.in_ld >= p.cin);
  PW_REQUIRE((p.w_ld & 3) == 0 && p.w_ld >= p.cout && p.out_ld >= p.cout);
  PW_REQUIRE(residual == nullptr || p.res_ld >= p.cout);
  PW_REQUIRE(p.act_channels >= 0 && (p.act_channels & 3) == 0);
  PW_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0);
  PW_REQUIRE(p.kd > 0 && p.kh > 0 && p.kw > 0 && p.sd > 0 && p.sh > 0 && p.sw > 0);
  long long M = (long long)p.n * p.od * p.oh * p.ow;
  PW_REQUIRE(M > 0 && M < (1ll << 31));
  PW_REQUIRE((long long)p.n * p.d * p.h * p.w < (1ll << 31));
  cudaStream_t st = (cudaStream_t)stream;
  if (p.cout <= 32) return launch<256, 32, 8, 4>(p, x, w, scale, bias, residual, y, st);
  // prefer the 128x128 tile when it still fills the machine
  long long tiles128 = (long long)pw_ceil_div(M, 128) * pw_ceil_div(p.cout, 128);
  if (p.cout >= 128 && (p.cout % 128 == 0) && tiles128 >= 296)
    return launch<128, 128, 8, 8>(p, x, w, scale, bias, residual, y, st);
  long long tiles64 = (long long)pw_ceil_div(M, 128) * pw_ceil_div(p.cout, 64);
  if (tiles64 >= 148) return launch<128, 64, 8, 4>(p, x, w, scale, bias, residual, y, st);
  return launch<64, 128, 8, 4>(p, x, w, scale, bias, residual, y, st);
}
