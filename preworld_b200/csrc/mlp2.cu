// Fused two-layer per-row MLP on the tensor cores (tcgen05 kind::tf32, 3xTF32):
//
//     y = act2( W2 . act1( W1 . x + b1 ) + b2 ) [+ residual]
//
// for x [M, 32] rows (voxels), hidden width H <= 256, up to 32 outputs.  The
// hidden tile never leaves the SM: per 128-row tile and per 32-column chunk j of
// the hidden layer
//     ACC1  = A1 . W1_j            (A1 = hi|lo split of the x tile, in TMEM)
//     A2    = split(act1(ACC1 + b1_j))      (workers: tcgen05.ld -> SIMT -> tcgen05.st)
//     ACC2 += A2 . W2_j
// so HBM sees x once and y once (81.9 MB in, 81.9 / 61.4 MB out at 640 000 voxels)
// instead of the 164 / 328 / 3 x 164 MB hidden tensors of the unfused chains.
//
// Replaces, on the reference path:
//   * detectors/preworld_temporal_traj.py:329-341,368 -- fusion_head on
//     cat([voxel, ego]) + residual (the ego half of fusion_head[0] arrives as the
//     per-sample bias b1): SURVEY §8b `pw_fusion_step`;
//   * detectors/preworld.py:81-105,251-254 -- density / semantic / color MLPs as ONE
//     launch (W1 = the three first layers stacked, W2 = block diagonal, final
//     Softplus on the density channels): SURVEY §8b `pw_attr_mlp`.
//
// CTA (persistent, one per SM): warp 0 TMA (weights once, then the x-tile ring),
// warps 1-2 MMA issue for set 0 / 1 (warp 2 owns TMEM), warps 3-10 two sets of four
// worker warps (warp % 4 = TMEM lane quadrant, thread = row).  The two sets work on
// alternate tiles, so one set's activation math runs under the other set's MMAs.
// Weights (pre-split hi / lo, K-major) stay resident in shared memory.
#include <stdio.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "../../include/preworld_b200.h"

namespace {

using namespace pwtc;

constexpr int KC = 32;                        // K chunk: 32 floats = one 128-byte swizzle row
constexpr int ROWB = 128;
constexpr int SETS = 2;
constexpr int TILE_BYTES = 128 * ROWB;        // one x tile: 128 rows x 32 floats
constexpr int FIRST_WORKER = 3;
constexpr int THREADS = (FIRST_WORKER + 4 * SETS) * 32;
constexpr int SET_COLS = 256;                 // TMEM columns per set: A1 | ACC1 | A2 | ACC2
constexpr int COL_A1 = 0, COL_ACC1 = 64, COL_A2 = 128, COL_ACC2 = 192;
constexpr int W1_CHUNK = 2 * KC * ROWB;       // hi rows then lo rows of one hidden chunk
constexpr int STAGE_PER_WARP = 32 * ROWB;

// Softplus(beta = 1, threshold = 20) of the hidden layer: max(x, 0) + ln2 * lg2(1 +
// ex2(-|x| log2e)) -- two MUFU + four FP32 instructions, branch-free (for x > 20 the
// second term is exactly 0, which is torch's threshold branch).  Absolute error
// <= 2e-7 (the lg2 of 1 + t loses t's low bits only where the result is itself below
// 1e-3): the hidden units are combined linearly by W2, so it is the absolute error
// that matters, and it is at the level of the fp32 rounding of the sums.  (The
// accurate log1pf(expf(x)) of common.cuh costs ~100 instructions with its range
// branches; at 128 x H evaluations per tile it set the kernel's time.)
__device__ __forceinline__ float softplus_fast(float x) {
  float t, l;
  const float a = -fabsf(x) * 1.4426950408889634f;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(a));
  const float u = 1.0f + t;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u));
  return fmaf(l, 0.6931471805599453f, fmaxf(x, 0.f));
}

struct Mlp2Params {
  long long M;
  int tiles;
  int H, J;                     // hidden width, J = H / 32 chunks
  int n2, n2p;                  // outputs (multiple of 4), padded to 16 or 32 MMA columns
  int nx;                       // x-tile ring depth
  int act1, act2, act2_channels;
  int out_ld, res_ld;
  float gain1, gain2;           // umma_chain_gain of the two accumulation chains
  const float* b1;
  const float* b2;
  const float* res;
  float* y;
};

__global__ void __launch_bounds__(THREADS, 1)
mlp2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1h,
            const __grid_constant__ CUtensorMap map_w1l,
            const __grid_constant__ CUtensorMap map_w2h,
            const __grid_constant__ CUtensorMap map_w2l, const Mlp2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int w2_chunk = 2 * p.n2p * ROWB;
  uint8_t* xring = smem;
  uint8_t* w1s = xring + (size_t)p.nx * TILE_BYTES;
  uint8_t* w2s = w1s + (size_t)p.J * W1_CHUNK;
  uint8_t* stage = w2s + (size_t)p.J * w2_chunk;
  float* b1s = reinterpret_cast<float*>(stage + 4 * SETS * STAGE_PER_WARP);
  float* b2s = b1s + p.H;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b2s + 32);
  // bars: w_full | x_full[nx] | x_empty[nx] | per set: a1_full acc1_full acc1_empty
  //       a2_full a2_empty acc2_full acc2_empty
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t w_full = bar0, x_full = bar0 + 8, x_empty = x_full + 8 * p.nx;
  const uint32_t set_bars = x_empty + 8 * p.nx;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 1 + 2 * p.nx + 7 * SETS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int s = 0; s < p.nx; ++s) {
      mbar_init(x_full + 8 * s, 1);
      mbar_init(x_empty + 8 * s, 4);
    }
    for (int s = 0; s < SETS; ++s) {
      const uint32_t b = set_bars + 56 * s;
      mbar_init(b + 0, 4);    // a1_full
      mbar_init(b + 8, 1);    // acc1_full   (tcgen05.commit)
      mbar_init(b + 16, 4);   // acc1_empty
      mbar_init(b + 24, 4);   // a2_full
      mbar_init(b + 32, 1);   // a2_empty    (tcgen05.commit)
      mbar_init(b + 40, 1);   // acc2_full   (tcgen05.commit)
      mbar_init(b + 48, 4);   // acc2_empty
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_holder), 512u);
  for (int i = threadIdx.x; i < p.H; i += THREADS) b1s[i] = p.b1 ? __ldg(p.b1 + i) : 0.f;
  if (threadIdx.x < 32) b2s[threadIdx.x] = (p.b2 && threadIdx.x < p.n2) ? __ldg(p.b2 + threadIdx.x) : 0.f;
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_x); prefetch_tensormap(&map_w1h); prefetch_tensormap(&map_w1l);
    prefetch_tensormap(&map_w2h); prefetch_tensormap(&map_w2l);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const int my_tiles = p.tiles > (int)blockIdx.x
                           ? (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    // ===================== TMA: weights once, then the x tiles ====================
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(w_full, (uint32_t)(p.J * (W1_CHUNK + w2_chunk)));
      for (int j = 0; j < p.J; ++j) {
        const uint32_t d1 = smem_u32(w1s + (size_t)j * W1_CHUNK);
        tma_load_2d(d1, &map_w1h, w_full, 0, j * KC);
        tma_load_2d(d1 + KC * ROWB, &map_w1l, w_full, 0, j * KC);
        const uint32_t d2 = smem_u32(w2s + (size_t)j * w2_chunk);
        tma_load_2d(d2, &map_w2h, w_full, j * KC, 0);
        tma_load_2d(d2 + p.n2p * ROWB, &map_w2l, w_full, j * KC, 0);
      }
    }
    __syncwarp();
    int slot = 0;
    uint32_t ph = 1;
    for (int i = 0; i < my_tiles; ++i) {
      mbar_wait(x_empty + 8 * slot, ph);
      if (leader) {
        mbar_expect_tx(x_full + 8 * slot, (uint32_t)TILE_BYTES);
        tma_load_2d(smem_u32(xring + (size_t)slot * TILE_BYTES), &map_x, x_full + 8 * slot, 0,
                    ((int)blockIdx.x + i * (int)gridDim.x) * 128);
      }
      __syncwarp();
      if (++slot == p.nx) { slot = 0; ph ^= 1; }
    }
  } else if (warp <= SETS) {
    // ===================== MMA issue for set `s` ==================================
    const int s = warp - 1;
    const bool leader = elect_one();
    const uint32_t sb = set_bars + 56 * s;
    const uint32_t a1_full = sb, acc1_full = sb + 8, acc1_empty = sb + 16, a2_full = sb + 24,
                   a2_empty = sb + 32, acc2_full = sb + 40, acc2_empty = sb + 48;
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0) + (uint32_t)(s * SET_COLS);
    const uint32_t id1_2n = umma_idesc_tf32_m128(2 * KC), id1_n = umma_idesc_tf32_m128(KC);
    const uint32_t id2_2n = umma_idesc_tf32_m128(2 * p.n2p), id2_n = umma_idesc_tf32_m128(p.n2p);
    const uint64_t w1d = umma_desc_sw128(smem_u32(w1s));
    const uint64_t w2d = umma_desc_sw128(smem_u32(w2s));
    const uint32_t w2_step = (uint32_t)(w2_chunk >> 4);
    mbar_wait(w_full, 0);
    uint32_t m1 = 0, c2 = 0;                     // MMA1 / MMA2 groups issued so far
    auto mma1 = [&](int j) {                     // ACC1 = A1 . [W1h_j ; W1l_j], corr += A1lo . W1h_j
      mbar_wait(acc1_empty, (m1 & 1u) ^ 1u);
      tc_fence_after();
      const uint64_t bd0 = w1d + (uint64_t)((uint32_t)(W1_CHUNK >> 4) * (uint32_t)j);
      if (leader) {
#pragma unroll
        for (int k = 0; k < KC / 8; ++k) {
          const uint64_t bd = bd0 + (uint64_t)(k * 2);
          umma_tf32_ts(tb + COL_ACC1, tb + COL_A1 + k * 8, bd, id1_2n, k > 0 ? 1u : 0u);
          umma_tf32_ts(tb + COL_ACC1 + KC, tb + COL_A1 + KC + k * 8, bd, id1_n, 1u);
        }
        umma_commit(acc1_full);
      }
      __syncwarp();
      ++m1;
    };
    int n = 0;
    for (int i = s; i < my_tiles; i += SETS, ++n) {
      mbar_wait(a1_full, (uint32_t)(n & 1));
      tc_fence_after();
      mma1(0);
      for (int j = 0; j < p.J; ++j) {
        if (j + 1 < p.J) mma1(j + 1);
        mbar_wait(a2_full, c2 & 1u);
        if (j == 0) mbar_wait(acc2_empty, (uint32_t)(n & 1) ^ 1u);
        tc_fence_after();
        const uint64_t bd0 = w2d + (uint64_t)(w2_step * (uint32_t)j);
        if (leader) {
#pragma unroll
          for (int k = 0; k < KC / 8; ++k) {
            const uint64_t bd = bd0 + (uint64_t)(k * 2);
            umma_tf32_ts(tb + COL_ACC2, tb + COL_A2 + k * 8, bd, id2_2n, (j | k) ? 1u : 0u);
            umma_tf32_ts(tb + COL_ACC2 + p.n2p, tb + COL_A2 + KC + k * 8, bd, id2_n, 1u);
          }
          umma_commit(a2_empty);
          if (j == p.J - 1) umma_commit(acc2_full);
        }
        __syncwarp();
        ++c2;
      }
    }
  } else {
    // ===================== workers ================================================
    const int sw = warp - FIRST_WORKER;
    const int s = sw >> 2;
    const int q = warp & 3;                      // TMEM lane quadrant
    const uint32_t lane_field = (uint32_t)(q * 32) << 16;
    const int R = q * 32 + lane;                 // row of the tile
    const uint32_t sb = set_bars + 56 * s;
    const uint32_t a1_full = sb, acc1_full = sb + 8, acc1_empty = sb + 16, a2_full = sb + 24,
                   a2_empty = sb + 32, acc2_full = sb + 40, acc2_empty = sb + 48;
    const uint32_t tb = tmem_base + lane_field + (uint32_t)(s * SET_COLS);
    const uint32_t stg = smem_u32(stage) + (uint32_t)(sw * STAGE_PER_WARP);
    const uint32_t xr0 = smem_u32(xring);
    const bool vec_res = p.res != nullptr;
    uint32_t c1 = 0, c2 = 0;                     // ACC1 read-outs / A2 stores so far
    int n = 0;
    for (int i = s; i < my_tiles; i += SETS, ++n) {
      const int slot = i % p.nx;
      const long long row0 = ((long long)blockIdx.x + (long long)i * gridDim.x) * 128;
      // ---- x row -> hi|lo -> A1 ----------------------------------------------------
      mbar_wait(x_full + 8 * slot, (uint32_t)((i / p.nx) & 1));
      {
        uint32_t hl[64];
        const uint32_t b2a = (xr0 + (uint32_t)(slot * TILE_BYTES + R * ROWB)) ^ (uint32_t)((R & 7) << 4);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = lds128(b2a ^ (uint32_t)(c << 4));
          float h0, h1, h2, h3, l0, l1, l2, l3;
          split2_rn(v.x, v.y, h0, h1, l0, l1);
          split2_rn(v.z, v.w, h2, h3, l2, l3);
          hl[c * 4] = __float_as_uint(h0); hl[c * 4 + 1] = __float_as_uint(h1);
          hl[c * 4 + 2] = __float_as_uint(h2); hl[c * 4 + 3] = __float_as_uint(h3);
          hl[KC + c * 4] = __float_as_uint(l0); hl[KC + c * 4 + 1] = __float_as_uint(l1);
          hl[KC + c * 4 + 2] = __float_as_uint(l2); hl[KC + c * 4 + 3] = __float_as_uint(l3);
        }
        // (A1 of the previous tile of this set is free: its last ACC1 was read out below)
        tmem_st64(tb + COL_A1, hl);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(a1_full); mbar_arrive(x_empty + 8 * slot); }
      }
      // ---- hidden chunks: ACC1 -> +b1, act1 -> hi|lo -> A2 -------------------------
      for (int j = 0; j < p.J; ++j) {
        uint32_t m0[16], m1[16], k0[16], k1[16];
        mbar_wait(acc1_full, c1 & 1u);
        tc_fence_after();
        tmem_ld16_nowait(tb + COL_ACC1, m0);
        tmem_ld16_nowait(tb + COL_ACC1 + 16, m1);
        tmem_ld16_nowait(tb + COL_ACC1 + KC, k0);
        tmem_ld16_nowait(tb + COL_ACC1 + KC + 16, k1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc1_empty);
        ++c1;
        uint32_t hl[64];
        const float* bj = b1s + j * KC;
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
          float u0 = fmaf(__uint_as_float(m0[c]) + __uint_as_float(k0[c]), p.gain1, bj[c]);
          float u1 = fmaf(__uint_as_float(m0[c + 1]) + __uint_as_float(k0[c + 1]), p.gain1, bj[c + 1]);
          float v0 = fmaf(__uint_as_float(m1[c]) + __uint_as_float(k1[c]), p.gain1, bj[16 + c]);
          float v1 = fmaf(__uint_as_float(m1[c + 1]) + __uint_as_float(k1[c + 1]), p.gain1, bj[16 + c + 1]);
          if (p.act1 == PW_ACT_SOFTPLUS) {
            u0 = softplus_fast(u0); u1 = softplus_fast(u1);
            v0 = softplus_fast(v0); v1 = softplus_fast(v1);
          } else if (p.act1 == PW_ACT_RELU) {
            u0 = fmaxf(u0, 0.f); u1 = fmaxf(u1, 0.f); v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f);
          }
          float h0, h1, l0, l1;
          split2_rn(u0, u1, h0, h1, l0, l1);
          hl[c] = __float_as_uint(h0); hl[c + 1] = __float_as_uint(h1);
          hl[KC + c] = __float_as_uint(l0); hl[KC + c + 1] = __float_as_uint(l1);
          split2_rn(v0, v1, h0, h1, l0, l1);
          hl[16 + c] = __float_as_uint(h0); hl[16 + c + 1] = __float_as_uint(h1);
          hl[KC + 16 + c] = __float_as_uint(l0); hl[KC + 16 + c + 1] = __float_as_uint(l1);
        }
        mbar_wait(a2_empty, (c2 & 1u) ^ 1u);       // MMA2 of the previous chunk retired
        tc_fence_after();
        tmem_st64(tb + COL_A2, hl);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a2_full);
        ++c2;
      }
      // ---- ACC2 -> +b2, act2 -> staging (transposed) -> coalesced rows ---------------
      mbar_wait(acc2_full, (uint32_t)(n & 1));
      tc_fence_after();
      {
        float o[32];
        if (p.n2p == 32) {
          uint32_t m0[16], m1[16], k0[16], k1[16];
          tmem_ld16_nowait(tb + COL_ACC2, m0);
          tmem_ld16_nowait(tb + COL_ACC2 + 16, m1);
          tmem_ld16_nowait(tb + COL_ACC2 + 32, k0);
          tmem_ld16_nowait(tb + COL_ACC2 + 48, k1);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            o[c] = __uint_as_float(m0[c]) + __uint_as_float(k0[c]);
            o[16 + c] = __uint_as_float(m1[c]) + __uint_as_float(k1[c]);
          }
        } else {
          uint32_t m0[16], k0[16];
          tmem_ld16_nowait(tb + COL_ACC2, m0);
          tmem_ld16_nowait(tb + COL_ACC2 + 16, k0);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            o[c] = __uint_as_float(m0[c]) + __uint_as_float(k0[c]);
            o[16 + c] = 0.f;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc2_empty);
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          o[c] = fmaf(o[c], p.gain2, b2s[c]);
          if (c < p.act2_channels) o[c] = pw_activate_slow(o[c], p.act2);
        }
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4)
          sts128(stg + (uint32_t)((lane * 8 + (c4 ^ (lane & 7))) << 4),
                 make_float4(o[c4 * 4], o[c4 * 4 + 1], o[c4 * 4 + 2], o[c4 * 4 + 3]));
      }
      __syncwarp();
      {
        const int ch4 = lane & 7;                  // lane -> (row group, 4 channels)
        if (ch4 * 4 < p.n2) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + (lane >> 3);
            const long long row = row0 + q * 32 + r;
            if (row < p.M) {
              float4 v = lds128(stg + (uint32_t)((r * 8 + (ch4 ^ (r & 7))) << 4));
              if (vec_res) {
                const float4 rr = pw_ldg4(p.res + row * p.res_ld + ch4 * 4);
                v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
              }
              *reinterpret_cast<float4*>(p.y + row * p.out_ld + ch4 * 4) = v;
            }
          }
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512u);
}

int mlp2_sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
    n = 148;
  return n;
}

}  // namespace

PW_API int pw_mlp2_supported(int c1, int hidden, int n2) {
  if (c1 != KC || hidden < KC || hidden % KC != 0 || hidden > 256) return 0;
  if (n2 < 4 || n2 > 32 || (n2 & 3) != 0) return 0;
  return encode_tiled_fn() != nullptr ? 1 : 0;
}

PW_API int pw_mlp2(const float* x, int x_ld, long long m, int c1, const float* w1_hi,
                   const float* w1_lo, const float* b1, int hidden, int act1, const float* w2_hi,
                   const float* w2_lo, const float* b2, int n2, int act2, int act2_channels,
                   const float* residual, int res_ld, float* y, int y_ld, void* stream) {
  PW_REQUIRE(x && w1_hi && w1_lo && w2_hi && w2_lo && y);
  PW_REQUIRE(c1 == KC && hidden >= KC && hidden % KC == 0 && hidden <= 256);
  PW_REQUIRE(n2 >= 4 && n2 <= 32 && (n2 & 3) == 0);
  PW_REQUIRE(m > 0 && m < (1ll << 31) - 128);
  PW_REQUIRE(x_ld >= c1 && (x_ld & 3) == 0 && ((uintptr_t)x & 15) == 0);
  PW_REQUIRE(y_ld >= n2 && (y_ld & 3) == 0 && ((uintptr_t)y & 15) == 0);
  PW_REQUIRE(residual == nullptr ||
             (res_ld >= n2 && (res_ld & 3) == 0 && ((uintptr_t)residual & 15) == 0));
  PW_REQUIRE(((uintptr_t)w1_hi & 15) == 0 && ((uintptr_t)w1_lo & 15) == 0 &&
             ((uintptr_t)w2_hi & 15) == 0 && ((uintptr_t)w2_lo & 15) == 0);
  PW_REQUIRE(act2_channels >= 0 && act2_channels <= n2);
  EncodeTiledFn enc = encode_tiled_fn();
  PW_REQUIRE(enc != nullptr);

  Mlp2Params p{};
  p.M = m;
  p.tiles = (int)((m + 127) / 128);
  p.H = hidden; p.J = hidden / KC;
  p.n2 = n2; p.n2p = n2 <= 16 ? 16 : 32;
  p.act1 = act1; p.act2 = act2; p.act2_channels = act2_channels;
  p.out_ld = y_ld; p.res_ld = res_ld;
  p.gain1 = umma_chain_gain(c1);
  p.gain2 = umma_chain_gain(hidden);
  p.b1 = b1; p.b2 = b2; p.res = residual; p.y = y;
  const int w2_chunk = 2 * p.n2p * ROWB;
  const size_t fixed = (size_t)p.J * (W1_CHUNK + w2_chunk) + 4 * SETS * STAGE_PER_WARP +
                       (size_t)(hidden + 32) * 4 + (1 + 2 * 4 + 7 * SETS) * 8 + 16 + 1024;
  p.nx = 4;
  while (p.nx > 2 && fixed + (size_t)p.nx * TILE_BYTES > 227 * 1024) --p.nx;
  const size_t smem = fixed + (size_t)p.nx * TILE_BYTES;
  PW_REQUIRE(smem <= 227 * 1024);

  CUtensorMap mx, m1h, m1l, m2h, m2l;
  auto encode2d = [&](CUtensorMap* mp, const float* base, cuuint64_t inner, cuuint64_t rows,
                      cuuint64_t pitch_floats, cuuint32_t box_rows) -> int {
    cuuint64_t gdim[2] = {inner, rows};
    cuuint64_t gstr[1] = {pitch_floats * 4};
    cuuint32_t box[2] = {(cuuint32_t)KC, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(mp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 1000 + (int)r;
  };
  if (int rc = encode2d(&mx, x, (cuuint64_t)c1, (cuuint64_t)m, (cuuint64_t)x_ld, 128)) return rc;
  if (int rc = encode2d(&m1h, w1_hi, (cuuint64_t)c1, (cuuint64_t)hidden, (cuuint64_t)c1, KC)) return rc;
  if (int rc = encode2d(&m1l, w1_lo, (cuuint64_t)c1, (cuuint64_t)hidden, (cuuint64_t)c1, KC)) return rc;
  if (int rc = encode2d(&m2h, w2_hi, (cuuint64_t)hidden, (cuuint64_t)p.n2p, (cuuint64_t)hidden,
                        (cuuint32_t)p.n2p)) return rc;
  if (int rc = encode2d(&m2l, w2_lo, (cuuint64_t)hidden, (cuuint64_t)p.n2p, (cuuint64_t)hidden,
                        (cuuint32_t)p.n2p)) return rc;

  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mlp2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int grid = p.tiles < mlp2_sm_count() ? p.tiles : mlp2_sm_count();
  mlp2_kernel<<<grid, THREADS, smem, (cudaStream_t)stream>>>(mx, m1h, m1l, m2h, m2l, p);
  PW_LAUNCH_CHECK();
  pw_count_launch(1);
  return 0;
}
