// Halo-resident tcgen05 implicit-GEMM convolution (channels-last fp32 in/out,
// 3xTF32 split for fp32-level accuracy) -- the second-generation tensor-core
// conv of this library.  Differences from conv_umma.cu, each aimed at the
// shared-memory / L2 bandwidth wall that kernel hits on the 3-D encoder
// (Cout 32..64, 27 taps):
//   * the input box of the CTA *including its halo* is loaded ONCE per
//     32-channel chunk (one 5-D TMA, padding = TMA zero fill) instead of once
//     per tap: 27x -> ~2.3x L2->smem traffic for a 3x3x3 conv;
//   * the A operand is fed to the tensor core from TENSOR MEMORY: the split
//     warps read their pixel row of the shifted view straight out of the halo
//     tile, split it into tf32 hi/lo in registers and tcgen05.st it -- no
//     shared-memory write-back, and the MMA reads no A bytes from smem;
//   * hi and lo weights sit back to back in smem, so  Ah*[Bh;Bl]  is ONE MMA
//     of width 2N (main | correction accumulator side by side in TMEM) and
//     Al*Bh a second one: 2 instructions per k-step instead of 3;
//   * a CTA owns up to two 128-pixel M tiles that share every weight stage;
//   * the epilogue transposes through (swizzled) smem so that global stores
//     and residual loads are full 128-byte rows;
//   * CTAs are PERSISTENT (one per SM, tiles b, b + gridDim.x, ...): TMEM,
//     barriers and tensor maps are set up once, every ring (halo slots, weight
//     stages, A slots) runs on across tiles, so the next tile's halo and first
//     weight stages load under the current tile's taps and epilogue; the MMA
//     warp waits on acc_empty (all split warps done reading the accumulators)
//     before it overwrites them.  -7..-22 % per launch against one tile per CTA;
//   * optionally the kw taps along x are folded into the MMA's N dimension
//     (HaloParams::fold, pw_conv_fold_fwd).
//
// Warp roles: 0 halo TMA, 1 weight TMA, 2 MMA issue + TMEM owner, 3..10 two
// sets of four split/epilogue warps (warp%4 = TMEM lane quadrant).  The
// mbarrier protocol between them is model-checked on the CPU by
// tools/halo_protocol_sim.py (tests/test_halo_protocol.py).
//
// Replaces the cuDNN convolutions of the reference path (mmdet ResNet,
// necks/fpn.py, necks/view_transformer.py:473-638, backbones/resnet.py:88-184,
// necks/lss_fpn.py:120-148, detectors/preworld.py:72-105,
// heads/occupancy_head.py:81-177).
#include <stdio.h>
#include <stdlib.h>

#include <mutex>
#include <string>
#include <unordered_map>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "../../include/preworld_b200.h"

namespace {

using namespace pwtc;

constexpr int BLOCK_K = 32;                  // floats = one 128-byte swizzle row
constexpr int ROW_BYTES = 128;
// Split-warp sets (4 warps = 4 TMEM lane quadrants each) per CTA.  Two variants
// are compiled: 2 sets with one CTA per SM (352 threads, up to 512 TMEM columns,
// 227 KB of shared memory) and 1 set with two co-resident CTAs per SM (224
// threads, <= 256 columns and <= 113 KB each), where one CTA's prologue, halo
// load and epilogue run under the other's main loop.  (3 sets x 1 CTA: +3 % on
// the 3x3x3 volume layers, -2 % on the whole step.)  The barrier protocol is
// checked for 1..4 sets by tools/halo_protocol_sim.py.
constexpr int MAX_SPLIT_SETS = 2;
constexpr int FIRST_SPLIT_WARP = 3;
constexpr int halo_threads(int sets) { return (FIRST_SPLIT_WARP + 4 * sets) * 32; }
constexpr int A_SLOT_COLS = 2 * BLOCK_K;     // hi | lo
constexpr int STAGE_BYTES_PER_WARP = 32 * ROW_BYTES;
constexpr int MAX_RING = 8;

// n / d for 0 <= n < 2^31 with a precomputed multiplier (the tile decode runs once per
// tile in three warp roles; four hardware divisions were ~140 instructions)
struct FastDiv {
  uint32_t mul, shift;
  int d;
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = d;
  uint32_t s = 0;
  while ((1u << s) < (uint32_t)d) ++s;
  f.shift = s;
  f.mul = (uint32_t)((((uint64_t)1 << 32) * (((uint64_t)1 << s) - (uint64_t)d)) / (uint64_t)d + 1);
  return f;
}
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) {
  return (int)((__umulhi((uint32_t)n, f.mul) + (uint32_t)n) >> f.shift);
}

struct HaloParams {
  int od, oh, ow, cout;
  int kh, kw, n_taps, chunks;
  int sd, sh, sw, pd, ph, pw, dd, dh, dw;
  int lbx, lby;                // log2 of the M-tile box (bx * by * bz == 128)
  int bx, by, bz;
  int mt, mt_axis;             // M tiles per CTA, stacked along axis 0:x 1:y 2:z
  int cbx, cby, cbz;           // CTA output box
  int hx, hy, hz;              // halo box (input pixels)
  int mt_halo_off;             // halo-row offset of M tile 1
  int halo_stride;             // bytes per halo slot (multiple of 1024)
  int halo_region;             // bytes reserved for the halo ring (>= staging)
  int tiles_x, tiles_y, tiles_z;
  int n_tile;
  int nh, nb;                  // ring depths: halo slots; (weight stage + A slots) entries
  int nacc;                    // accumulator copies per M tile (k-step k -> copy k % nacc):
                               // independent MMA chains hide the accumulate latency at small N
  int tmem_cols;
  // x-tap folding (fold > 0): the kw taps along x become N columns -- the MMA
  // computes P[row, kx*fold_n + n] over INPUT columns and the epilogue adds the
  // shifted partial rows, out[x] = sum_kx P[x + kx*fold_shift, kx, :].  The main
  // loop then runs kh*kd taps with N = fold*fold_n and every A row is split once
  // per (ky, kz) instead of once per tap.  n_tile == fold * fold_n.
  int fold, fold_n, fold_shift;
  int out_bx;                  // valid outputs per x row group (== bx unless folded)
  // Tile loop: CTA b runs tiles b, b + gridDim.x, ... (tile = spatial tile + sp_tiles *
  // N slab).  One tile per CTA unless the plan is persistent: then TMEM, barriers and
  // tensor maps are set up once, the rings run on across tiles (the halo and the first
  // weight stages of the next tile load under the current tile's taps and epilogue),
  // and the epilogue stages through its own shared-memory region (stage_off).
  int total_tiles, sp_tiles, slabs;
  FastDiv fd_sp, fd_tx, fd_ty, fd_tz;   // dividers of the tile decode
  int stage_off;               // byte offset of the epilogue staging (0: aliases the halo ring)
  int stage_bytes;             // bytes reserved between the weight ring and the barriers
  int cps;                     // CTAs per SM the plan counts on (1 or 2)
  int pdl;                     // launched with programmatic stream serialization
  int zero;                    // always 0: added to values read through a scoreboarded load so that
                               // later uses depend on an ALU result (see tmem_base below)
  int dbg;                     // PW_HALO_DBG knock-out bits (timing experiments only)
  long long* ts;               // PW_HALO_TS: per-CTA clock64 milestones (debug)
  int out_ld, res_ld, act, act_channels;
  float gain;                  // umma_chain_gain(): undoes the accumulator's truncation bias
  const float* scale;
  const float* bias;
  const float* res;
  float* y;
};

// Timing instrumentation / knock-out experiments (PW_HALO_TS, PW_HALO_DBG) are
// compiled in only with -DPW_HALO_DEBUG: in the product build they are constant
// false, so the hot loops carry none of their loads and branches.
#ifdef PW_HALO_DEBUG
#define PW_TSON (p.ts != nullptr)
#define PW_DBG(bit) ((p.dbg & (bit)) != 0)
#elif defined(PW_HALO_KO)
// compile-time knock-outs (make ko KO=<bits>): no run-time checks in the loops, so
// the variant runs at the product build's speed minus the removed work
#define PW_TSON false
#define PW_DBG(bit) (((PW_HALO_KO) & (bit)) != 0)
#else
#define PW_TSON false
#define PW_DBG(bit) false
#endif
#define PW_TS(k) do { if (PW_TSON) p.ts[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 + (k)] = clock64(); } while (0)

template <int SPLIT_SETS, int MIN_CTAS>
__global__ void __launch_bounds__(halo_threads(SPLIT_SETS), MIN_CTAS)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a,
                 const __grid_constant__ CUtensorMap map_bh,
                 const __grid_constant__ CUtensorMap map_bl, const HaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int b_stage = 2 * p.n_tile * ROW_BYTES;
  uint8_t* b_ring = smem + p.halo_region;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + (size_t)p.nb * b_stage + p.stage_bytes);
  // bars: halo_full[nh] halo_empty[nh] b_full[nb] empty[nb] a_full[nb*mt] accum acc_empty
  // Ring entry r = (chunk,tap) % nb owns weight stage r and A slots r*mt + m; ONE
  // tcgen05.commit per entry frees both (a commit costs ~400 issue cycles).
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t halo_full = bar0, halo_empty = halo_full + 8 * p.nh;
  const uint32_t b_full = halo_empty + 8 * p.nh, ring_empty = b_full + 8 * p.nb;
  const uint32_t a_full = ring_empty + 8 * p.nb;
  const uint32_t accum_bar = a_full + 8 * p.nb * p.mt;
  const uint32_t acc_empty = accum_bar + 8;      // epilogue done reading the accumulators
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * p.nh + p.nb * (2 + p.mt) + 2);
  // halo-row offset of every tap (the split loop indexes it instead of carrying kx/ky/kz)
  int* tap_tab = reinterpret_cast<int*>(tmem_holder + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) PW_TS(0);


  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nh; ++s) {
      mbar_init(halo_full + 8 * s, 1);
      mbar_init(halo_empty + 8 * s, 4 * SPLIT_SETS);     // one arrive per split warp
    }
    for (int s = 0; s < p.nb; ++s) {
      mbar_init(b_full + 8 * s, 1);
      mbar_init(ring_empty + 8 * s, 1);
    }
    for (int s = 0; s < p.nb * p.mt; ++s) mbar_init(a_full + 8 * s, 4);   // four warps of a set
    mbar_init(accum_bar, 1);
    mbar_init(acc_empty, 4 * SPLIT_SETS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int t = threadIdx.x; t < p.n_taps; t += blockDim.x) {
    const int kx = t % p.kw, ky = (t / p.kw) % p.kh, kz = t / (p.kw * p.kh);
    tap_tab[t] = ((kz * p.dd) * p.hy + ky * p.dh) * p.hx + kx * p.dw;
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_holder), (uint32_t)p.tmem_cols);
  if (warp == 0 && lane == 0) prefetch_tensormap(&map_a);
  if (warp == 1 && lane == 0) { prefetch_tensormap(&map_bh); prefetch_tensormap(&map_bl); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // `+ p.zero`: the sum is an ALU result.  Used directly, every later use of the loaded
  // register statically depends on the load's scoreboard slot, which ptxas re-uses for the
  // epilogue's residual loads -- the TMEM loads then waited for the residual prefetch
  // (17 % of all stall samples on the 64->256 pointwise layer).
  uint32_t tmem_base;
  {
    const uint32_t loaded = *tmem_holder + (uint32_t)p.zero;
    asm volatile("mov.u32 %0, %1;" : "=r"(tmem_base) : "r"(loaded));   // not re-materialised later
  }
  if (threadIdx.x == 0) PW_TS(1);
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, tensor-map
  // prefetch, tap table) may run while the previous kernel of the stream drains; nothing below
  // touches global memory before that kernel has completed and flushed.  Our own dependents
  // may be scheduled as soon as every CTA of this grid is running.
  if (p.pdl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  const int tile_cols = p.nacc * 2 * p.n_tile;          // TMEM columns of one M tile's accumulators
  const uint32_t a_ring_col = (uint32_t)(p.mt * tile_cols);

  // ---- tile coordinates ------------------------------------------------------
  struct Tile { int x0, y0, z0, img, slab; };
  auto decode = [&](int t) {
    Tile tl;
    tl.slab = fdiv(t, p.fd_sp);
    int tt = t - tl.slab * p.sp_tiles;
    int qd = fdiv(tt, p.fd_tx);
    const int tx = tt - qd * p.tiles_x; tt = qd;
    qd = fdiv(tt, p.fd_ty);
    const int ty = tt - qd * p.tiles_y; tt = qd;
    tl.img = fdiv(tt, p.fd_tz);
    const int tz = tt - tl.img * p.tiles_z;
    tl.x0 = tx * p.cbx; tl.y0 = ty * p.cby; tl.z0 = tz * p.cbz;
    return tl;
  };
  const int n_real = p.fold ? p.fold_n : p.n_tile;       // output channels of one N slab
  const int T = p.n_taps;
  const int total_ct = p.chunks * T;            // (chunk, tap) pairs

  // Warps 0..2 run their loops warp-CONVERGED; only the TMA / MMA / commit
  // instruction itself is issued by one elected lane.  (A whole loop under
  // `if (lane == 0)` makes every uniform-register operand a waterfall loop:
  // measured ~65 issue cycles per tcgen05.mma.)
  if (warp == 0) {
    // ===================== halo producer ======================================
    const bool leader = elect_one();
    const uint32_t bytes = (uint32_t)(p.hx * p.hy * p.hz) * ROW_BYTES;
    int s = 0;
    uint32_t ph = 1;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const Tile tl = decode(t);
      const int cx = tl.x0 * p.sw - p.pw, cy = tl.y0 * p.sh - p.ph, cz = tl.z0 * p.sd - p.pd;
      for (int c = 0; c < p.chunks; ++c) {
        mbar_wait(halo_empty + 8 * s, ph);
        if (leader) {
          mbar_expect_tx(halo_full + 8 * s, bytes);
          tma_load_5d(smem_u32(smem + (size_t)s * p.halo_stride), &map_a, halo_full + 8 * s,
                      c * BLOCK_K, cx, cy, cz, tl.img);
        }
        __syncwarp();
        if (++s == p.nh) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== weight producer ====================================
    const bool leader = elect_one();
    const uint32_t b_ring_u32 = smem_u32(b_ring);
    int s = 0;
    uint32_t ph = 1;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int n0w = fdiv(tile, p.fd_sp) * p.n_tile;    // weight rows of this N slab
      for (int c = 0; c < p.chunks; ++c) {
        for (int t = 0; t < T; ++t) {
          mbar_wait(ring_empty + 8 * s, ph);
          if (leader && PW_DBG(64)) mbar_arrive(b_full + 8 * s);   // (64: no weight loads)
          else if (leader) {
            mbar_expect_tx(b_full + 8 * s, (uint32_t)b_stage);
            const uint32_t dst = b_ring_u32 + (uint32_t)(s * b_stage);
            const int k0 = (t * p.chunks + c) * BLOCK_K; // weights are tap-major in K
            tma_load_2d(dst, &map_bh, b_full + 8 * s, k0, n0w);
            tma_load_2d(dst + p.n_tile * ROW_BYTES, &map_bl, b_full + 8 * s, k0, n0w);
          }
          __syncwarp();
          if (++s == p.nb) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== MMA issuer ==========================================
    const bool leader = elect_one();
    const uint32_t idesc_2n = umma_idesc_tf32_m128(2 * p.n_tile);
    const uint32_t idesc_n = umma_idesc_tf32_m128(p.n_tile);
    const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t desc0 = umma_desc_sw128(smem_u32(b_ring));
    const uint32_t desc_stage = (uint32_t)(b_stage >> 4);      // descriptor address units
    const uint32_t acc_stride = (uint32_t)(2 * p.n_tile);
    int r = 0;
    uint32_t ph = 0;
    long long cyc_b = 0, cyc_a = 0, cyc_i = 0, tq = 0;
    int iter = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++iter) {
    if (iter > 0) {                                            // accumulators read out?
      mbar_wait(acc_empty, (uint32_t)((iter - 1) & 1));
      tc_fence_after();
    }
    uint32_t accumulate = 0;                                   // first k-steps overwrite
    for (int ct = 0; ct < total_ct; ++ct) {
      if (PW_TSON) tq = clock64();
      mbar_wait(b_full + 8 * r, ph);
      if (PW_TSON) { cyc_b += clock64() - tq; }
      const uint64_t bdesc = desc0 + (uint64_t)(desc_stage * (uint32_t)r);
      for (int m = 0; m < p.mt; ++m) {
        if (PW_TSON) tq = clock64();
        mbar_wait(a_full + 8 * (r * p.mt + m), ph);
        tc_fence_after();
        if (PW_TSON) { cyc_a += clock64() - tq; tq = clock64(); }
        const uint32_t a_hi = tbase + a_ring_col + (uint32_t)((r * p.mt + m) * A_SLOT_COLS);
        const uint32_t acc = tbase + (uint32_t)(m * tile_cols);
        if (leader && !PW_DBG(1)) {                              // (1: commits only, no MMA)
#pragma unroll
          for (int k = 0; k < BLOCK_K / 8; ++k) {
            const uint64_t bd = bdesc + (uint64_t)(k * 2);     // +32 bytes inside the swizzle row
            const uint32_t acc_k = acc + (uint32_t)(k & (p.nacc - 1)) * acc_stride;
            // [main | corr] (+)= Ah * [Bh ; Bl]
            umma_tf32_ts(acc_k, a_hi + k * 8, bd, idesc_2n,
                         (k < p.nacc) ? accumulate : 1u);
            // corr += Al * Bh
            umma_tf32_ts(acc_k + p.n_tile, a_hi + BLOCK_K + k * 8, bd, idesc_n, 1u);
          }
        }
        __syncwarp();
        if (PW_TSON) cyc_i += clock64() - tq;
      }
      accumulate = 1;
      if (leader) {                                  // weight stage + A slots free on retire
        if (PW_DBG(8)) mbar_arrive(ring_empty + 8 * r);      // (8, with 1: plain arrive, no commit)
        else umma_commit(ring_empty + 8 * r);
      }
      __syncwarp();
      if (++r == p.nb) { r = 0; ph ^= 1; }
    }
    if (leader) umma_commit(accum_bar);
    __syncwarp();
    }
    if (PW_TSON && leader) {
      long long* t = p.ts + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16;
      t[6] = clock64(); t[12] = t[0] + cyc_b; t[13] = t[0] + cyc_a; t[14] = t[0] + cyc_i;
    }
  } else {
    // ===================== split warps, then epilogue ==========================
    const int sw_id = warp - FIRST_SPLIT_WARP;
    const int set = sw_id >> 2;
    const int q = warp & 3;                              // TMEM lane quadrant of this warp
    const uint32_t lane_field = (uint32_t)(q * 32) << 16;
    const int R = q * 32 + lane;                         // row of the M tile
    const int rx = R & (p.bx - 1);
    const int ry = (R >> p.lbx) & (p.by - 1);
    const int rz = R >> (p.lbx + p.lby);
    const int row_base = ((rz * p.sd) * p.hy + ry * p.sh) * p.hx + rx * p.sw;
    const uint32_t a_ring = tmem_base + lane_field + a_ring_col;
    const uint32_t smem_base_u32 = smem_u32(smem);

    // This set builds the A tiles of M tile `m_set` for taps tap0, tap0+tstep, ...
    // of every chunk (it = (c*T + t)*mt + m; sets take it = set, set+SETS, ...).  All state advances incrementally -- the loop body
    // is the critical instruction stream of the kernel (8 warps on 4 schedulers).
    int m_cur = set & (p.mt - 1);                        // M tile of the current iteration (mt is 1 or 2)
    const int lmt = p.mt - 1;                            // log2(mt)
    int c = 0, t = 0, r = 0, hs = 0;                     // chunk, tap, ring entry, halo slot
    uint32_t eph = 1u, hph = 0u;                         // parities: ring_empty, halo_full
    const uint32_t tap_tab_u32 = smem_u32(tap_tab);
    auto advance_taps = [&](int n) {                     // n <= nb (enforced by the planner)
      r += n;
      const int wrap = r >= p.nb ? 1 : 0;                // (selects: a branch here cost ~6 % of the
      r -= wrap ? p.nb : 0;                              //  split warps' samples in branch resolution)
      eph ^= (uint32_t)wrap;
      t += n;
      while (__builtin_expect(t >= T, 0)) {              // next chunk (T may be 1: twice)
        t -= T;
        ++c;
        if (++hs == p.nh) { hs = 0; hph ^= 1u; }
      }
    };
    float4 raw[8];
    auto load_row = [&]() {
      int toff;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(toff) : "r"(tap_tab_u32 + (uint32_t)(t << 2)));
      const int hrow = row_base + m_cur * p.mt_halo_off + toff;
      // 16-byte chunk j of row hrow sits at ((j ^ (hrow & 7)) << 4) (TMA 128B swizzle)
      const uint32_t b2 = (smem_base_u32 + (uint32_t)(hs * p.halo_stride + hrow * ROW_BYTES)) ^
                          (uint32_t)((hrow & 7) << 4);
#pragma unroll
      for (int j = 0; j < 8; ++j) raw[j] = lds128(b2 ^ (uint32_t)(j << 4));
    };
    // Every split warp arrives once per chunk on halo_empty.  A chunk this warp
    // never read (T*mt < SPLIT_SETS: the set skips it) may only be released once
    // its load has been ISSUED -- i.e. after the previous occupant of the slot was
    // released by everybody -- or the arrival would land in the previous phase
    // and free the slot under a slower set: wait for its halo_full first.
    int gc0 = 0;                                         // chunks of the tiles before this one
    auto release_chunk = [&](int ch) {
      const int g = gc0 + ch;                            // the halo ring runs on across tiles
      const int slot = g % p.nh;
      mbar_wait(halo_full + 8 * slot, (uint32_t)((g / p.nh) & 1));
      __syncwarp();
      if (lane == 0) mbar_arrive(halo_empty + 8 * slot);
    };
    advance_taps(set >> lmt);                            // first tap of this set
    // The (chunk, tap, M tile) sequence of a persistent CTA simply runs on across its
    // tiles: a set whose stride overshoots a tile's end starts the next one mid-way.
    int iter = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++iter) {
    const Tile tl = decode(tile);
    const int x0 = tl.x0, y0 = tl.y0, z0 = tl.z0, img = tl.img;
    const int n0 = tl.slab * n_real;
    int released = 0;                                    // halo chunks this warp has released
    int pending = -1;                                    // A slot stored but not yet published
    long long sp_wait = 0, sp_busy = 0, sp_t0 = 0;       // PW_HALO_TS: cycles waiting / storing
#pragma unroll 1
    while (released < min(c, p.chunks)) {                // chunks before this set's first one
      release_chunk(released);
      ++released;
    }
    if (c < p.chunks) {
      mbar_wait(halo_full + 8 * hs, hph);
      if (sw_id == 0 && lane == 0) PW_TS(3);
      load_row();
    }
    while (c < p.chunks) {
      const int slot = r * p.mt + m_cur;
      const uint32_t a_col = a_ring + (uint32_t)(slot * A_SLOT_COLS);
      const int ring_r = r;
      const uint32_t ring_ph = eph;
      // hi (columns 0..31) and lo (32..63) of this row, stored by ONE tcgen05.st
      uint32_t hl[64];
      if (!PW_DBG(2))                                      // (2: no hi/lo split)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 v = raw[j];
        float h0, h1, h2, h3, l0, l1, l2, l3;
        // round-to-nearest split, two elements per packed instruction (tc_ptx.cuh)
        split2_rn(v.x, v.y, h0, h1, l0, l1);
        split2_rn(v.z, v.w, h2, h3, l2, l3);
        hl[j * 4] = __float_as_uint(h0); hl[j * 4 + 1] = __float_as_uint(h1);
        hl[j * 4 + 2] = __float_as_uint(h2); hl[j * 4 + 3] = __float_as_uint(h3);
        hl[BLOCK_K + j * 4] = __float_as_uint(l0); hl[BLOCK_K + j * 4 + 1] = __float_as_uint(l1);
        hl[BLOCK_K + j * 4 + 2] = __float_as_uint(l2); hl[BLOCK_K + j * 4 + 3] = __float_as_uint(l3);
      }
      // raw[] is consumed: advance to this set's next row (it += SPLIT_SETS in (tap, m)
      // space) and, inside the same chunk (the common case), fetch it NOW -- the
      // shared-memory latency then runs under the barrier traffic below instead of
      // heading the next iteration's dependency chain
      const int c_prev = c;
      {
        const int mm = m_cur + SPLIT_SETS;
        m_cur = mm & (p.mt - 1);
        advance_taps(mm >> lmt);
      }
      const bool same_chunk = c == c_prev;
      if (same_chunk && !PW_DBG(4)) load_row();            // (4: no halo row loads)
      // the PREVIOUS row's store is published only now: its completion latency
      // ran under this row's hi/lo split
      if (pending >= 0) {
        if (!PW_DBG(16)) tmem_st_wait();                   // (16: no TMEM store, no wait for it)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full + 8 * pending);
        if (PW_TSON) sp_busy += clock64() - sp_t0;
      }
      {
        long long tw = 0;
        if (PW_TSON) tw = clock64();
        mbar_wait(ring_empty + 8 * ring_r, ring_ph);     // MMAs reading this slot retired
        tc_fence_after();
        if (PW_TSON) { sp_wait += clock64() - tw; sp_t0 = clock64(); }
      }
      if (PW_DBG(1024)) tmem_st32(a_col, hl);            // (timing experiment: half the bytes)
      else if (!PW_DBG(2048) && !PW_DBG(16)) tmem_st64(a_col, hl);    // (2048: no store at all)
      pending = slot;
      if (!PW_DBG(4096)) {
        // Published at once: the main loop is bound by the ring round trip (MMA retire ->
        // commit -> store -> publish -> MMA) with 3 entries in flight, not by the split
        // warps' issue rate -- deferring the publication by one row to hide the store's
        // completion latency (round 1; bit 4096 restores it) lengthened that loop: +2 %.
        if (!PW_DBG(16)) tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full + 8 * pending);
        pending = -1;
      }
      if (!same_chunk) {                                 // chunk boundary (rare)
        const int upto = c < p.chunks ? c : p.chunks;
#pragma unroll 1
        while (released < upto) {                        // chunks this warp is done reading
          release_chunk(released);
          ++released;
        }
        if (c < p.chunks) {
          mbar_wait(halo_full + 8 * hs, hph);
          if (!PW_DBG(4)) load_row();
        }
      }
    }
    if (pending >= 0) {
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full + 8 * pending);
    }
#pragma unroll 1
    while (released < p.chunks) {                        // a set with no work in the last chunks
      release_chunk(released);
      ++released;
    }

    // ---- epilogue: TMEM -> regs (main + corr) -> swizzled smem -> coalesced
    //      affine / residual / activation / store (kept small: it is straight-line
    //      code after the main loop and must not thrash the instruction cache) ----
    if (sw_id == 0 && lane == 0) PW_TS(11);
    if (PW_TSON && sw_id == 0 && lane == 0) {
      long long* t = p.ts + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16;
      t[15] = t[0] + sp_wait; t[4] = t[0] + sp_busy;
    }
    const uint32_t stage = smem_base_u32 + (uint32_t)(p.stage_off + sw_id * STAGE_BYTES_PER_WARP);
    // an epilogue item = (M tile, group of cgw channels); the sets take items in turn.
    // Folded launches use 16-channel groups so that both sets share the one M tile.
    const int cgw = p.fold ? 16 : 32;
    const int ncg = (n_real + cgw - 1) / cgw;
    const int items = p.mt * ncg;
    const int act_end = p.act_channels > 0 ? p.act_channels : p.cout;
    const bool vec_ok = ((p.out_ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0) &&
                        (p.res == nullptr || ((p.res_ld & 3) == 0 &&
                                              (reinterpret_cast<uintptr_t>(p.res) & 15) == 0));
    const int ch4 = lane & 7;
    // pixel index (within the output tensor) of the 8 rows this lane stores, for
    // M tile 0; tile m adds m * mt_pix.  -1 = outside the output.
    int rowpix[2][8];
#pragma unroll
    for (int mm = 0; mm < 2; ++mm) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int Rr = q * 32 + i * 4 + (lane >> 3);
        int ox = x0 + (Rr & (p.bx - 1));
        int oy = y0 + ((Rr >> p.lbx) & (p.by - 1));
        int oz = z0 + (Rr >> (p.lbx + p.lby));
        if (p.mt_axis == 0) ox += mm * p.bx;
        else if (p.mt_axis == 1) oy += mm * p.by;
        else oz += mm * p.bz;
        const bool ok = mm < p.mt && (Rr & (p.bx - 1)) < p.out_bx && ox < p.ow && oy < p.oh &&
                        oz < p.od;
        rowpix[mm][i] = ok ? ((img * p.od + oz) * p.oh + oy) * p.ow + ox : -1;
      }
    }
    // residual rows of an item are fetched ahead of use: for the first item while
    // the last MMAs are still running, for the others under the TMEM loads
    float4 rr[8];
    auto prefetch_res = [&](int item) {
      const int m = item / ncg;
      const int cb = n0 + (item - m * ncg) * cgw + ch4 * 4;
      const bool on = p.res != nullptr && vec_ok && cb + 4 <= p.cout && ch4 * 4 < cgw &&
                      (item - m * ncg) * cgw + ch4 * 4 < n_real && !PW_DBG(256);   // (256: no residual loads)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int pixi = m == 0 ? rowpix[0][i] : rowpix[1][i];
        rr[i] = (on && pixi >= 0) ? pw_ldg4(p.res + (size_t)pixi * p.res_ld + cb)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    if (set < items) prefetch_res(set);
    if (sw_id == 0 && lane == 0) PW_TS(2);
    mbar_wait(accum_bar, (uint32_t)(iter & 1));
    tc_fence_after();
    if (sw_id == 0 && lane == 0) PW_TS(7);
    for (int item = set; item < items && !PW_DBG(32); item += SPLIT_SETS) {   // (32: no epilogue)
      const int m = item / ncg;
      const int col0 = (item - m * ncg) * cgw;
      const int ncol = min(cgw, n_real - col0);          // 16 or 32 (warp-uniform)
      if (item != set) prefetch_res(item);
      // phase 1: this thread's accumulator row (32 channels) -> its staging row
      if (!p.fold && p.nacc == 1 && ncol == 32) {
        // common case: all four TMEM loads of the item (main | corr, two halves) in
        // flight under ONE wait -- a tcgen05.ld round trip is ~600 cycles here and the
        // epilogue is half of a pointwise layer's tile time.  (Issuing the NEXT item's
        // loads before this item's store phase was tried: the 64 loop-carried
        // registers spill into the split loop, 2x slower.)
        uint32_t a0[16], b0[16], a1[16], b1[16];
        const uint32_t taddr = tmem_base + lane_field + (uint32_t)(m * tile_cols + col0);
        tmem_ld16_nowait(taddr, a0);
        tmem_ld16_nowait(taddr + p.n_tile, b0);
        tmem_ld16_nowait(taddr + 16, a1);
        tmem_ld16_nowait(taddr + 16 + p.n_tile, b1);
        tmem_ld_wait();
        if (item == set && sw_id == 0 && lane == 0) PW_TS(5);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          sts128(stage + (uint32_t)((lane * 8 + (j4 ^ (lane & 7))) << 4),
                 make_float4(__uint_as_float(a0[j4 * 4]) + __uint_as_float(b0[j4 * 4]),
                             __uint_as_float(a0[j4 * 4 + 1]) + __uint_as_float(b0[j4 * 4 + 1]),
                             __uint_as_float(a0[j4 * 4 + 2]) + __uint_as_float(b0[j4 * 4 + 2]),
                             __uint_as_float(a0[j4 * 4 + 3]) + __uint_as_float(b0[j4 * 4 + 3])));
          sts128(stage + (uint32_t)((lane * 8 + ((4 + j4) ^ (lane & 7))) << 4),
                 make_float4(__uint_as_float(a1[j4 * 4]) + __uint_as_float(b1[j4 * 4]),
                             __uint_as_float(a1[j4 * 4 + 1]) + __uint_as_float(b1[j4 * 4 + 1]),
                             __uint_as_float(a1[j4 * 4 + 2]) + __uint_as_float(b1[j4 * 4 + 2]),
                             __uint_as_float(a1[j4 * 4 + 3]) + __uint_as_float(b1[j4 * 4 + 3])));
        }
      } else
      for (int half = 0; half * 16 < ncol; ++half) {
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
        if (p.fold) {
          // out[row] = sum_kx P_kx[row + kx * shift]: the partial rows of the other
          // x taps sit `shift` lanes further down in this warp (bx <= 32)
          for (int kx = 0; kx < p.fold; ++kx) {
            uint32_t a[16], b[16];
            const uint32_t taddr = tmem_base + lane_field +
                                   (uint32_t)(m * tile_cols + kx * p.fold_n + col0 + half * 16);
            tmem_ld16_nowait(taddr, a);
            tmem_ld16_nowait(taddr + p.n_tile, b);
            tmem_ld_wait();
            const int sh = kx * p.fold_shift;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float v = __uint_as_float(a[j]) + __uint_as_float(b[j]);
              acc[j] += __shfl_down_sync(0xffffffffu, v, sh);
            }
          }
        } else
        for (int copy = 0; copy < p.nacc; ++copy) {
          uint32_t a[16], b[16];
          const uint32_t taddr = tmem_base + lane_field +
                                 (uint32_t)(m * tile_cols + copy * 2 * p.n_tile + col0 + half * 16);
          tmem_ld16_nowait(taddr, a);
          tmem_ld16_nowait(taddr + p.n_tile, b);
          tmem_ld_wait();
          if (item == set && half == 0 && sw_id == 0 && lane == 0) PW_TS(5);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += __uint_as_float(a[j]) + __uint_as_float(b[j]);
        }
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4)
          sts128(stage + (uint32_t)((lane * 8 + ((half * 4 + j4) ^ (lane & 7))) << 4),
                 make_float4(acc[j4 * 4], acc[j4 * 4 + 1], acc[j4 * 4 + 2], acc[j4 * 4 + 3]));
      }
      __syncwarp();
      if (item == set && sw_id == 0 && lane == 0) PW_TS(10);
      // phase 2: lane -> (row group, 4 channels); 8 lanes cover one 128-byte row
      const int cbase = n0 + col0 + ch4 * 4;
      if (ch4 * 4 < ncol && cbase < p.cout) {
        float sc[4], bi[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool ok = cbase + e < p.cout;
          sc[e] = ((p.scale && ok) ? __ldg(p.scale + cbase + e) : 1.f) * p.gain;
          bi[e] = (p.bias && ok) ? __ldg(p.bias + cbase + e) : 0.f;
        }
        const int a_ = (cbase < act_end) ? p.act : PW_ACT_NONE;     // act_end % 4 == 0
        const bool vec = vec_ok && cbase + 4 <= p.cout;
        const float* yb = p.y + cbase;
        if (vec && (a_ == PW_ACT_NONE || a_ == PW_ACT_RELU)) {
          // fast path (every hot layer): branch-free, fully unrolled, residual
          // already in registers
          const bool relu = a_ == PW_ACT_RELU;
          const int ldo = p.out_ld;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + (lane >> 3);
            const int pixi = m == 0 ? rowpix[0][i] : rowpix[1][i];
            const float4 v4 = lds128(stage + (uint32_t)((r * 8 + (ch4 ^ (r & 7))) << 4));
            float4 o;
            o.x = fmaf(v4.x, sc[0], bi[0]) + rr[i].x;
            o.y = fmaf(v4.y, sc[1], bi[1]) + rr[i].y;
            o.z = fmaf(v4.z, sc[2], bi[2]) + rr[i].z;
            o.w = fmaf(v4.w, sc[3], bi[3]) + rr[i].w;
            o.x = relu ? fmaxf(o.x, 0.f) : o.x;
            o.y = relu ? fmaxf(o.y, 0.f) : o.y;
            o.z = relu ? fmaxf(o.z, 0.f) : o.z;
            o.w = relu ? fmaxf(o.w, 0.f) : o.w;
            if (pixi >= 0 && !PW_DBG(128))                     // (128: no output stores)
              *reinterpret_cast<float4*>(const_cast<float*>(yb) + (size_t)pixi * ldo) = o;
          }
        } else if (vec && a_ == PW_ACT_GELU) {
          // Swin FFN (fc1 + GELU, 4C wide): the fast path's shape, erff inline.  (In the
          // opt-in <1, 2> variant this branch costs 140 bytes of extra spill traffic.)
          const int ldo = p.out_ld;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + (lane >> 3);
            const int pixi = m == 0 ? rowpix[0][i] : rowpix[1][i];
            const float4 v4 = lds128(stage + (uint32_t)((r * 8 + (ch4 ^ (r & 7))) << 4));
            float4 o;
            o.x = pw_activate(fmaf(v4.x, sc[0], bi[0]) + rr[i].x, PW_ACT_GELU);
            o.y = pw_activate(fmaf(v4.y, sc[1], bi[1]) + rr[i].y, PW_ACT_GELU);
            o.z = pw_activate(fmaf(v4.z, sc[2], bi[2]) + rr[i].z, PW_ACT_GELU);
            o.w = pw_activate(fmaf(v4.w, sc[3], bi[3]) + rr[i].w, PW_ACT_GELU);
            if (pixi >= 0)
              *reinterpret_cast<float4*>(const_cast<float*>(yb) + (size_t)pixi * ldo) = o;
          }
        } else {
#pragma unroll 1
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + (lane >> 3);
            const int pixi = m == 0 ? rowpix[0][i] : rowpix[1][i];
            if (pixi < 0) continue;
            const float4 v4 = lds128(stage + (uint32_t)((r * 8 + (ch4 ^ (r & 7))) << 4));
            float va = fmaf(v4.x, sc[0], bi[0]), vb = fmaf(v4.y, sc[1], bi[1]);
            float vc = fmaf(v4.z, sc[2], bi[2]), vd = fmaf(v4.w, sc[3], bi[3]);
            float* yrow = p.y + (size_t)pixi * p.out_ld + cbase;
            const float* rrow = p.res ? p.res + (size_t)pixi * p.res_ld + cbase : nullptr;
            if (rrow) {
              va += __ldg(rrow);
              if (cbase + 1 < p.cout) vb += __ldg(rrow + 1);
              if (cbase + 2 < p.cout) vc += __ldg(rrow + 2);
              if (cbase + 3 < p.cout) vd += __ldg(rrow + 3);
            }
            va = pw_activate_slow(va, a_); vb = pw_activate_slow(vb, a_);
            vc = pw_activate_slow(vc, a_); vd = pw_activate_slow(vd, a_);
            if (vec) {
              *reinterpret_cast<float4*>(yrow) = make_float4(va, vb, vc, vd);
              continue;
            }
            yrow[0] = va;
            if (cbase + 1 < p.cout) yrow[1] = vb;
            if (cbase + 2 < p.cout) yrow[2] = vc;
            if (cbase + 3 < p.cout) yrow[3] = vd;
          }
        }
      }
      __syncwarp();
    }
    // accumulators of this tile are read out: the MMA warp may start the next tile
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(acc_empty);
    c -= p.chunks;                                       // position within the next tile
    gc0 += p.chunks;
    }
  }

  if (warp == FIRST_SPLIT_WARP && lane == 0) PW_TS(8);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) PW_TS(9);
  if (warp == 2) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ---- host side -----------------------------------------------------------------
inline int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }
inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int SMEM_LIMIT_2CTA = (228 * 1024 - 2 * 1024) / 2;   // two CTAs + 1 KB reserved each
constexpr int MAX_TAPS = 64;                 // tap table entries (7x7 stem, 3x3x3 volume convs)
constexpr int SMEM_SLACK = 1024 /*align*/ + (6 * MAX_RING + 1) * 8 + 16 +   // barriers: 2nh + nb(2+mt) + 2
                           MAX_TAPS * 4;                                   // + tap table

struct HaloPlan {
  bool ok = false;
  pw_conv_desc c;              // possibly flattened (1x1 convs)
  HaloParams p;
  size_t smem = 0;
  dim3 grid;
  int sets = 2;                // kernel variant: split sets per CTA (2: 1 CTA/SM, 1: 2 CTAs/SM)
  double cost = -1;            // cycle-model estimate of the whole launch
};

// Tuning knobs for experiments (tools/umma_probe.py), read once per process.
struct HaloKnobs {
  int nt = 0, nacc = 0, nb = 0, mt = 0, dbg = 0, sets = 0;
  int fold = -1;               // PW_HALO_FOLD=0: pw_conv_fold_supported() always says no
  int persist = 1;             // PW_HALO_PERSIST=0: one tile per CTA (the pre-persistent launch)
  bool ts = false;
  HaloKnobs() {
    auto geti = [](const char* k) { const char* e = getenv(k); return e ? atoi(e) : 0; };
    nt = geti("PW_HALO_NT"); nacc = geti("PW_HALO_NACC"); nb = geti("PW_HALO_NB");
    mt = geti("PW_HALO_MT"); dbg = geti("PW_HALO_DBG"); sets = geti("PW_HALO_SETS");
    ts = getenv("PW_HALO_TS") != nullptr;
    if (getenv("PW_HALO_FOLD")) fold = geti("PW_HALO_FOLD");
    if (getenv("PW_HALO_PERSIST")) persist = geti("PW_HALO_PERSIST");
  }
};
const HaloKnobs& knobs() {
  static const HaloKnobs k;
  return k;
}

// Picks box / M-tile count / ring depths with a small cycle model:
// waves * (per-CTA mainloop + halo load).
HaloPlan make_plan_uncached(const pw_conv_desc& in);

// The plan depends only on the descriptor: cache it (a forward re-issues the
// same ~100 layer shapes every step; the search below costs ~50 us of host time).
const HaloPlan& make_plan(const pw_conv_desc& in) {
  static std::mutex mu;
  static std::unordered_map<std::string, HaloPlan> cache;
  pw_conv_desc key = in;
  key.act = 0; key.act_channels = 0; key.out_ld = 0; key.res_ld = 0; key.w_ld = 0;
  std::string k(reinterpret_cast<const char*>(&key), sizeof(key));
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(k);
  if (it == cache.end()) {
    it = cache.emplace(k, make_plan_uncached(key)).first;
    static const bool show = getenv("PW_HALO_PLAN") != nullptr;
    if (show && it->second.ok) {
      const HaloPlan& pl = it->second;
      const HaloParams& p = pl.p;
      fprintf(stderr, "[halo plan] %dx%dx%dx%d cin %d cout %d k %dx%dx%d: grid %u mt %d(axis %d) n_tile %d nh %d nb %d "
              "box %dx%dx%d halo %dx%dx%d (%d B) smem %zu tmem %d sets %d tiles %d\n",
              in.n, in.d, in.h, in.w, in.cin, in.cout, in.kd, in.kh, in.kw, pl.grid.x, p.mt, p.mt_axis,
              p.n_tile, p.nh, p.nb, p.cbx, p.cby, p.cbz, p.hx, p.hy, p.hz, p.halo_stride, pl.smem,
              p.tmem_cols, pl.sets, p.total_tiles);
    }
  }
  return it->second;
}

// Best plan for one kernel variant: `sets` split sets per CTA, the CTA limited to
// `tmem_limit` TMEM columns and `smem_limit` bytes, `cps` CTAs resident per SM.
HaloPlan search_plan(const pw_conv_desc& in, int sets, int tmem_limit, int smem_limit, int cps,
                     int fold_n, bool persist);

// Persistent when that plan fits without giving up ring depth: the persistent CTA
// needs a second halo slot even for a one-chunk conv plus its own 32 KB epilogue
// staging; where that costs a weight stage or a halo slot (the stride-2 layers, whose
// halo is 2x per axis) the one-tile-per-CTA plan stays the faster one (measured).
HaloPlan pick_plan(const pw_conv_desc& in, int sets, int tmem_limit, int smem_limit, int cps,
                   int fold_n) {
  HaloPlan one_tile = search_plan(in, sets, tmem_limit, smem_limit, cps, fold_n, false);
  if (knobs().persist != 1) return one_tile;
  HaloPlan loop = search_plan(in, sets, tmem_limit, smem_limit, cps, fold_n, true);
  if (!loop.ok) return one_tile;
  if (!one_tile.ok) return loop;
  if (loop.p.nb < one_tile.p.nb || loop.p.nh < one_tile.p.nh) return one_tile;
  return loop;
}

// Plan of the x-tap-folded variant (pw_conv_fold_fwd), cached like make_plan.
const HaloPlan& make_fold_plan(const pw_conv_desc& in, int fold_n) {
  static std::mutex mu;
  static std::unordered_map<std::string, HaloPlan> cache;
  pw_conv_desc key = in;
  key.act = 0; key.act_channels = 0; key.out_ld = 0; key.res_ld = 0; key.w_ld = 0;
  std::string k(reinterpret_cast<const char*>(&key), sizeof(key));
  k.push_back((char)fold_n);
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(k);
  if (it == cache.end())
    it = cache.emplace(k, pick_plan(key, 2, 512, SMEM_LIMIT, 1, fold_n)).first;
  return it->second;
}

HaloPlan make_plan_uncached(const pw_conv_desc& in) {
  HaloPlan one_cta = pick_plan(in, 2, 512, SMEM_LIMIT, 1, 0);
  // The 1-set / 2-CTAs-per-SM variant is opt-in (PW_HALO_SETS=1): measured on the
  // path's layers it wins 7-9 % on two shapes (32->64 k333, 224->32 k111) and loses
  // up to 75 % where the 256-column TMEM share forces one M tile per CTA.
  if (knobs().sets != 1) return one_cta;
  HaloPlan two_ctas = pick_plan(in, 1, 256, SMEM_LIMIT_2CTA, 2, 0);
  return two_ctas.ok ? two_ctas : one_cta;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

HaloPlan search_plan(const pw_conv_desc& in, const int SPLIT_SETS, const int tmem_limit,
                     const int smem_limit, const int cps, const int fold_n, const bool persist) {
  HaloPlan plan;
  plan.sets = SPLIT_SETS;
  pw_conv_desc c = in;
  // x-tap folding: stride 1 along x, 2..4 taps, N = kw * fold_n <= 128
  const int fold = fold_n > 0 ? c.kw : 0;
  if (fold_n > 0 && (c.sw != 1 || c.kw < 2 || c.kw > 4 || c.kw * fold_n > 128 ||
                     (fold_n != 16 && fold_n != 32)))
    return plan;
  if (c.cin % BLOCK_K != 0 || c.in_ld % 4 != 0 || c.cout < 1) return plan;
  if (c.n < 1 || c.od < 1 || c.oh < 1 || c.ow < 1) return plan;
  const bool pointwise = c.kd == 1 && c.kh == 1 && c.kw == 1 && c.sd == 1 && c.sh == 1 &&
                         c.sw == 1 && c.pd == 0 && c.ph == 0 && c.pw == 0;
  if (pointwise) {
    // a 1x1 conv is a plain [pixels, cin] x [cin, cout] GEMM: flatten to one row
    const long long px = (long long)c.n * c.d * c.h * c.w;
    if (px >= (1ll << 31)) return plan;
    c.w = c.ow = (int)px;
    c.n = c.d = c.h = c.od = c.oh = 1;
  }
  const int chunks = c.cin / BLOCK_K;
  const int T = c.kd * c.kh * (fold ? 1 : c.kw);         // taps the main loop iterates
  if (T > MAX_TAPS) return plan;
  static const int boxes3[][3] = {{8, 4, 4}, {8, 8, 2}, {16, 4, 2}, {16, 8, 1}, {8, 16, 1},
                                  {32, 4, 1}, {16, 2, 4}, {32, 2, 2}, {8, 2, 8}};
  static const int boxes2[][3] = {{16, 8, 1}, {8, 16, 1}, {32, 4, 1}, {64, 2, 1}, {128, 1, 1}};
  static const int boxes1[][3] = {{128, 1, 1}};
  const int (*boxes)[3] = pointwise ? boxes1 : (c.od > 1 ? boxes3 : boxes2);
  const int nboxes = pointwise ? 1 : (c.od > 1 ? 9 : 5);
  const int n_full = fold ? fold * fold_n : min(round_up(c.cout, 16), 128);

  double best = -1;
  for (int ntry = 0; ntry < 3; ++ntry) {
    const int n_tile = ntry == 0 ? n_full : (ntry == 1 ? 64 : 32);
    if (ntry > 0 && (fold || n_tile >= n_full)) continue;
    if (knobs().nt && knobs().nt != n_tile && knobs().nt < n_full) continue;
    const int slabs = fold ? pw_ceil_div(c.cout, fold_n) : pw_ceil_div(c.cout, n_tile);
    const int b_stage = 2 * n_tile * ROW_BYTES;
    for (int bi = 0; bi < nboxes; ++bi) {
      for (int mt = 1; mt <= 2; ++mt) {
        for (int axis = 0; axis < (mt == 1 ? 1 : 3); ++axis) {
          int b[3] = {boxes[bi][0], boxes[bi][1], boxes[bi][2]};
          int cb[3] = {b[0], b[1], b[2]};
          if (fold) {
            // an x row group (bx input columns) must sit inside one warp and keep
            // at least half of its rows as outputs
            if (mt != 1 || b[0] > 32) continue;
            cb[0] = b[0] - (c.kw - 1) * c.dw;
            if (cb[0] * 2 < b[0]) continue;
          }
          if (mt == 2) cb[axis] *= 2;
          if (pointwise && axis != 0) continue;
          const int ext[3] = {c.ow, c.oh, c.od};
          if (mt == 2 && ext[axis] <= b[axis]) continue;       // second tile would be empty
          const int st[3] = {c.sw, c.sh, c.sd}, dl[3] = {c.dw, c.dh, c.dd};
          const int kk[3] = {c.kw, c.kh, c.kd};
          int h[3];
          bool fit = true;
          for (int a = 0; a < 3; ++a) {
            h[a] = (cb[a] - 1) * st[a] + (kk[a] - 1) * dl[a] + 1;
            if (h[a] > 256) fit = false;
          }
          if (!fit) continue;
          const int hrows = h[0] * h[1] * h[2];
          const int halo_stride = round_up(hrows * ROW_BYTES, 1024);
          // TMEM: accumulators (mt * nacc * 2n) + A ring (nb * mt * 64 columns)
          int nacc = 1;
          if (knobs().nacc == 2 && !fold) nacc = 2;
          if (mt * nacc * 2 * n_tile + 2 * mt * A_SLOT_COLS > tmem_limit) nacc = 1;
          const int acc_cols = mt * nacc * 2 * n_tile;
          int nb = min(T * chunks, min(4, (tmem_limit - acc_cols) / (mt * A_SLOT_COLS)));
          if (knobs().nb) nb = min(nb, max(1, knobs().nb));
          if (nb < 1 || (nb < 2 && T * chunks > 1)) continue;
          // A set advances by SPLIT_SETS / mt ring entries per row; with a shallower
          // ring its parity wait on ring_empty would alias a phase it never observed.
          if (T * chunks > nb && nb < (SPLIT_SETS + mt - 1) / mt) continue;
          if (knobs().mt && knobs().mt != mt) continue;
          // halo ring; a persistent CTA keeps two slots even for a one-chunk conv (the
          // next tile's halo loads under this tile) and stages its epilogue elsewhere
          const int stage_bytes = 4 * SPLIT_SETS * STAGE_BYTES_PER_WARP;
          int nh = persist ? 2 : min(chunks, 2);
          auto smem_need = [&](int nh_, int nb_) {
            const long long halo = persist ? (long long)nh_ * halo_stride + stage_bytes
                                           : (long long)max(nh_ * halo_stride, stage_bytes);
            return halo + (long long)nb_ * b_stage + SMEM_SLACK;
          };
          while (nb > max(2, (SPLIT_SETS + mt - 1) / mt) && smem_need(nh, nb) > smem_limit) --nb;
          // a multi-chunk conv wants two halo slots (load of chunk c+1 under the taps of
          // chunk c); where they do not fit (stride-2 3x3x3 convs: 180 KB per slot) a
          // one-tile plan runs with ONE slot -- the next chunk's halo then loads after the
          // current chunk's last tap (protocol checked by tools/halo_protocol_sim.py); still
          // ahead of conv_umma.cu's per-tap tiles (64 -> 256 k333 s2: 64 TFLOP/s there)
          if (smem_need(nh, nb) > smem_limit && !persist && nh == 2) nh = 1;
          if (smem_need(nh, nb) > smem_limit) continue;
          while (nh < min(chunks, MAX_RING) && nh * halo_stride < 64 * 1024 &&
                 smem_need(nh + 1, nb) <= smem_limit)
            ++nh;
          const long long tiles = (long long)pw_ceil_div(c.ow, cb[0]) * pw_ceil_div(c.oh, cb[1]) *
                                  pw_ceil_div(c.od, cb[2]) * c.n * slabs;
          if (tiles > 0x7fffffffLL) continue;
          // cycle model
          // measured: ~430 cycles per tcgen05.commit, >= ~50 per MMA
          const double kstep = 1.5 * n_tile > 100.0 ? 1.5 * n_tile : 100.0;
          const double per_ct = 430.0 + mt * 4.0 * kstep;
          double cta = chunks * (T * per_ct + (nh > 1 ? 0.25 : 1.0) * hrows * 2.0) + 3000.0 +
                       mt * n_tile * 40.0;
          // Two co-resident CTAs overlap one CTA's prologue / halo load / epilogue
          // with the other's main loop, but each has half the split warps: the main
          // loop of a CTA runs at the split rate (measured ~1250 cycles per 128-row
          // A tile per set) when that is slower than the MMA issue rate.
          if (cps == 2) {
            const double split_ct = mt * 1250.0;
            const double main_ct = per_ct > split_ct ? per_ct : split_ct;
            const double fixed = cta - chunks * T * per_ct;
            cta = chunks * T * main_ct + 0.35 * fixed;
          }
          const double waves = (double)((tiles + 148 * cps - 1) / (148 * cps));
          const double cost = waves * cta;
          if (best < 0 || cost < best) {
            best = cost;
            HaloParams& p = plan.p;
            p = HaloParams{};
            p.bx = b[0]; p.by = b[1]; p.bz = b[2];
            p.lbx = ilog2(b[0]); p.lby = ilog2(b[1]);
            p.mt = mt; p.mt_axis = axis;
            p.cbx = cb[0]; p.cby = cb[1]; p.cbz = cb[2];
            p.hx = h[0]; p.hy = h[1]; p.hz = h[2];
            const int hstr[3] = {1, h[0], h[0] * h[1]};
            p.mt_halo_off = mt == 2 ? b[axis] * st[axis] * hstr[axis] : 0;
            p.halo_stride = halo_stride;
            p.halo_region = persist ? nh * halo_stride : max(nh * halo_stride, stage_bytes);
            p.stage_bytes = persist ? stage_bytes : 0;
            p.stage_off = persist ? p.halo_region + nb * b_stage : 0;
            p.total_tiles = (int)tiles; p.slabs = slabs; p.sp_tiles = (int)(tiles / slabs);
            p.fold = fold; p.fold_n = fold_n; p.fold_shift = c.dw;
            p.out_bx = fold ? cb[0] : b[0];
            p.tiles_x = pw_ceil_div(c.ow, cb[0]); p.tiles_y = pw_ceil_div(c.oh, cb[1]);
            p.tiles_z = pw_ceil_div(c.od, cb[2]);
            p.fd_sp = make_fastdiv(p.sp_tiles); p.fd_tx = make_fastdiv(p.tiles_x);
            p.fd_ty = make_fastdiv(p.tiles_y); p.fd_tz = make_fastdiv(p.tiles_z);
            p.n_tile = n_tile; p.nh = nh; p.nb = nb; p.nacc = nacc;
            int cols = acc_cols + nb * mt * A_SLOT_COLS, pw2 = 32;
            while (pw2 < cols) pw2 <<= 1;
            p.tmem_cols = pw2;
            p.cps = cps;
            plan.smem = (size_t)p.halo_region + (size_t)nb * b_stage + p.stage_bytes + SMEM_SLACK;
            plan.grid = dim3((unsigned)(persist ? min(tiles, (long long)sm_count() * cps) : tiles));
            plan.ok = true;
            plan.cost = cost;
          }
        }
      }
    }
  }
  if (!plan.ok) return plan;
  HaloParams& p = plan.p;
  p.od = c.od; p.oh = c.oh; p.ow = c.ow; p.cout = c.cout;
  p.kh = c.kh; p.kw = fold ? 1 : c.kw; p.n_taps = T; p.chunks = chunks;
  p.sd = c.sd; p.sh = c.sh; p.sw = c.sw; p.pd = c.pd; p.ph = c.ph; p.pw = c.pw;
  p.dd = c.dd; p.dh = c.dh; p.dw = c.dw;
  p.out_ld = c.out_ld; p.res_ld = c.res_ld; p.act = c.act; p.act_channels = c.act_channels;
  plan.c = c;
  return plan;
}

}  // namespace

PW_API int pw_conv_halo_supported(const pw_conv_desc* d) {
  if (!d) return 0;
  if (d->sd < 1 || d->sh < 1 || d->sw < 1) return 0;
  if (encode_tiled_fn() == nullptr) return 0;
  return make_plan(*d).ok ? 1 : 0;
}

namespace {

// Encodes the three tensor maps of a plan and launches its kernel variant.
// `w_rows` = rows of the (pre-split) weight matrices [w_rows, n_taps * cin].
int launch_halo(HaloPlan plan /* copy: per-launch fields are filled here */,
                const pw_conv_desc* d, long long w_rows, const float* x, const float* wt_hi,
                const float* wt_lo, const float* scale, const float* bias,
                const float* residual, float* y, void* stream) {
  EncodeTiledFn enc = encode_tiled_fn();
  PW_REQUIRE(enc != nullptr);
  const pw_conv_desc& c = plan.c;
  HaloParams& p = plan.p;
  p.out_ld = d->out_ld; p.res_ld = d->res_ld; p.act = d->act; p.act_channels = d->act_channels;
  p.scale = scale; p.bias = bias; p.res = residual; p.y = y;
  p.dbg = knobs().dbg;
  // one accumulator column receives n_taps * chunks * 4 / nacc dependent MMAs of K = 8
  p.gain = umma_chain_gain((long long)p.n_taps * p.chunks * BLOCK_K / p.nacc);
  cudaStream_t st = (cudaStream_t)stream;

  CUtensorMap ma, mbh, mbl;
  {
    cuuint64_t gdim[5] = {(cuuint64_t)c.cin, (cuuint64_t)c.w, (cuuint64_t)c.h, (cuuint64_t)c.d,
                          (cuuint64_t)c.n};
    cuuint64_t gstr[4] = {(cuuint64_t)c.in_ld * 4, (cuuint64_t)c.w * c.in_ld * 4,
                          (cuuint64_t)c.h * c.w * c.in_ld * 4,
                          (cuuint64_t)c.d * c.h * c.w * c.in_ld * 4};
    cuuint32_t box[5] = {BLOCK_K, (cuuint32_t)p.hx, (cuuint32_t)p.hy, (cuuint32_t)p.hz, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(x), gdim, gstr,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 1000 + (int)r;
  }
  const long long K = (long long)p.n_taps * c.cin;
  for (int i = 0; i < 2; ++i) {
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)w_rows};
    cuuint64_t gstr[1] = {(cuuint64_t)K * 4};
    cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)p.n_tile};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(i == 0 ? &mbh : &mbl, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<float*>(i == 0 ? wt_hi : wt_lo), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 1000 + (int)r;
  }

  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<2, 1>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(conv_halo_kernel<1, 2>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT_2CTA);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const bool want_ts = knobs().ts;
  const size_t n_cta = (size_t)plan.grid.x * plan.grid.y;
  if (want_ts) {
    cudaMalloc(&p.ts, n_cta * 16 * sizeof(long long));
    cudaMemset(p.ts, 0, n_cta * 16 * sizeof(long long));
  }
  static const int pdl_knob = getenv("PW_HALO_PDL") ? atoi(getenv("PW_HALO_PDL")) : 1;
  p.pdl = pdl_knob && !want_ts;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = plan.grid;
    cfg.blockDim = dim3((unsigned)halo_threads(plan.sets == 1 ? 1 : 2));
    cfg.dynamicSmemBytes = plan.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = p.pdl ? 1 : 0;
    cudaError_t e = plan.sets == 1
                        ? cudaLaunchKernelEx(&cfg, conv_halo_kernel<1, 2>, ma, mbh, mbl, p)
                        : cudaLaunchKernelEx(&cfg, conv_halo_kernel<2, 1>, ma, mbh, mbl, p);
    if (e != cudaSuccess) return (int)e;
  }
  PW_LAUNCH_CHECK();
  if (want_ts) {
    cudaStreamSynchronize(st);
    long long* h = (long long*)malloc(n_cta * 16 * sizeof(long long));
    cudaMemcpy(h, p.ts, n_cta * 16 * sizeof(long long), cudaMemcpyDeviceToHost);
    double sum[16] = {0};
    for (size_t i = 0; i < n_cta; ++i)
      for (int k = 0; k < 16; ++k) sum[k] += (double)(h[i * 16 + k] - h[i * 16]);
    fprintf(stderr, "[halo ts] grid %u x %u mt %d n_tile %d nh %d nb %d nacc %d box %dx%dx%d halo %dx%dx%d smem %zu tmem %d sets %d fold %d:",
            plan.grid.x, plan.grid.y, p.mt, p.n_tile, p.nh, p.nb, p.nacc, p.cbx, p.cby, p.cbz,
            p.hx, p.hy, p.hz, plan.smem, p.tmem_cols, plan.sets, p.fold);
    for (int k = 0; k < 16; ++k) fprintf(stderr, " t%d=%.0f", k, sum[k] / n_cta);
    fprintf(stderr, "\n");
    free(h);
    cudaFree(p.ts);
  }
  pw_count_launch(1);
  return 0;
}

int check_halo_args(const pw_conv_desc* d, const float* x, const float* wt_hi,
                    const float* wt_lo, const float* residual, const float* y) {
  PW_REQUIRE(d && x && wt_hi && wt_lo && y);
  PW_REQUIRE(d->sd >= 1 && d->sh >= 1 && d->sw >= 1);
  PW_REQUIRE(d->out_ld >= d->cout);
  PW_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)wt_hi & 15) == 0 &&
             ((uintptr_t)wt_lo & 15) == 0);
  PW_REQUIRE(residual == nullptr || d->res_ld >= d->cout);
  PW_REQUIRE(d->act_channels >= 0 && (d->act_channels & 3) == 0);
  return 0;
}

}  // namespace

PW_API int pw_conv_halo_fwd(const pw_conv_desc* d, const float* x, const float* wt_hi,
                            const float* wt_lo, const float* scale, const float* bias,
                            const float* residual, float* y, void* stream) {
  if (int rc = check_halo_args(d, x, wt_hi, wt_lo, residual, y)) return rc;
  const HaloPlan& plan = make_plan(*d);
  PW_REQUIRE(plan.ok);
  return launch_halo(plan, d, d->cout, x, wt_hi, wt_lo, scale, bias, residual, y, stream);
}

// ---- x-tap-folded variant --------------------------------------------------------
PW_API int pw_conv_fold_n(int cout) { return cout <= 16 ? 16 : 32; }

PW_API int pw_conv_fold_supported(const pw_conv_desc* d) {
  if (!d || knobs().fold == 0) return 0;
  if (d->sd < 1 || d->sh < 1 || d->sw != 1 || d->kw < 2) return 0;
  if (encode_tiled_fn() == nullptr) return 0;
  // Feasibility only; the caller decides (preworld_b200/ops.py: USE_FOLD, off by
  // default).  Measured (tools/umma_probe.py and the full step): the main loop gets
  // ~3x shorter, but a folded CTA covers one M tile of (bx - 2) outputs per row
  // group, so the per-CTA prologue + halo latency + epilogue (not overlapped: one
  // CTA per SM) is paid 2.3x as often -- inside the step 32->32 k333 runs 334 us
  // folded vs 309 us unfolded, and every extra N slab re-splits the halo (32->64:
  // +70 %).  It becomes the faster kernel once CTAs are persistent and prefetch
  // the next tile's halo.
  return make_fold_plan(*d, pw_conv_fold_n(d->cout)).ok ? 1 : 0;
}

PW_API int pw_conv_fold_fwd(const pw_conv_desc* d, const float* x, const float* wf_hi,
                            const float* wf_lo, const float* scale, const float* bias,
                            const float* residual, float* y, void* stream) {
  if (int rc = check_halo_args(d, x, wf_hi, wf_lo, residual, y)) return rc;
  const int fold_n = pw_conv_fold_n(d->cout);
  const HaloPlan& plan = make_fold_plan(*d, fold_n);
  PW_REQUIRE(plan.ok);
  const long long w_rows = (long long)pw_ceil_div(d->cout, fold_n) * d->kw * fold_n;
  return launch_halo(plan, d, w_rows, x, wf_hi, wf_lo, scale, bias, residual, y, stream);
}
