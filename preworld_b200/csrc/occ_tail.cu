// OccHead tail, fused: occ_pred_conv (1x1x1 16->8 + BN + ReLU, 1x1x1 8->18) ->
// argmax over the classes -> uint8 occupancy (+ geometry) grid in the reference's
// [X,Y,Z] order.  Reference: heads/occupancy_head.py:95-105,147-162 and
// detectors/preworld.py:196-221.  SURVEY §8b `pw_occhead_argmax`.
//
// HBM-bound and tiny in arithmetic (272 FMA per voxel): the 16-channel feature rows
// are read once, coalesced, in the library's [Z,Y,X] voxel order; the 8- and
// 18-channel intermediates (20 + 46 MB at 200x200x16) never exist unless the caller
// asks for the logits; the class bytes are transposed through shared memory so that
// the [X,Y,Z] grid is written in 16-byte runs along z.
#include "common.cuh"
#include "../../include/preworld_b200.h"

namespace {

constexpr int MAX_IN = 32, MAX_MID = 16, MAX_CLS = 32;
constexpr int XT = 32;                     // x voxels per block tile

// weights in shared memory: w0 [mid][cin], s0/b0 [mid], w1 [ncls][mid], b1 [ncls]
template <int CIN, int MID>
__global__ void __launch_bounds__(512)
occ_tail_kernel(const float* __restrict__ feat, int feat_ld, const float* __restrict__ w0,
                const float* __restrict__ s0, const float* __restrict__ b0,
                const float* __restrict__ w1, const float* __restrict__ b1, int ncls,
                float* __restrict__ logits, int logits_ld, unsigned char* __restrict__ occ,
                unsigned char* __restrict__ geo, int free_idx, int geo_value, int gx, int gy,
                int gz, int zt) {
  __shared__ float sw0[MID * CIN], ss0[MID], sb0[MID], sw1[MAX_CLS * MID], sb1[MAX_CLS];
  __shared__ unsigned char scls[XT * 32];  // [xl][z within the z tile]
  for (int i = threadIdx.x; i < MID * CIN; i += blockDim.x) sw0[i] = __ldg(w0 + i);
  for (int i = threadIdx.x; i < MID; i += blockDim.x) {
    ss0[i] = s0 ? __ldg(s0 + i) : 1.f;
    sb0[i] = b0 ? __ldg(b0 + i) : 0.f;
  }
  for (int i = threadIdx.x; i < ncls * MID; i += blockDim.x) sw1[i] = __ldg(w1 + i);
  for (int i = threadIdx.x; i < ncls; i += blockDim.x) sb1[i] = b1 ? __ldg(b1 + i) : 0.f;
  __syncthreads();

  const int xl = threadIdx.x % XT, zl = threadIdx.x / XT;   // zl < zt (blockDim = XT * zt)
  const int xtiles = (gx + XT - 1) / XT, ztiles = (gz + zt - 1) / zt;
  const long long tiles = (long long)xtiles * gy * ztiles;
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int xt = (int)(t % xtiles);
    long long r = t / xtiles;
    const int y = (int)(r % gy);
    const int zb = (int)(r / gy) * zt;
    const int x = xt * XT + xl, z = zb + zl;
    int arg = 0;
    if (x < gx && z < gz) {
      const long long v = ((long long)z * gy + y) * gx + x;
      const float4* src = reinterpret_cast<const float4*>(feat + v * feat_ld);
      float f[CIN];
#pragma unroll
      for (int c = 0; c < CIN / 4; ++c) {
        const float4 q = __ldg(src + c);
        f[c * 4] = q.x; f[c * 4 + 1] = q.y; f[c * 4 + 2] = q.z; f[c * 4 + 3] = q.w;
      }
      float h[MID];
#pragma unroll
      for (int m = 0; m < MID; ++m) {
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < CIN; ++c) a = fmaf(f[c], sw0[m * CIN + c], a);
        h[m] = fmaxf(fmaf(a, ss0[m], sb0[m]), 0.f);
      }
      float best = 0.f;
      float* lrow = logits ? logits + v * logits_ld : nullptr;
      for (int k = 0; k < ncls; ++k) {
        float a = 0.f;
#pragma unroll
        for (int m = 0; m < MID; ++m) a = fmaf(h[m], sw1[k * MID + m], a);
        a += sb1[k];
        if (lrow) lrow[k] = a;
        if (k == 0 || a > best) { best = a; arg = k; }   // first maximum wins (torch.argmax)
      }
    }
    scls[xl * 32 + zl] = (unsigned char)arg;
    __syncthreads();
    // output order: (x, y, z) with z fastest -> consecutive threads write consecutive z
    {
      const int oz = threadIdx.x % zt, ox = threadIdx.x / zt;
      const int xo = xt * XT + ox, zo = zb + oz;
      if (xo < gx && zo < gz) {
        const unsigned char c = scls[ox * 32 + oz];
        const long long o = ((long long)xo * gy + y) * gz + zo;
        occ[o] = c;
        if (geo) geo[o] = (unsigned char)(c != free_idx ? 0 : geo_value);
      }
    }
    __syncthreads();
  }
}

}  // namespace

PW_API int pw_occhead_tail(const float* feat, int feat_ld, int cin, const float* w0,
                           const float* scale0, const float* bias0, int mid, const float* w1,
                           const float* bias1, int ncls, float* logits, int logits_ld,
                           unsigned char* occ, unsigned char* geo, int free_idx, int geo_value,
                           int gx, int gy, int gz, void* stream) {
  PW_REQUIRE(feat && w0 && w1 && occ);
  PW_REQUIRE(cin == 16 && mid == 8);                   // the OccHead of the PreWorld configs
  PW_REQUIRE(ncls >= 1 && ncls <= MAX_CLS);
  PW_REQUIRE(feat_ld >= cin && (feat_ld & 3) == 0 && ((uintptr_t)feat & 15) == 0);
  PW_REQUIRE(logits == nullptr || logits_ld >= ncls);
  PW_REQUIRE(gx > 0 && gy > 0 && gz > 0 && (long long)gx * gy * gz < (1ll << 31));
  const int zt = gz < 16 ? gz : 16;
  const int xtiles = (gx + XT - 1) / XT, ztiles = (gz + zt - 1) / zt;
  const long long tiles = (long long)xtiles * gy * ztiles;
  const int blocks = (int)(tiles < 148 * 8 ? tiles : 148 * 8);
  occ_tail_kernel<16, 8><<<blocks, XT * zt, 0, (cudaStream_t)stream>>>(
      feat, feat_ld, w0, scale0, bias0, w1, bias1, ncls, logits, logits_ld, occ, geo, free_idx,
      geo_value, gx, gy, gz, zt);
  PW_LAUNCH_CHECK();
  pw_count_launch(1);
  return 0;
}
