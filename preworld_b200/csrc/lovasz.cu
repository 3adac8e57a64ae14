// Lovasz-Softmax loss: lovasz_softmax / lovasz_softmax_flat /
// lovasz_grad / flatten_probas of mmdet3d/models/detectors/lovasz_softmax.py:20-33,
// 157-239 as called from preworld.py:155 (classes='present', per_image=False,
// ignore=empty_idx, camera_mask).
//
//   1. lovasz_keys_kernel: one pass over the probabilities (or logits: softmax in
//      registers): per class c the error |[t==c] - p_c| of every voxel becomes a sort
//      key, (voxel index | fg bit) its value; voxels that flatten_probas drops
//      (label == ignore, camera mask off) get key -1 and sort behind every kept voxel.
//      Counts: kept voxels P, foreground G_c per class.
//   2. cub::DeviceRadixSort (library sort, like the reference's torch.sort): ONE sort of
//      64-bit keys (class << 32 | ~error bits), 37 significant bits -- the segmented
//      variant runs one CTA per segment and took 5 ms for 18 x 640 k.
//   3. lovasz_scan_kernel: one CTA per present class walks its sorted segment with a
//      running foreground count -- Jaccard index, its first difference (lovasz_grad) and
//      the dot product with the sorted errors, all in fp32 like the reference; writes
//      d loss / d p[v,c] = sign(p - fg) * grad_rank / #present back to the voxel.
//   4. softmax_backward_kernel (only when the input was logits).
#include <cub/cub.cuh>

#include "common.cuh"
#include "../../include/preworld_b200.h"

namespace {

constexpr int LV_MAX_CL = 32;
constexpr int LV_TPB = 256;
constexpr unsigned FG_BIT = 0x80000000u;

struct LvLayout {          // carve-up of the caller's workspace
  size_t keys_in, keys_out, vals_in, vals_out, offsets, counts, cub_temp, cub_bytes, total;
};

constexpr int KEY_BITS = 32 + 5;       // error bits + class id (n_cls <= 32)

size_t cub_temp_bytes(long long n_vox, int n_cls) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr,
                                  (unsigned long long*)nullptr, (const unsigned*)nullptr,
                                  (unsigned*)nullptr, (int)(n_vox * n_cls), 0, KEY_BITS);
  return bytes;
}

LvLayout lv_layout(long long n_vox, int n_cls) {
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  LvLayout L;
  const size_t kv = up((size_t)n_vox * n_cls * 4);
  size_t o = 0;
  L.keys_in = o; o += 2 * kv;          // 64-bit keys
  L.keys_out = o; o += 2 * kv;
  L.vals_in = o; o += kv;
  L.vals_out = o; o += kv;
  L.offsets = o; o += up((size_t)(n_cls + 1) * 4);
  L.counts = o; o += up((size_t)(LV_MAX_CL + 4) * 8);      // double: P, #present, loss, G_c...
  L.cub_bytes = cub_temp_bytes(n_vox, n_cls);
  L.cub_temp = o; o += up(L.cub_bytes);
  L.total = o;
  return L;
}

// counts (unsigned long long): [0] P kept voxels, [1 + c] G_c
template <int C_MAX>
__global__ void __launch_bounds__(LV_TPB)
lovasz_keys_kernel(const float* __restrict__ x, int ld, int is_logits,
                   const unsigned char* __restrict__ target, const unsigned char* __restrict__ cam,
                   long long n, int C, int ignore_label, unsigned long long* __restrict__ keys,
                   unsigned* __restrict__ vals, unsigned long long* __restrict__ counts) {
  __shared__ unsigned s_cnt[C_MAX + 1];
  for (int i = threadIdx.x; i <= C; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n;
       v += (long long)gridDim.x * blockDim.x) {
    const int t = target[v];
    const bool kept = t != ignore_label && (cam == nullptr || cam[v] != 0);
    float p[C_MAX];
    float mx = -3.0e38f;
#pragma unroll
    for (int c = 0; c < C_MAX; ++c)
      if (c < C) { p[c] = __ldg(x + v * ld + c); mx = fmaxf(mx, p[c]); }
    if (is_logits) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < C_MAX; ++c)
        if (c < C) { p[c] = expf(p[c] - mx); s += p[c]; }
      const float inv = 1.f / s;
#pragma unroll
      for (int c = 0; c < C_MAX; ++c)
        if (c < C) p[c] *= inv;
    }
    if (kept) {
      atomicAdd(&s_cnt[0], 1u);
      if (t < C) atomicAdd(&s_cnt[1 + t], 1u);
    }
#pragma unroll
    for (int c = 0; c < C_MAX; ++c)
      if (c < C) {
        const bool fg = t == c;
        // ascending key order == class by class, errors descending, dropped voxels last
        // (errors are >= 0, so their bit patterns are monotone)
        const unsigned eb = kept ? ~__float_as_uint(fabsf((fg ? 1.f : 0.f) - p[c])) : 0xFFFFFFFFu;
        keys[(long long)c * n + v] = ((unsigned long long)c << 32) | eb;
        vals[(long long)c * n + v] = (unsigned)v | ((kept && fg) ? FG_BIT : 0u);
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= C; i += blockDim.x)
    if (s_cnt[i]) atomicAdd(counts + i, (unsigned long long)s_cnt[i]);
}

// One CTA per class.  out[0] += loss_c / #present; grad_p (optional) [n, C].
__global__ void __launch_bounds__(1024)
lovasz_scan_kernel(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ vals, long long n,
                   int C, const unsigned long long* __restrict__ counts,
                   const float* __restrict__ x_unused, double* __restrict__ loss_acc,
                   float* __restrict__ grad_p) {
  const int c = blockIdx.x;
  const long long P = (long long)counts[0];
  const float G = (float)counts[1 + c];
  int present = 0;
  for (int k = 0; k < C; ++k) present += counts[1 + k] > 0 ? 1 : 0;
  const bool active = counts[1 + c] > 0 && present > 0;
  const float inv_present = present > 0 ? 1.f / (float)present : 0.f;
  const unsigned long long* kc = keys + (long long)c * n;
  const unsigned* vc = vals + (long long)c * n;
  __shared__ float s_warp[32];
  __shared__ float s_carry;       // foreground count before this chunk (exact in fp32: < 2^24)
  __shared__ double s_loss;
  if (threadIdx.x == 0) { s_carry = 0.f; s_loss = 0.0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float my_loss = 0.f;
  for (long long base = 0; base < n; base += blockDim.x) {
    const long long i = base + threadIdx.x;
    const bool in = i < n;
    const unsigned val = in ? vc[i] : 0u;
    const float err = in ? __uint_as_float(~(unsigned)kc[i]) : 0.f;   // (dropped: beyond P)
    const float fg = (val & FG_BIT) ? 1.f : 0.f;
    // inclusive scan of fg over the chunk
    float sc = fg;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float up = __shfl_up_sync(0xffffffffu, sc, o);
      if (lane >= o) sc += up;
    }
    if (lane == 31) s_warp[wid] = sc;
    __syncthreads();
    if (wid == 0) {
      float w = lane < (blockDim.x >> 5) ? s_warp[lane] : 0.f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += up;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const float carry = s_carry;
    const float cum = carry + sc + (wid > 0 ? s_warp[wid - 1] : 0.f);     // inclusive
    float g = 0.f;
    if (in && i < P && active) {
      // lovasz_grad (lovasz_softmax.py:20-33), fp32 like the reference
      const float pos = (float)(i + 1);
      const float jac = 1.f - (G - cum) / (G + (pos - cum));
      float jac_prev = 0.f;
      if (i > 0) {
        const float cum_prev = cum - fg;
        jac_prev = 1.f - (G - cum_prev) / (G + ((pos - 1.f) - cum_prev));
      }
      g = jac - jac_prev;
      my_loss += err * g;
    }
    if (grad_p != nullptr && in) {
      // errors = |fg - p|: d err / d p = sign(p - fg) = fg ? -1 : +1  (0 where err == 0
      // in torch's abs backward; p == fg exactly contributes nothing either way)
      const unsigned idx = val & ~FG_BIT;
      const float sgn = err == 0.f ? 0.f : (fg != 0.f ? -1.f : 1.f);
      grad_p[(long long)idx * C + c] = sgn * g * inv_present;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = cum;
    __syncthreads();
  }
  // block reduce the loss
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my_loss += __shfl_xor_sync(0xffffffffu, my_loss, o);
  if (lane == 0 && my_loss != 0.f) atomicAdd(&s_loss, (double)my_loss);
  __syncthreads();
  if (threadIdx.x == 0 && active) atomicAdd(loss_acc, s_loss * (double)inv_present);
}

__global__ void lovasz_finalize_kernel(const double* __restrict__ loss_acc, float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) loss[0] = (float)loss_acc[0];
}

// d L / d logits from d L / d probas (softmax rows recomputed from the logits)
template <int C_MAX>
__global__ void __launch_bounds__(LV_TPB)
softmax_backward_kernel(const float* __restrict__ logits, int ld, const float* __restrict__ gp,
                        long long n, int C, float* __restrict__ gl) {
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n;
       v += (long long)gridDim.x * blockDim.x) {
    float p[C_MAX];
    float mx = -3.0e38f;
#pragma unroll
    for (int c = 0; c < C_MAX; ++c)
      if (c < C) { p[c] = __ldg(logits + v * ld + c); mx = fmaxf(mx, p[c]); }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C_MAX; ++c)
      if (c < C) { p[c] = expf(p[c] - mx); s += p[c]; }
    const float inv = 1.f / s;
    float dot = 0.f;
    float g[C_MAX];
#pragma unroll
    for (int c = 0; c < C_MAX; ++c)
      if (c < C) { p[c] *= inv; g[c] = __ldg(gp + v * C + c); dot += p[c] * g[c]; }
#pragma unroll
    for (int c = 0; c < C_MAX; ++c)
      if (c < C) gl[v * C + c] = p[c] * (g[c] - dot);
  }
}

}  // namespace

PW_API long long pw_lovasz_workspace_bytes(long long n_vox, int n_cls) {
  if (n_vox <= 0 || n_cls <= 0 || n_cls > LV_MAX_CL || n_vox * n_cls >= (1ll << 31)) return -1;
  return (long long)lv_layout(n_vox, n_cls).total;
}

PW_API int pw_lovasz_softmax(const float* x, int ld, int is_logits, const unsigned char* target,
                             const unsigned char* camera_mask, long long n_vox, int n_cls,
                             int ignore_label, void* workspace, long long workspace_bytes,
                             float* loss, float* grad_probas, void* stream) {
  PW_REQUIRE(x && target && workspace && loss);
  PW_REQUIRE(n_vox > 0 && n_cls > 0 && n_cls <= LV_MAX_CL && ld >= n_cls);
  PW_REQUIRE(n_vox * n_cls < (1ll << 31));
  const LvLayout L = lv_layout(n_vox, n_cls);
  PW_REQUIRE(workspace_bytes >= (long long)L.total);
  char* ws = (char*)workspace;
  unsigned long long* keys_in = (unsigned long long*)(ws + L.keys_in);
  unsigned long long* keys_out = (unsigned long long*)(ws + L.keys_out);
  unsigned* vals_in = (unsigned*)(ws + L.vals_in);
  unsigned* vals_out = (unsigned*)(ws + L.vals_out);
  unsigned long long* counts = (unsigned long long*)(ws + L.counts);
  double* loss_acc = (double*)(counts + LV_MAX_CL + 2);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(counts, 0, (LV_MAX_CL + 4) * 8, st);
  if (e != cudaSuccess) return (int)e;
  const int blocks = (int)min((long long)148 * 8, (n_vox + LV_TPB - 1) / LV_TPB);
  if (n_cls <= 20)
    lovasz_keys_kernel<20><<<blocks, LV_TPB, 0, st>>>(x, ld, is_logits, target, camera_mask, n_vox,
                                                      n_cls, ignore_label, keys_in, vals_in, counts);
  else
    lovasz_keys_kernel<LV_MAX_CL><<<blocks, LV_TPB, 0, st>>>(
        x, ld, is_logits, target, camera_mask, n_vox, n_cls, ignore_label, keys_in, vals_in, counts);
  PW_LAUNCH_CHECK();
  size_t cub_bytes = L.cub_bytes;
  e = cub::DeviceRadixSort::SortPairs(ws + L.cub_temp, cub_bytes, keys_in, keys_out, vals_in,
                                      vals_out, (int)(n_vox * n_cls), 0, KEY_BITS, st);
  if (e != cudaSuccess) return (int)e;
  lovasz_scan_kernel<<<n_cls, 1024, 0, st>>>(keys_out, vals_out, n_vox, n_cls, counts, nullptr,
                                             loss_acc, grad_probas);
  PW_LAUNCH_CHECK();
  lovasz_finalize_kernel<<<1, 32, 0, st>>>(loss_acc, loss);
  PW_LAUNCH_CHECK(); pw_count_launch(3);
  return 0;
}

PW_API int pw_softmax_backward(const float* logits, int ld, const float* grad_probas,
                               long long n_vox, int n_cls, float* grad_logits, void* stream) {
  PW_REQUIRE(logits && grad_probas && grad_logits && n_vox > 0 && n_cls > 0 &&
             n_cls <= LV_MAX_CL && ld >= n_cls);
  const int blocks = (int)min((long long)148 * 8, (n_vox + LV_TPB - 1) / LV_TPB);
  if (n_cls <= 20)
    softmax_backward_kernel<20><<<blocks, LV_TPB, 0, (cudaStream_t)stream>>>(
        logits, ld, grad_probas, n_vox, n_cls, grad_logits);
  else
    softmax_backward_kernel<LV_MAX_CL><<<blocks, LV_TPB, 0, (cudaStream_t)stream>>>(
        logits, ld, grad_probas, n_vox, n_cls, grad_logits);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}
