// Voxel SSC training losses: CE_ssc_loss, sem_scal_loss and
// geo_scal_loss of mmdet3d/models/detectors/loss.py:20-113 (called from
// preworld.py:151-154 / preworld_temporal_traj.py:190-193) as ONE pass over the
// logits plus a second pass for the gradient.
//
// The reference evaluates softmax three times and then ~55 masked reductions
// over [B,C,H,W,D] (three per class and loss); every one of them is a sum over
// voxels of a function of (p[v,:], target[v], mask[v]).  voxel_loss_stats_kernel
// reads each voxel's logits once (HBM bound: 4*C bytes + 2 per voxel), forms
// the softmax in registers and accumulates all those sums -- 3 + 3C + 5 numbers --
// in fp64 (warp shuffle -> shared -> one atomic per CTA and number).  The
// losses are closed forms of the sums (voxel_loss_finalize_kernel), and so are
// their gradients: voxel_loss_grad_kernel re-reads the logits once and writes
// d(w_ce*CE + w_sem*sem + w_geo*geo)/dlogits.
#include "common.cuh"
#include "../../include/preworld_b200.h"

namespace {

constexpr int MAX_CL = 32;
constexpr int TPB_L = 256;

// stats layout (doubles)
struct Idx {
  int C;
  __host__ __device__ int ce_num() const { return 0; }
  __host__ __device__ int ce_den() const { return 1; }
  __host__ __device__ int n_mask() const { return 2; }
  __host__ __device__ int sum_p(int i) const { return 3 + i; }
  __host__ __device__ int nom(int i) const { return 3 + C + i; }
  __host__ __device__ int cnt(int i) const { return 3 + 2 * C + i; }
  __host__ __device__ int g_inter() const { return 3 + 3 * C; }
  __host__ __device__ int g_snp() const { return 4 + 3 * C; }     // sum of non-empty probs
  __host__ __device__ int g_st() const { return 5 + 3 * C; }      // sum of the non-empty target
  __host__ __device__ int g_specnum() const { return 6 + 3 * C; }
  __host__ __device__ int g_specden() const { return 7 + 3 * C; }
  __host__ __device__ int total() const { return 8 + 3 * C; }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// softmax of one voxel's logits into p[0..C) (registers: every loop is unrolled over
// C_MAX and predicated on c < C), the logit of class t, the max and log(sum exp(l - max))
template <int C_MAX>
__device__ __forceinline__ void voxel_softmax(const float* __restrict__ row, int C, int t,
                                              float (&p)[C_MAX], float& l_t, float& mx, float& lse) {
  mx = -3.0e38f;
  l_t = 0.f;
#pragma unroll
  for (int c = 0; c < C_MAX; ++c)
    if (c < C) {
      p[c] = __ldg(row + c);
      mx = fmaxf(mx, p[c]);
      if (c == t) l_t = p[c];
    }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < C_MAX; ++c)
    if (c < C) {
      p[c] = expf(p[c] - mx);
      s += p[c];
    }
  const float inv = 1.f / s;
#pragma unroll
  for (int c = 0; c < C_MAX; ++c)
    if (c < C) p[c] *= inv;
  lse = logf(s);
}

template <int C_MAX>
__global__ void __launch_bounds__(TPB_L)
voxel_loss_stats_kernel(const float* __restrict__ logits, int ld, const unsigned char* __restrict__ target,
                        const unsigned char* __restrict__ cam, long long n, int C, int ignore_index,
                        int empty_idx, const float* __restrict__ class_w, double* __restrict__ stats) {
  const Idx ix{C};
  __shared__ double s_acc[8 + 3 * C_MAX];
  for (int i = threadIdx.x; i < ix.total(); i += blockDim.x) s_acc[i] = 0.0;
  __syncthreads();

  // per-thread partial sums: the per-class ones are kept in registers as floats
  // over a bounded number of voxels (<= 64 per flush) and flushed to fp64
  float sum_p[C_MAX], nom[C_MAX], cnt[C_MAX];
#pragma unroll
  for (int c = 0; c < C_MAX; ++c) sum_p[c] = nom[c] = cnt[c] = 0.f;
  double ce_num = 0, ce_den = 0, n_mask = 0, g_inter = 0, g_snp = 0, g_st = 0, g_sn = 0, g_sd = 0;
  int since_flush = 0;
  const int lane = threadIdx.x & 31;

  auto flush = [&]() {
#pragma unroll
    for (int c = 0; c < C_MAX; ++c) {
      if (c < C) {
        // <= 64 voxels x 32 lanes of probabilities in [0,1]: fp32 is exact enough
        // here (rel. 1e-7); the running totals are fp64
        const float a = warp_sum(sum_p[c]), b = warp_sum(nom[c]), d = warp_sum(cnt[c]);
        if (lane == 0) {
          if (a != 0.f) atomicAdd(&s_acc[ix.sum_p(c)], (double)a);
          if (b != 0.f) atomicAdd(&s_acc[ix.nom(c)], (double)b);
          if (d != 0.f) atomicAdd(&s_acc[ix.cnt(c)], (double)d);
        }
      }
      sum_p[c] = nom[c] = cnt[c] = 0.f;
    }
    since_flush = 0;
  };

  const long long stride = (long long)gridDim.x * blockDim.x;
  // all lanes of a warp run the same number of iterations (flush() shuffles)
  const long long n_iter = (n + stride - 1) / stride;
  long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (long long it = 0; it < n_iter; ++it, v += stride) {
    if (v < n) {
      float p[C_MAX];
      float l_t, mx, lse;
      const int t = target[v];
      voxel_softmax<C_MAX>(logits + v * ld, C, t, p, l_t, mx, lse);
      const bool cam_ok = cam == nullptr || cam[v] != 0;
      // CE (loss.py:20-30): ignore_index voxels drop out, no camera mask
      if (t != ignore_index && t < C) {
        const float w = __ldg(class_w + t);
        const float logp = l_t - mx - lse;
        ce_num += (double)(-w * logp);
        ce_den += (double)w;
      }
      // sem_scal (loss.py:33-80): mask = target != ignore [& camera]
      if (t != ignore_index && cam_ok) {
        n_mask += 1.0;
#pragma unroll
        for (int c = 0; c < C_MAX; ++c)
          if (c < C) {
            sum_p[c] += p[c];
            if (t == c) { nom[c] += p[c]; cnt[c] += 1.f; }
          }
      }
      // geo_scal (loss.py:83-113): every voxel; target = (t != empty) [& camera]
      float pe = 0.f;
#pragma unroll
      for (int c = 0; c < C_MAX; ++c)
        if (c == empty_idx) pe = p[c];
      const bool tg = (t != empty_idx) && cam_ok;
      g_snp += (double)(1.f - pe);
      if (tg) { g_inter += (double)(1.f - pe); g_st += 1.0; }
      else { g_sn += (double)pe; g_sd += 1.0; }
    }
    if (++since_flush == 64) flush();
  }
  flush();
  double sc[8] = {ce_num, ce_den, n_mask, g_inter, g_snp, g_st, g_sn, g_sd};
  const int slot[8] = {ix.ce_num(), ix.ce_den(), ix.n_mask(), ix.g_inter(),
                       ix.g_snp(),  ix.g_st(),   ix.g_specnum(), ix.g_specden()};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double a = warp_sum(sc[k]);
    if (lane == 0 && a != 0.0) atomicAdd(&s_acc[slot[k]], a);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ix.total(); i += blockDim.x)
    if (s_acc[i] != 0.0) atomicAdd(stats + i, s_acc[i]);
}

__device__ __forceinline__ double bce_to_one(double x) {      // -max(log x, -100)
  const double l = log(x);
  return -(l < -100.0 ? -100.0 : l);
}

// Per-class gradient coefficients of sem_scal wrt the masked probabilities:
//   dL/dp[v,i] = a[i] + (t_v == i ? b[i] : c[i])      (v in the mask)
struct SemCoef { double a, b, c; bool active; };
__device__ __forceinline__ SemCoef sem_coef(const double* st, const Idx& ix, int i, double count) {
  SemCoef k{0, 0, 0, false};
  const double cnt = st[ix.cnt(i)];
  if (!(cnt > 0.0) || !(count > 0.0)) return k;
  k.active = true;
  const double sp = st[ix.sum_p(i)], nom = st[ix.nom(i)], nm = st[ix.n_mask()];
  const double inv = 1.0 / count;
  if (nom > 0.0) {
    if (sp > 0.0) { k.a += inv / sp; k.b -= inv / nom; }      // precision
    k.b -= inv / nom;                                         // recall
  }
  const double specden = nm - cnt;
  const double specnum = nm - sp - cnt + nom;
  if (specden > 0.0 && specnum > 0.0) k.c += inv / specnum;   // specificity
  return k;
}

__global__ void voxel_loss_finalize_kernel(const double* __restrict__ st, int C,
                                           float* __restrict__ losses) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const Idx ix{C};
  losses[0] = (float)(st[ix.ce_num()] / st[ix.ce_den()]);
  double loss = 0.0, count = 0.0;
  const double nm = st[ix.n_mask()];
  for (int i = 0; i < C; ++i) {
    const double cnt = st[ix.cnt(i)];
    if (!(cnt > 0.0)) continue;
    count += 1.0;
    const double sp = st[ix.sum_p(i)], nom = st[ix.nom(i)];
    if (sp > 0.0) loss += bce_to_one(nom / sp);
    loss += bce_to_one(nom / cnt);
    if (nm - cnt > 0.0) loss += bce_to_one((nm - sp - cnt + nom) / (nm - cnt));
  }
  losses[1] = (float)(loss / count);
  const double inter = st[ix.g_inter()];
  losses[2] = (float)(bce_to_one(inter / st[ix.g_snp()]) + bce_to_one(inter / st[ix.g_st()]) +
                      bce_to_one(st[ix.g_specnum()] / st[ix.g_specden()]));
}

template <int C_MAX>
__global__ void __launch_bounds__(TPB_L)
voxel_loss_grad_kernel(const float* __restrict__ logits, int ld, const unsigned char* __restrict__ target,
                       const unsigned char* __restrict__ cam, long long n, int C, int ignore_index,
                       int empty_idx, const float* __restrict__ class_w,
                       const double* __restrict__ st, float w_ce, float w_sem, float w_geo,
                       float* __restrict__ grad, int grad_ld) {
  const Idx ix{C};
  __shared__ float s_a[C_MAX], s_b[C_MAX], s_c[C_MAX];
  __shared__ float s_geo[3], s_ce;
  if (threadIdx.x < C) {
    double count = 0.0;
    for (int i = 0; i < C; ++i) count += st[ix.cnt(i)] > 0.0 ? 1.0 : 0.0;
    const SemCoef k = sem_coef(st, ix, threadIdx.x, count);
    s_a[threadIdx.x] = (float)(w_sem * k.a);
    s_b[threadIdx.x] = (float)(w_sem * k.b);
    s_c[threadIdx.x] = (float)(w_sem * k.c);
  }
  if (threadIdx.x == 0) {
    // dgeo/dp_e = tgt * 2/inter - 1/snp - (1 - tgt)/specnum
    const double inter = st[ix.g_inter()], snp = st[ix.g_snp()], sn = st[ix.g_specnum()];
    s_geo[0] = inter > 0.0 ? (float)(w_geo * 2.0 / inter) : 0.f;
    s_geo[1] = snp > 0.0 ? (float)(w_geo / snp) : 0.f;
    s_geo[2] = sn > 0.0 ? (float)(w_geo / sn) : 0.f;
    s_ce = st[ix.ce_den()] > 0.0 ? (float)(w_ce / st[ix.ce_den()]) : 0.f;
  }
  __syncthreads();
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n;
       v += (long long)gridDim.x * blockDim.x) {
    float p[C_MAX];
    float l_t, mx, lse;
    const int t = target[v];
    voxel_softmax<C_MAX>(logits + v * ld, C, t, p, l_t, mx, lse);
    const bool cam_ok = cam == nullptr || cam[v] != 0;
    const bool in_sem = t != ignore_index && cam_ok;
    const bool tg = (t != empty_idx) && cam_ok;
    // g[c] = dL/dp[v,c]; dL/dl[v,c] = p_c (g_c - sum_k p_k g_k) + CE term
    float g[C_MAX];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < C_MAX; ++c)
      if (c < C) {
        float gc = in_sem ? s_a[c] + (t == c ? s_b[c] : s_c[c]) : 0.f;
        if (c == empty_idx) gc += (tg ? s_geo[0] : -s_geo[2]) - s_geo[1];
        g[c] = gc;
        dot += p[c] * gc;
      }
    const float wce = (t != ignore_index && t < C) ? __ldg(class_w + t) * s_ce : 0.f;
    float* out = grad + v * grad_ld;
#pragma unroll
    for (int c = 0; c < C_MAX; ++c)
      if (c < C) out[c] = p[c] * (g[c] - dot) + wce * (p[c] - (t == c ? 1.f : 0.f));
  }
}

// few, long-lived CTAs: the per-warp flush of 3C partial sums costs as much as a
// dozen voxels, so every thread should see many of them
int loss_blocks(long long n) { return (int)min((long long)148 * 2, (n + TPB_L - 1) / TPB_L); }
int grad_blocks(long long n) { return (int)min((long long)148 * 8, (n + TPB_L - 1) / TPB_L); }

}  // namespace

PW_API int pw_voxel_loss_stats_size(int n_cls) { return Idx{n_cls}.total(); }

PW_API int pw_voxel_loss_stats(const float* logits, int ld, const unsigned char* target,
                               const unsigned char* camera_mask, long long n_vox, int n_cls,
                               int ignore_index, int empty_idx, const float* class_weights,
                               double* stats, float* losses, void* stream) {
  PW_REQUIRE(logits && target && class_weights && stats && losses);
  PW_REQUIRE(n_vox > 0 && n_cls > 1 && n_cls <= MAX_CL && ld >= n_cls);
  PW_REQUIRE(empty_idx >= 0 && empty_idx < n_cls);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * Idx{n_cls}.total(), st);
  if (e != cudaSuccess) return (int)e;
  if (n_cls <= 20)
    voxel_loss_stats_kernel<20><<<loss_blocks(n_vox), TPB_L, 0, st>>>(
        logits, ld, target, camera_mask, n_vox, n_cls, ignore_index, empty_idx, class_weights,
        stats);
  else
    voxel_loss_stats_kernel<MAX_CL><<<loss_blocks(n_vox), TPB_L, 0, st>>>(
        logits, ld, target, camera_mask, n_vox, n_cls, ignore_index, empty_idx, class_weights,
        stats);
  PW_LAUNCH_CHECK();
  voxel_loss_finalize_kernel<<<1, 32, 0, st>>>(stats, n_cls, losses);
  PW_LAUNCH_CHECK(); pw_count_launch(2);
  return 0;
}

PW_API int pw_voxel_loss_grad(const float* logits, int ld, const unsigned char* target,
                              const unsigned char* camera_mask, long long n_vox, int n_cls,
                              int ignore_index, int empty_idx, const float* class_weights,
                              const double* stats, float w_ce, float w_sem, float w_geo,
                              float* grad_logits, int grad_ld, void* stream) {
  PW_REQUIRE(logits && target && class_weights && stats && grad_logits);
  PW_REQUIRE(n_vox > 0 && n_cls > 1 && n_cls <= MAX_CL && ld >= n_cls && grad_ld >= n_cls);
  PW_REQUIRE(empty_idx >= 0 && empty_idx < n_cls);
  if (n_cls <= 20)
    voxel_loss_grad_kernel<20><<<grad_blocks(n_vox), TPB_L, 0, (cudaStream_t)stream>>>(
        logits, ld, target, camera_mask, n_vox, n_cls, ignore_index, empty_idx, class_weights,
        stats, w_ce, w_sem, w_geo, grad_logits, grad_ld);
  else
    voxel_loss_grad_kernel<MAX_CL><<<grad_blocks(n_vox), TPB_L, 0, (cudaStream_t)stream>>>(
        logits, ld, target, camera_mask, n_vox, n_cls, ignore_index, empty_idx, class_weights,
        stats, w_ce, w_sem, w_geo, grad_logits, grad_ld);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

// ---- depth loss (view_transformer.py:736-789) -----------------------------------
// get_downsampled_gt_depth + get_depth_loss fused: one warp per feature-map cell takes
// the minimum non-zero lidar depth of its downsample x downsample patch, turns it into
// a bin label, and (foreground cells only) sums the binary cross entropy of the D
// depth probabilities against the one-hot label.
namespace {

__device__ __forceinline__ float bce_term(float p, bool one) {   // torch: log clamped at -100
  const float l = one ? logf(p) : log1pf(-p);
  return -fmaxf(l, -100.f);
}

__global__ void __launch_bounds__(256)
depth_loss_kernel(const float* __restrict__ gt, int bn, int H, int W, int ds,
                  const float* __restrict__ pred, long long s_img, long long s_d, long long s_y,
                  long long s_x, int D, float c0, float c2, int* __restrict__ labels,
                  double* __restrict__ sums) {
  const int h = H / ds, w = W / ds;
  const long long cells = (long long)bn * h * w;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
  double loss = 0.0, nfg = 0.0;
  for (long long cell = warp0; cell < cells; cell += nwarp) {
    const int x = (int)(cell % w);
    const int y = (int)((cell / w) % h);
    const long long img = cell / ((long long)w * h);
    // minimum over the patch, zeros (no lidar return) count as 1e5
    float mn = 1e5f;
    const float* g0 = gt + (img * H + (long long)y * ds) * W + (long long)x * ds;
    for (int i = lane; i < ds * ds; i += 32) {
      const float v = __ldg(g0 + (long long)(i / ds) * W + (i % ds));
      mn = fminf(mn, v == 0.f ? 1e5f : v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    const float bin = (mn - c0) / c2;                       // (d - (d0 - dstep)) / dstep
    const float kept = (bin < (float)(D + 1) && bin >= 0.f) ? bin : 0.f;
    const int label = (int)kept - 1;                        // .long(), one-hot class 0 dropped
    if (lane == 0) labels[cell] = label;
    if (label < 0) continue;                                // background cell
    const float* p0 = pred + img * s_img + y * s_y + x * s_x;
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) acc += bce_term(__ldg(p0 + d * s_d), d == label);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    loss += (double)acc;
    nfg += 1.0;
  }
  if (lane == 0 && nfg > 0.0) {
    atomicAdd(sums, loss);
    atomicAdd(sums + 1, nfg);
  }
}

__global__ void depth_loss_finalize_kernel(const double* __restrict__ sums, float weight,
                                           float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0)
    loss[0] = (float)(weight * sums[0] / (sums[1] > 1.0 ? sums[1] : 1.0));
}

// d loss / d pred = weight / max(1, n_fg) * (p - y) / max((1 - p) p, 1e-12)   (torch's BCE
// backward), zero on background cells; grad is [cells, D] contiguous
__global__ void __launch_bounds__(256)
depth_loss_grad_kernel(const int* __restrict__ labels, long long cells, int h, int w,
                       const float* __restrict__ pred, long long s_img, long long s_d,
                       long long s_y, long long s_x, int D, const double* __restrict__ sums,
                       float weight, float* __restrict__ grad) {
  const float scale = (float)(weight / (sums[1] > 1.0 ? sums[1] : 1.0));
  const long long total = cells * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long cell = i / D;
    const int d = (int)(i - cell * D);
    const int label = labels[cell];
    float g = 0.f;
    if (label >= 0) {
      const int x = (int)(cell % w);
      const int y = (int)((cell / w) % h);
      const long long img = cell / ((long long)w * h);
      const float p = __ldg(pred + img * s_img + y * s_y + x * s_x + d * s_d);
      g = scale * (p - (d == label ? 1.f : 0.f)) / fmaxf((1.f - p) * p, 1e-12f);
    }
    grad[i] = g;
  }
}

}  // namespace

PW_API int pw_depth_loss(const float* gt_depth, int bn, int H, int W, int downsample,
                         const float* depth_pred, long long stride_img, long long stride_d,
                         long long stride_y, long long stride_x, int D, float depth_min,
                         float depth_step, float weight, int* labels, double* sums, float* loss,
                         void* stream) {
  PW_REQUIRE(gt_depth && depth_pred && labels && sums && loss);
  PW_REQUIRE(bn > 0 && downsample > 0 && H % downsample == 0 && W % downsample == 0 && D > 0);
  PW_REQUIRE(depth_step > 0.f);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(sums, 0, 2 * sizeof(double), st);
  if (e != cudaSuccess) return (int)e;
  const long long cells = (long long)bn * (H / downsample) * (W / downsample);
  const int blocks = (int)min((long long)148 * 4, (cells + 7) / 8);
  // the reference subtracts the python double (d0 - dstep) from an fp32 tensor
  const float c0 = (float)((double)depth_min - (double)depth_step);
  depth_loss_kernel<<<blocks, 256, 0, st>>>(gt_depth, bn, H, W, downsample, depth_pred, stride_img,
                                            stride_d, stride_y, stride_x, D, c0, depth_step,
                                            labels, sums);
  PW_LAUNCH_CHECK();
  depth_loss_finalize_kernel<<<1, 32, 0, st>>>(sums, weight, loss);
  PW_LAUNCH_CHECK(); pw_count_launch(2);
  return 0;
}

PW_API int pw_depth_loss_grad(const int* labels, int bn, int h, int w, const float* depth_pred,
                              long long stride_img, long long stride_d, long long stride_y,
                              long long stride_x, int D, const double* sums, float weight,
                              float* grad, void* stream) {
  PW_REQUIRE(labels && depth_pred && sums && grad && bn > 0 && h > 0 && w > 0 && D > 0);
  const long long cells = (long long)bn * h * w;
  const int blocks = (int)min((long long)148 * 8, (cells * D + 255) / 256);
  depth_loss_grad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      labels, cells, h, w, depth_pred, stride_img, stride_d, stride_y, stride_x, D, sums, weight,
      grad);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

// ---- CustomFocalLoss (loss_utils/focal_loss.py:162-273; PreWorld's default voxel CE
// term, preworld.py:43,116-117,146-148) -------------------------------------------------
// Sigmoid focal loss per (voxel, class) in the form of mmcv 1.6.0's CUDA op
// (sigmoid_focal_loss_cuda_kernel.cuh; the file's own py_sigmoid_focal_loss is the same
// function up to rounding), times class_weight[c] * radial[h,w], summed over classes,
// mean over the kept voxels (target != ignore [& camera mask]), times loss_weight.
namespace {

__device__ __forceinline__ float focal_term(float x, bool hit, float gamma, float alpha) {
  const float p = 1.f / (1.f + expf(-x));
  if (hit) return -alpha * powf(1.f - p, gamma) * logf(fmaxf(p, 1.17549435e-38f));
  return -(1.f - alpha) * powf(p, gamma) * logf(fmaxf(1.f - p, 1.17549435e-38f));
}
__device__ __forceinline__ float focal_term_grad(float x, bool hit, float gamma, float alpha) {
  const float p = 1.f / (1.f + expf(-x));
  if (hit)
    return -alpha * powf(1.f - p, gamma) *
           (1.f - p - gamma * p * logf(fmaxf(p, 1.17549435e-38f)));
  return -(1.f - alpha) * powf(p, gamma) *
         (gamma * (1.f - p) * logf(fmaxf(1.f - p, 1.17549435e-38f)) - p);
}

// grad == nullptr: sums[0] += sum of weighted losses, sums[1] += kept voxels.
// grad != nullptr: grad[v, c] = loss_weight / sums[1] * w_c * radial * d term / d x.
__global__ void __launch_bounds__(TPB_L)
focal_loss_kernel(const float* __restrict__ logits, int ld, const unsigned char* __restrict__ target,
                  const unsigned char* __restrict__ cam, long long n, int C, int ignore_index,
                  const float* __restrict__ class_w, const float* __restrict__ radial, int hw,
                  int depth, float gamma, float alpha, float loss_weight,
                  double* __restrict__ sums, float* __restrict__ grad) {
  const float scale = grad ? (float)(loss_weight / (sums[1] > 0.0 ? sums[1] : 1.0)) : 0.f;
  double acc = 0.0, cnt = 0.0;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n;
       v += (long long)gridDim.x * blockDim.x) {
    const int t = target[v];
    const bool kept = t != ignore_index && (cam == nullptr || cam[v] != 0);
    if (!kept) {
      if (grad)
        for (int c = 0; c < C; ++c) grad[v * C + c] = 0.f;
      continue;
    }
    const float rw = radial ? __ldg(radial + (v / depth) % hw) : 1.f;
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
      const float x = __ldg(logits + v * ld + c);
      const float wm = __ldg(class_w + c) * rw;                 // weight_mask[v, c]
      if (grad) grad[v * C + c] = scale * wm * focal_term_grad(x, t == c, gamma, alpha);
      else s += focal_term(x, t == c, gamma, alpha) * wm;
    }
    acc += (double)s;
    cnt += 1.0;
  }
  if (grad) return;
  acc = warp_sum(acc);
  cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0 && cnt > 0.0) {
    atomicAdd(sums, acc);
    atomicAdd(sums + 1, cnt);
  }
}

__global__ void focal_finalize_kernel(const double* __restrict__ sums, float loss_weight,
                                      float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) loss[0] = (float)(loss_weight * sums[0] / sums[1]);
}

}  // namespace

PW_API int pw_focal_loss(const float* logits, int ld, const unsigned char* target,
                         const unsigned char* camera_mask, long long n_vox, int n_cls,
                         int ignore_index, const float* class_weights, const float* radial, int hw,
                         int depth, float gamma, float alpha, float loss_weight, double* sums,
                         float* loss, void* stream) {
  PW_REQUIRE(logits && target && class_weights && sums && loss);
  PW_REQUIRE(n_vox > 0 && n_cls > 0 && ld >= n_cls && (radial == nullptr || (hw > 0 && depth > 0)));
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(sums, 0, 2 * sizeof(double), st);
  if (e != cudaSuccess) return (int)e;
  focal_loss_kernel<<<grad_blocks(n_vox), TPB_L, 0, st>>>(
      logits, ld, target, camera_mask, n_vox, n_cls, ignore_index, class_weights, radial, hw, depth,
      gamma, alpha, loss_weight, sums, nullptr);
  PW_LAUNCH_CHECK();
  focal_finalize_kernel<<<1, 32, 0, st>>>(sums, loss_weight, loss);
  PW_LAUNCH_CHECK(); pw_count_launch(2);
  return 0;
}

PW_API int pw_focal_loss_grad(const float* logits, int ld, const unsigned char* target,
                              const unsigned char* camera_mask, long long n_vox, int n_cls,
                              int ignore_index, const float* class_weights, const float* radial,
                              int hw, int depth, float gamma, float alpha, float loss_weight,
                              const double* sums, float* grad, void* stream) {
  PW_REQUIRE(logits && target && class_weights && sums && grad);
  PW_REQUIRE(n_vox > 0 && n_cls > 0 && ld >= n_cls && (radial == nullptr || (hw > 0 && depth > 0)));
  focal_loss_kernel<<<grad_blocks(n_vox), TPB_L, 0, (cudaStream_t)stream>>>(
      logits, ld, target, camera_mask, n_vox, n_cls, ignore_index, class_weights, radial, hw, depth,
      gamma, alpha, loss_weight, const_cast<double*>(sums), grad);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}
