// Volume rendering head.
//
// (1) Drop-ins for the reference's JIT CUDA extensions on raw pointers:
//       pw_raw2alpha     <- render_utils_cuda.raw2alpha
//                           (nerf/cuda/render_utils_kernel.cu:431-443,460-481)
//       pw_alpha2weight  <- render_utils_cuda.alpha2weight
//                           (render_utils_kernel.cu:577-651; the host-side
//                           `i_end[ray_id[n-1]] = n` sync at :635 is done on
//                           the device)
//       pw_cumdist_thres <- ub360_utils_cuda.cumdist_thres
//                           (nerf/cuda/ub360_utils_kernel.cu:12-47)
// (2) pw_render_rays: the whole of NerfHead.render_one_scene +
//     render_depth/semantic/color (nerf/nerf_head.py:32-55,165-269,332-353)
//     as ONE warp-per-ray march: sample, contract, bda, cumdist mask,
//     trilinear gather (density first, the other 20 channels only for samples
//     that survive the alpha filter), transmittance scan with the reference's
//     float/double mixed rounding and early stop, per-ray sums.  None of the
//     reference's [R,417,*] intermediates (~190 MB each) is materialised.
#include "common.cuh"
#include "../../include/preworld_b200.h"

namespace {

__global__ void raw2alpha_kernel(const float* __restrict__ density, float shift, float interval,
                                 long long n, float* __restrict__ exp_d, float* __restrict__ alpha) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float e = expf(density[i] + shift);   // can be inf
    if (exp_d) exp_d[i] = e;
    alpha[i] = 1.f - powf(1.f + e, -interval);
  }
}

__global__ void a2w_init_kernel(long long n_pts, int n_rays, float* weight, float* T, float* last,
                                long long* i_start, long long* i_end) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n_pts) { weight[i] = 0.f; T[i] = 1.f; }
  if (i < n_rays) { last[i] = 1.f; i_start[i] = 0; i_end[i] = 0; }
}

__global__ void a2w_segments_kernel(const long long* __restrict__ ray_id, long long n_pts,
                                    long long* i_start, long long* i_end) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_pts) return;
  if (i > 0 && ray_id[i] != ray_id[i - 1]) {
    i_start[ray_id[i]] = i;
    i_end[ray_id[i - 1]] = i;
  }
  if (i == n_pts - 1) i_end[ray_id[i]] = n_pts;
}

__global__ void a2w_scan_kernel(const float* __restrict__ alpha, int n_rays, float* weight,
                                float* T, float* last, const long long* i_start,
                                long long* i_end) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  long long i_s = i_start[r], i_e = i_end[r];
  float T_cum = 1.f;
  long long i;
  for (i = i_s; i < i_e; ++i) {
    T[i] = T_cum;
    weight[i] = T_cum * alpha[i];
    T_cum = (float)((double)T_cum * (1. - (double)alpha[i]));
    if ((double)T_cum < 1e-3) { i += 1; break; }
  }
  i_end[r] = i;
  last[r] = T_cum;
}

__global__ void cumdist_thres_kernel(const float* __restrict__ dist, float thres, int n_rays,
                                     int n_pts, unsigned char* __restrict__ mask) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  float cum = 0.f;
  for (int i = 0; i < n_pts; ++i) {
    cum += dist[(long long)r * n_pts + i];
    bool over = cum > thres;
    cum *= (float)(!over);
    mask[(long long)r * n_pts + i] = (unsigned char)over;
  }
}

// ---------------------------------------------------------------------------
struct Sample {
  float x, y, z;   // after contraction + bda
  bool inner;
};

__device__ __forceinline__ Sample make_sample(const pw_render_desc& D, const float o[3],
                                              const float d[3], const float* bda, float t) {
  // ray_pts = rays_o + rays_d * t  (separate mul/add, as the torch ops)
  float px = __fadd_rn(o[0], __fmul_rn(d[0], t));
  float py = __fadd_rn(o[1], __fmul_rn(d[1], t));
  float pz = __fadd_rn(o[2], __fmul_rn(d[2], t));
  float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz)));
  Sample s;
  s.inner = nrm <= 1.f;
  if (!s.inner) {
    float k = __fsub_rn(__fadd_rn(1.f, D.bg_len), __fdiv_rn(D.bg_len, nrm));
    px = __fmul_rn(__fdiv_rn(px, nrm), k);
    py = __fmul_rn(__fdiv_rn(py, nrm), k);
    pz = __fmul_rn(__fdiv_rn(pz, nrm), k);
  }
  s.x = fmaf(bda[2], pz, fmaf(bda[1], py, bda[0] * px));
  s.y = fmaf(bda[5], pz, fmaf(bda[4], py, bda[3] * px));
  s.z = fmaf(bda[8], pz, fmaf(bda[7], py, bda[6] * px));
  return s;
}

struct Tri {
  int x0, y0, z0;          // voxel indices (x: X axis, ...), may be out of range
  float wx1, wy1, wz1;     // fractional parts
};

__device__ __forceinline__ Tri tri_setup(const pw_render_desc& D, const Sample& s) {
  // ind_norm = ((xyz - xyz_min) / (xyz_max - xyz_min)) * 2 - 1 ; grid_sample
  // align_corners=True: idx = ((g + 1) / 2) * (size - 1)
  float gx = __fdiv_rn(s.x - D.xyz_min[0], D.xyz_max[0] - D.xyz_min[0]) * 2.f - 1.f;
  float gy = __fdiv_rn(s.y - D.xyz_min[1], D.xyz_max[1] - D.xyz_min[1]) * 2.f - 1.f;
  float gz = __fdiv_rn(s.z - D.xyz_min[2], D.xyz_max[2] - D.xyz_min[2]) * 2.f - 1.f;
  float fx = ((gx + 1.f) * 0.5f) * (float)(D.gx - 1);
  float fy = ((gy + 1.f) * 0.5f) * (float)(D.gy - 1);
  float fz = ((gz + 1.f) * 0.5f) * (float)(D.gz - 1);
  float x0 = floorf(fx), y0 = floorf(fy), z0 = floorf(fz);
  Tri t;
  t.wx1 = fx - x0; t.wy1 = fy - y0; t.wz1 = fz - z0;
  t.x0 = (int)fminf(fmaxf(x0, -2.f), (float)D.gx + 1.f);
  t.y0 = (int)fminf(fmaxf(y0, -2.f), (float)D.gy + 1.f);
  t.z0 = (int)fminf(fmaxf(z0, -2.f), (float)D.gz + 1.f);
  return t;
}

// corner k: bit0 -> z+1 (torch "ix", the W dim), bit1 -> y+1, bit2 -> x+1
// (torch "iz", the D dim) == torch's tnw,tne,tsw,tse,bnw,bne,bsw,bse order.
__device__ __forceinline__ float tri_weight(const Tri& t, int k) {
  float wz = (k & 1) ? t.wz1 : 1.f - t.wz1;
  float wy = (k & 2) ? t.wy1 : 1.f - t.wy1;
  float wx = (k & 4) ? t.wx1 : 1.f - t.wx1;
  return wz * wy * wx;
}

__device__ __forceinline__ long long tri_voxel(const pw_render_desc& D, const Tri& t, int k) {
  int z = t.z0 + (k & 1), y = t.y0 + ((k >> 1) & 1), x = t.x0 + ((k >> 2) & 1);
  if ((unsigned)z >= (unsigned)D.gz || (unsigned)y >= (unsigned)D.gy ||
      (unsigned)x >= (unsigned)D.gx)
    return -1;
  return (long long)z * D.vs_z + (long long)y * D.vs_y + (long long)x * D.vs_x;
}

constexpr int MAX_SEM = 20;

__global__ void __launch_bounds__(128)
render_rays_kernel(const pw_render_desc D, const float* __restrict__ rays, int n_rays,
                   const float* __restrict__ tvals, int n_steps, const float* __restrict__ bda_g,
                   const float* __restrict__ density, int dld, const float* __restrict__ sem,
                   int sld, const float* __restrict__ col, int cld, float* __restrict__ o_depth,
                   float* __restrict__ o_sem, float* __restrict__ o_col, float* __restrict__ o_last,
                   unsigned char* __restrict__ o_valid) {
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rays) return;
  const float* ray = rays + (long long)r * 16;
  float gt = __ldg(ray + 2);
  bool valid = gt > 0.f && !(gt > D.max_depth);
  const int ns = D.n_sem;
  if (!valid) {
    if (lane == 0) { o_depth[r] = 0.f; o_last[r] = 0.f; o_valid[r] = 0; }
    for (int k = lane; k < ns; k += 32) o_sem[(long long)r * ns + k] = 0.f;
    if (lane < 3) o_col[(long long)r * 3 + lane] = 0.f;
    return;
  }
  float bda[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) bda[k] = __ldg(bda_g + k);
  float o[3], d[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    o[k] = __fdiv_rn(__ldg(ray + 4 + k) - D.scene_center[k], D.scene_radius[k]);
    d[k] = __ldg(ray + 7 + k);
  }
  float dn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])),
                             __fmul_rn(d[2], d[2])));
#pragma unroll
  for (int k = 0; k < 3; ++k) d[k] = __fdiv_rn(d[k], dn);

  const float dist_thres = (2.f + 2.f * D.bg_len) / (float)D.world_len * D.step_size * 0.95f;

  float acc_sem[MAX_SEM];
#pragma unroll
  for (int k = 0; k < MAX_SEM; ++k) acc_sem[k] = 0.f;
  float acc_col[3] = {0.f, 0.f, 0.f};
  float acc_depth = 0.f;
  float T_cum = 1.f;
  float cum = 0.f;          // cumdist state carried across chunks
  bool prev_over = false;   // `over` flag of sample (chunk_start - 1)
  bool stopped = false;

  for (int base = 0; base < n_steps && !stopped; base += 32) {
    const int i = base + lane;
    const bool in_range = i < n_steps;
    float t = in_range ? __ldg(tvals + i) : 0.f;
    Sample s = make_sample(D, o, d, bda, t);
    // distance to the next sample (dist[i] = |p[i+1] - p[i]|)
    float dist = 0.f;
    if (i + 1 < n_steps) {
      Sample s1 = make_sample(D, o, d, bda, __ldg(tvals + i + 1));
      float ex = s1.x - s.x, ey = s1.y - s.y, ez = s1.z - s.z;
      dist = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez)));
    }
    // sequential cumdist over this chunk (all lanes redundantly)
    unsigned over_bits = 0;
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
      float dj = __shfl_sync(0xffffffffu, dist, j);
      if (base + j + 1 < n_steps) {
        cum += dj;
        bool over = cum > dist_thres;
        if (over) { cum = 0.f; over_bits |= 1u << j; }
      }
    }
    // mask[i] = inner[i] | over[i-1]
    bool over_prev = lane == 0 ? prev_over : ((over_bits >> (lane - 1)) & 1u);
    prev_over = (over_bits >> 31) & 1u;
    bool keep = in_range && (s.inner || over_prev);

    // density + alpha
    Tri tr;
    float alpha = 0.f;
    if (keep) {
      tr = tri_setup(D, s);
      float dens = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        long long v = tri_voxel(D, tr, k);
        if (v >= 0) dens = fmaf(__ldg(density + v * dld), tri_weight(tr, k), dens);
      }
      float e = expf(dens + D.act_shift);
      alpha = 1.f - powf(1.f + e, -D.interval);
    }
    bool act = keep && alpha > D.fast_color_thres;
    unsigned act_bits = __ballot_sync(0xffffffffu, act);
    // transmittance scan in sample order (render_utils_kernel.cu:577-605)
    float w_mine = 0.f;
    while (act_bits) {
      int j = __ffs(act_bits) - 1;
      act_bits &= act_bits - 1;
      float aj = __shfl_sync(0xffffffffu, alpha, j);
      float wj = T_cum * aj;
      if (lane == j) w_mine = wj;
      T_cum = (float)((double)T_cum * (1. - (double)aj));
      if ((double)T_cum < 1e-3) { stopped = true; break; }
    }
    if (w_mine > D.fast_color_thres) {
      float sdepth = 1.f - 1.f / (1.f + t);
      acc_depth = fmaf(w_mine, sdepth, acc_depth);
      float sv[MAX_SEM];
#pragma unroll
      for (int c = 0; c < MAX_SEM; ++c) sv[c] = 0.f;
      float cv[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        long long v = tri_voxel(D, tr, k);
        if (v < 0) continue;
        float wk = tri_weight(tr, k);
        const float* sp = sem + v * sld;
#pragma unroll
        for (int c = 0; c < MAX_SEM; ++c)
          if (c < ns) sv[c] = fmaf(__ldg(sp + c), wk, sv[c]);
        const float* cp = col + v * cld;
#pragma unroll
        for (int c = 0; c < 3; ++c) cv[c] = fmaf(__ldg(cp + c), wk, cv[c]);
      }
#pragma unroll
      for (int c = 0; c < MAX_SEM; ++c)
        if (c < ns) acc_sem[c] = fmaf(w_mine, sv[c], acc_sem[c]);
#pragma unroll
      for (int c = 0; c < 3; ++c) acc_col[c] = fmaf(w_mine, cv[c], acc_col[c]);
    }
  }

  // per-ray sums
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    acc_depth += __shfl_xor_sync(0xffffffffu, acc_depth, off);
#pragma unroll
    for (int c = 0; c < 3; ++c) acc_col[c] += __shfl_xor_sync(0xffffffffu, acc_col[c], off);
#pragma unroll
    for (int c = 0; c < MAX_SEM; ++c)
      if (c < ns) acc_sem[c] += __shfl_xor_sync(0xffffffffu, acc_sem[c], off);
  }
  if (lane == 0) {
    o_depth[r] = (acc_depth + 1e-7f) * D.radius;
    o_last[r] = T_cum;
    o_valid[r] = 1;
#pragma unroll
    for (int c = 0; c < MAX_SEM; ++c)
      if (c < ns) o_sem[(long long)r * ns + c] = acc_sem[c];
#pragma unroll
    for (int c = 0; c < 3; ++c) o_col[(long long)r * 3 + c] = acc_col[c];
  }
}

}  // namespace

PW_API int pw_raw2alpha(const float* density, float shift, float interval, long long n,
                        float* exp_d, float* alpha, void* stream) {
  PW_REQUIRE(n >= 0);
  if (n == 0) return 0;
  PW_REQUIRE(density && alpha);
  int blocks = (int)min((long long)148 * 16, (n + 255) / 256);
  raw2alpha_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(density, shift, interval, n, exp_d,
                                                             alpha);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_alpha2weight(const float* alpha, const long long* ray_id, long long n_pts,
                           int n_rays, float* weight, float* T, float* alphainv_last,
                           long long* i_start, long long* i_end, void* stream) {
  PW_REQUIRE(n_pts >= 0 && n_rays >= 0 && alphainv_last && i_start && i_end);
  cudaStream_t st = (cudaStream_t)stream;
  long long m = n_pts > n_rays ? n_pts : n_rays;
  if (m == 0) return 0;
  PW_REQUIRE(n_pts == 0 || (alpha && ray_id && weight && T));
  a2w_init_kernel<<<pw_ceil_div(m, 256), 256, 0, st>>>(n_pts, n_rays, weight, T, alphainv_last,
                                                       i_start, i_end);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  if (n_pts == 0 || n_rays == 0) return 0;
  a2w_segments_kernel<<<pw_ceil_div(n_pts, 256), 256, 0, st>>>(ray_id, n_pts, i_start, i_end);
  PW_LAUNCH_CHECK();
  a2w_scan_kernel<<<pw_ceil_div(n_rays, 128), 128, 0, st>>>(alpha, n_rays, weight, T,
                                                            alphainv_last, i_start, i_end);
  PW_LAUNCH_CHECK(); pw_count_launch(2);
  return 0;
}

PW_API int pw_cumdist_thres(const float* dist, float thres, int n_rays, int n_pts,
                            unsigned char* mask, void* stream) {
  PW_REQUIRE(n_rays >= 0 && n_pts >= 0);
  if (n_rays == 0 || n_pts == 0) return 0;
  PW_REQUIRE(dist && mask);
  cumdist_thres_kernel<<<pw_ceil_div(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(
      dist, thres, n_rays, n_pts, mask);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_render_rays(const pw_render_desc* desc, const float* rays, int n_rays,
                          const float* t_vals, int n_steps, const float* bda,
                          const float* density, int density_ld, const float* semantic, int sem_ld,
                          const float* color, int col_ld, float* out_depth, float* out_sem,
                          float* out_col, float* out_last, unsigned char* out_valid,
                          void* stream) {
  PW_REQUIRE(desc && n_rays >= 0);
  if (n_rays == 0) return 0;
  PW_REQUIRE(rays && t_vals && bda && density && semantic && color && out_depth && out_sem &&
             out_col && out_last && out_valid);
  PW_REQUIRE(desc->n_sem > 0 && desc->n_sem <= MAX_SEM && n_steps > 0);
  int blocks = pw_ceil_div((long long)n_rays * 32, 128);
  render_rays_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(
      *desc, rays, n_rays, t_vals, n_steps, bda, density, density_ld, semantic, sem_ld, color,
      col_ld, out_depth, out_sem, out_col, out_last, out_valid);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

// ---------------------------------------------------------------------------
// Reduction of the renderings to NerfHead.compute_loss's sums
// (nerf/nerf_head.py:271-291 with silog_loss / l1_loss, nerf/utils.py:71-87, and
// nn.CrossEntropyLoss(weight, reduction='mean')): one thread per ray, masked rays
// (0 < depth <= 52, the kernel's `valid`) only, fp64 accumulation.
//   sums[0] n   [1] sum d   [2] sum d^2        d = log(depth + 1e-7) - log(gt_depth)
//   sums[3] sum w[t] * nll  [4] sum w[t]       nll = logsumexp(sem) - sem[t]
//   sums[5..7] sum |color - gt_color| per channel
//   sums[8] sum p log p + (1 - p) log(1 - p),  p = clamp(alphainv_last, 1e-6, 1 - 1e-6)
namespace {
constexpr int RL_SUMS = 9;

__global__ void render_loss_kernel(const float* __restrict__ rays, int n_rays, int n_sem,
                                   const float* __restrict__ depth, const float* __restrict__ sem,
                                   const float* __restrict__ col, const float* __restrict__ last,
                                   const unsigned char* __restrict__ valid,
                                   const float* __restrict__ class_w, double* __restrict__ sums) {
  double v[RL_SUMS];
#pragma unroll
  for (int k = 0; k < RL_SUMS; ++k) v[k] = 0.0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rays; r += gridDim.x * blockDim.x) {
    if (!valid[r]) continue;
    const float* ray = rays + (size_t)r * 16;
    const float d = logf(depth[r] + 1e-7f) - logf(ray[2]);
    v[0] += 1.0; v[1] += d; v[2] += (double)d * d;
    const int t = (int)ray[3];                            // target_semantic.long()
    const float* s = sem + (size_t)r * n_sem;
    float mx = s[0];
    for (int c = 1; c < n_sem; ++c) mx = fmaxf(mx, s[c]);
    float se = 0.f;
    for (int c = 0; c < n_sem; ++c) se += expf(s[c] - mx);
    if (t >= 0 && t < n_sem) {
      const float w = class_w[t];
      v[3] += (double)(w * (logf(se) + mx - s[t]));
      v[4] += w;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) v[5 + c] += fabsf(col[(size_t)r * 3 + c] - ray[13 + c]);
    const float p = fminf(fmaxf(last[r], 1e-6f), 1.f - 1e-6f);
    v[8] += (double)(p * logf(p) + (1.f - p) * logf(1.f - p));
  }
  __shared__ double red[RL_SUMS][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < RL_SUMS; ++k) {
    double x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) red[k][warp] = x;
  }
  __syncthreads();
  if (threadIdx.x < RL_SUMS) {
    double x = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) x += red[threadIdx.x][w];
    atomicAdd(sums + threadIdx.x, x);
  }
}
}  // namespace

PW_API int pw_render_loss_sums(const float* rays, int n_rays, int n_sem, const float* depth,
                               const float* sem, const float* col, const float* last,
                               const unsigned char* valid, const float* class_weights,
                               double* sums, void* stream) {
  PW_REQUIRE(n_rays >= 0 && n_sem > 0 && sums);
  cudaError_t e = cudaMemsetAsync(sums, 0, RL_SUMS * sizeof(double), (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  if (n_rays == 0) return 0;
  PW_REQUIRE(rays && depth && sem && col && last && valid && class_weights);
  const int blocks = min(pw_ceil_div(n_rays, 256), 148 * 4);
  render_loss_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(rays, n_rays, n_sem, depth, sem,
                                                              col, last, valid, class_weights, sums);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}
