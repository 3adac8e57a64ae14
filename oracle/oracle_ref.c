/*
 * TEST INFRASTRUCTURE ONLY -- the product never links or calls this file.
 *
 * Plain-C CPU restatement of the CUDA-only pieces of the reference's
 * camera->voxel path (getterupper/PreWorld @ 0b0e021).  Each function cites
 * the reference file:line it follows.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may use it.
 *
 * Parity pin: pw_ref_bev_pool_v2_{fwd,bwd} reproduce the single known-answer
 * test the reference ships (mmdet3d/ops/bev_pool_v2/bev_pool.py:145-176);
 * see tests/test_oracle.py.  The other functions have no reference KAT
 * ("parity unpinned" by the reference); they are checked against the
 * reference's own Python run verbatim in oracle/make_golden.py.
 *
 * Floating-point contract: the reference kernels are compiled by nvcc with
 * the default -fmad=true, so `acc += a * b` is one fused multiply-add.  The
 * restatement therefore uses fmaf() in the same loop order; the CUDA product
 * kernels use the same order and are compared bit-for-bit.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:21-48 (one thread per
 * (interval, channel); serial sum over the interval's points). `out` must be
 * zero-initialised by the caller (bev_pool.py:27). */
void pw_ref_bev_pool_v2_fwd(int c, int n_intervals, const float *depth,
                            const float *feat, const int *ranks_depth,
                            const int *ranks_feat, const int *ranks_bev,
                            const int *interval_starts,
                            const int *interval_lengths, float *out) {
  for (int index = 0; index < n_intervals; ++index) {
    int s = interval_starts[index], len = interval_lengths[index];
    for (int cur_c = 0; cur_c < c; ++cur_c) {
      float psum = 0.f;
      for (int i = 0; i < len; ++i)
        psum = fmaf(feat[(int64_t)ranks_feat[s + i] * c + cur_c],
                    depth[ranks_depth[s + i]], psum);
      out[(int64_t)ranks_bev[s] * c + cur_c] = psum;
    }
  }
}

/* mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:67-121 (one thread per
 * interval of the ranks_feat-sorted point list, bev_pool.py:47-57). */
void pw_ref_bev_pool_v2_bwd(int c, int n_intervals, const float *out_grad,
                            const float *depth, const float *feat,
                            const int *ranks_depth, const int *ranks_feat,
                            const int *ranks_bev, const int *interval_starts,
                            const int *interval_lengths, float *depth_grad,
                            float *feat_grad) {
  for (int idx = 0; idx < n_intervals; ++idx) {
    int s = interval_starts[idx], len = interval_lengths[idx];
    for (int i = 0; i < len; ++i) {
      const float *og = out_grad + (int64_t)ranks_bev[s + i] * c;
      const float *f = feat + (int64_t)ranks_feat[s + i] * c;
      float g = 0.f;
      for (int cc = 0; cc < c; ++cc) g = fmaf(og[cc], f[cc], g);
      depth_grad[ranks_depth[s + i]] = g;
    }
    for (int cc = 0; cc < c; ++cc) {
      float g = 0.f;
      for (int i = 0; i < len; ++i)
        g = fmaf(out_grad[(int64_t)ranks_bev[s + i] * c + cc],
                 depth[ranks_depth[s + i]], g);
      feat_grad[(int64_t)ranks_feat[s] * c + cc] = g;
    }
  }
}

/* ---- lift geometry ------------------------------------------------------
 * view_transformer.py:114-153 (get_lidar_coor) + :226-245
 * (voxel_pooling_prepare_v2 up to ranks_bev).  The reference evaluates this
 * with a chain of torch ops whose 3-term dot products have a library-defined
 * order; the restatement fixes the order (mul, fma, fma) and the 3x3 inverses
 * (fp64 cofactor inverse rounded to fp32).  oracle/make_golden.py measures the
 * number of frustum points whose voxel differs from the reference's own run
 * (they sit within a few ulp of a voxel face). */
static void inv3_f64(const float *m, float *out) {
  double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6],
         h = m[7], i = m[8];
  double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  double det = a * A + b * B + c * C;
  double r = 1.0 / det;
  out[0] = (float)(A * r);
  out[1] = (float)(-(b * i - c * h) * r);
  out[2] = (float)((b * f - c * e) * r);
  out[3] = (float)(B * r);
  out[4] = (float)((a * i - c * g) * r);
  out[5] = (float)(-(a * f - c * d) * r);
  out[6] = (float)(C * r);
  out[7] = (float)(-(a * h - b * g) * r);
  out[8] = (float)((a * e - b * d) * r);
}

static inline float dot3(const float *m, float x, float y, float z) {
  float acc = m[0] * x;
  acc = fmaf(m[1], y, acc);
  return fmaf(m[2], z, acc);
}

/* Per-camera constants: cam[24] = inv(post_rot)[9], post_tran[3],
 * combine = sensor2ego[:3,:3] @ inv(intrin) [9], sensor2ego[:3,3] [3].
 * view_transformer.py:141,148,150. */
void pw_ref_lift_camera_params(int n_cams, const float *sensor2ego /*[n,4,4]*/,
                               const float *intrin /*[n,3,3]*/,
                               const float *post_rot /*[n,3,3]*/,
                               const float *post_tran /*[n,3]*/,
                               float *cam /*[n,24]*/) {
  for (int n = 0; n < n_cams; ++n) {
    float *o = cam + n * 24;
    float invk[9];
    inv3_f64(post_rot + n * 9, o);
    memcpy(o + 9, post_tran + n * 3, 3 * sizeof(float));
    inv3_f64(intrin + n * 9, invk);
    const float *s = sensor2ego + n * 16;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        float acc = s[r * 4 + 0] * invk[0 * 3 + c];
        acc = fmaf(s[r * 4 + 1], invk[1 * 3 + c], acc);
        o[12 + r * 3 + c] = fmaf(s[r * 4 + 2], invk[2 * 3 + c], acc);
      }
    for (int r = 0; r < 3; ++r) o[21 + r] = s[r * 4 + 3];
  }
}

/* rank[p] = voxel id b*ZYX + z*YX + y*X + x of frustum point
 * p = (((b*N+n)*D+d)*H+h)*W+w, or -1 when outside the grid.
 * The (coor-lower)/interval -> .long() truncation toward zero keeps points in
 * (-1,0) voxel units in voxel 0, as view_transformer.py:226-236 does. */
void pw_ref_lift_ranks(int B, int N, int D, int H, int W, const float *xs,
                       const float *ys, const float *ds,
                       const float *cam /*[B*N,24]*/, const float *bda /*[B,9]*/,
                       const float *lower, const float *interval, int gx,
                       int gy, int gz, int *rank) {
  int64_t p = 0;
  for (int b = 0; b < B; ++b)
    for (int n = 0; n < N; ++n) {
      const float *c = cam + (b * N + n) * 24;
      const float *bd = bda + b * 9;
      for (int d = 0; d < D; ++d)
        for (int h = 0; h < H; ++h)
          for (int w = 0; w < W; ++w, ++p) {
            float px = xs[w] - c[9], py = ys[h] - c[10], pz = ds[d] - c[11];
            float qx = dot3(c + 0, px, py, pz), qy = dot3(c + 3, px, py, pz),
                  qz = dot3(c + 6, px, py, pz);
            qx = qx * qz;
            qy = qy * qz;
            float ex = dot3(c + 12, qx, qy, qz) + c[21],
                  ey = dot3(c + 15, qx, qy, qz) + c[22],
                  ez = dot3(c + 18, qx, qy, qz) + c[23];
            float fx = dot3(bd + 0, ex, ey, ez), fy = dot3(bd + 3, ex, ey, ez),
                  fz = dot3(bd + 6, ex, ey, ez);
            int64_t ix = (int64_t)((fx - lower[0]) / interval[0]);
            int64_t iy = (int64_t)((fy - lower[1]) / interval[1]);
            int64_t iz = (int64_t)((fz - lower[2]) / interval[2]);
            int ok = ix >= 0 && ix < gx && iy >= 0 && iy < gy && iz >= 0 &&
                     iz < gz;
            rank[p] = ok ? (int)(((b * (int64_t)gz + iz) * gy + iy) * gx + ix)
                         : -1;
          }
    }
}

/* mmdet3d/models/nerf/cuda/render_utils_kernel.cu:431-443 */
void pw_ref_raw2alpha(int n, const float *density, float shift, float interval,
                      float *alpha) {
  for (int i = 0; i < n; ++i) {
    float e = expf(density[i] + shift);
    alpha[i] = 1 - powf(1 + e, -interval);
  }
}

/* mmdet3d/models/nerf/cuda/render_utils_kernel.cu:577-651.  ray_id is sorted;
 * segment bounds come from its change points (:607-617, :635).  Note the
 * reference's mixed precision: `T_cum *= (1. - alpha[i])` and
 * `T_cum < 1e-3` are evaluated in double and rounded back to float. */
void pw_ref_alpha2weight(int n_pts, int n_rays, const float *alpha,
                         const int64_t *ray_id, float *weight, float *T,
                         float *alphainv_last, int64_t *i_start,
                         int64_t *i_end) {
  for (int i = 0; i < n_pts; ++i) { weight[i] = 0.f; T[i] = 1.f; }
  for (int r = 0; r < n_rays; ++r) {
    alphainv_last[r] = 1.f; i_start[r] = 0; i_end[r] = 0;
  }
  if (n_pts == 0) return;
  for (int i = 1; i < n_pts; ++i)
    if (ray_id[i] != ray_id[i - 1]) {
      i_start[ray_id[i]] = i;
      i_end[ray_id[i - 1]] = i;
    }
  i_end[ray_id[n_pts - 1]] = n_pts;
  for (int r = 0; r < n_rays; ++r) {
    int i_s = (int)i_start[r], i_e = (int)i_end[r];
    float T_cum = 1.f;
    int i;
    for (i = i_s; i < i_e; ++i) {
      T[i] = T_cum;
      weight[i] = T_cum * alpha[i];
      T_cum = (float)((double)T_cum * (1. - (double)alpha[i]));
      if ((double)T_cum < 1e-3) { i += 1; break; }
    }
    i_end[r] = i;
    alphainv_last[r] = T_cum;
  }
}

/* mmdet3d/models/nerf/cuda/render_utils_kernel.cu:507-517 (scalar_t = float):
 * `min(exp_d, 1e10)` promotes to double, `pow(1+exp_d, -interval-1)` is the
 * float overload, the product chain is evaluated left to right in double and
 * rounded once on the store. */
void pw_ref_raw2alpha_bwd(int n, const float *exp_d, const float *grad_back,
                          float interval, float *grad) {
  for (int i = 0; i < n; ++i) {
    double m = (double)exp_d[i] < 1e10 ? (double)exp_d[i] : 1e10;
    float pw = powf(1 + exp_d[i], -interval - 1);
    grad[i] = (float)(m * (double)pw * (double)interval * (double)grad_back[i]);
  }
}

/* mmdet3d/models/nerf/cuda/render_utils_kernel.cu:654-676; `grad` must be
 * zero-initialised by the caller (torch::zeros_like, :682).  back_cum is a
 * float; `1-alpha+1e-10` and the division / subtraction are double. */
void pw_ref_alpha2weight_bwd(int n_rays, const float *alpha, const float *weight,
                             const float *T, const float *alphainv_last,
                             const int64_t *i_start, const int64_t *i_end,
                             const float *grad_weights, const float *grad_last,
                             float *grad) {
  for (int r = 0; r < n_rays; ++r) {
    int i_s = (int)i_start[r], i_e = (int)i_end[r];
    float back_cum = grad_last[r] * alphainv_last[r];
    for (int i = i_e - 1; i >= i_s; --i) {
      float gt = grad_weights[i] * T[i];
      grad[i] = (float)((double)gt - (double)back_cum / ((double)(1 - alpha[i]) + 1e-10));
      back_cum += grad_weights[i] * weight[i];
    }
  }
}

/* mmdet3d/models/nerf/cuda/ub360_utils_kernel.cu:12-32 */
void pw_ref_cumdist_thres(int n_rays, int n_pts, const float *dist, float thres,
                          uint8_t *mask) {
  for (int r = 0; r < n_rays; ++r) {
    float cum = 0.f;
    for (int i = 0; i < n_pts; ++i) {
      cum += dist[(int64_t)r * n_pts + i];
      int over = cum > thres;
      cum *= (float)(!over);
      mask[(int64_t)r * n_pts + i] = (uint8_t)over;
    }
  }
}
