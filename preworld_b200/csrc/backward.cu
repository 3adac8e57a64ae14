// Backward kernels of the path's custom ops: the steps
// either side of the forward that make the pre-training / fine-tuning
// gradients run on this library.
//
//   pw_bev_pool_v2_grad        <- bev_pool_v2_ext.bev_pool_v2_backward
//        (mmdet3d/ops/bev_pool_v2/src/bev_pool.cpp:74-111, kernel
//         src/bev_pool_cuda.cu:67-121; caller side bev_pool.py:43-83)
//   pw_raw2alpha_backward      <- render_utils_cuda.raw2alpha_backward
//        (nerf/cuda/render_utils_kernel.cu:507-537)
//   pw_alpha2weight_backward   <- render_utils_cuda.alpha2weight_backward
//        (nerf/cuda/render_utils_kernel.cu:654-707)
//
// All three keep the reference's arithmetic order (sequential fmaf chains,
// its float/double mixing), so they are bit-exact against oracle/oracle_ref.c.
#include "common.cuh"
#include "../../include/preworld_b200.h"

namespace {

// One warp per interval of the ranks_feat-sorted point list (the reference
// runs one THREAD per interval with two serial loops).
//   depth_grad[p_i] = sum_c out_grad[bev_i, c] * feat[feat_i, c]   lane <-> point i,
//                                                                   serial over c
//   feat_grad[feat, c] = sum_i out_grad[bev_i, c] * depth[p_i]     lane <-> channel c,
//                                                                   serial over i
__global__ void __launch_bounds__(256)
bev_pool_v2_grad_kernel(int c, int n_intervals, const float* __restrict__ out_grad,
                        const float* __restrict__ depth, const float* __restrict__ feat,
                        const int* __restrict__ ranks_depth, const int* __restrict__ ranks_feat,
                        const int* __restrict__ ranks_bev, const int* __restrict__ interval_starts,
                        const int* __restrict__ interval_lengths, float* __restrict__ depth_grad,
                        float* __restrict__ feat_grad) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long idx = warp; idx < n_intervals; idx += nwarps) {
    const int s = __ldg(interval_starts + idx);
    const int len = __ldg(interval_lengths + idx);
    for (int i = lane; i < len; i += 32) {
      const float* og = out_grad + (long long)__ldg(ranks_bev + s + i) * c;
      const float* f = feat + (long long)__ldg(ranks_feat + s + i) * c;
      float g = 0.f;
      for (int cc = 0; cc < c; ++cc) g = fmaf(__ldg(og + cc), __ldg(f + cc), g);
      depth_grad[__ldg(ranks_depth + s + i)] = g;
    }
    const long long frow = (long long)__ldg(ranks_feat + s) * c;
    for (int cc = lane; cc < c; cc += 32) {
      float g = 0.f;
      for (int i = 0; i < len; ++i)
        g = fmaf(__ldg(out_grad + (long long)__ldg(ranks_bev + s + i) * c + cc),
                 __ldg(depth + __ldg(ranks_depth + s + i)), g);
      feat_grad[frow + cc] = g;
    }
  }
}

__global__ void raw2alpha_backward_kernel(const float* __restrict__ exp_d,
                                          const float* __restrict__ grad_back, float interval,
                                          long long n, float* __restrict__ grad) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float e = exp_d[i];
    // min(float, 1e10) is a double; pow(float, float) the float overload
    const double m = fmin((double)e, 1e10);
    const float pw = powf(1 + e, -interval - 1);
    grad[i] = (float)(m * (double)pw * (double)interval * (double)grad_back[i]);
  }
}

// one thread per ray, walking its samples backwards (the scan is a serial
// dependence: back_cum)
__global__ void alpha2weight_backward_kernel(const float* __restrict__ alpha,
                                             const float* __restrict__ weight,
                                             const float* __restrict__ T,
                                             const float* __restrict__ alphainv_last,
                                             const long long* __restrict__ i_start,
                                             const long long* __restrict__ i_end, int n_rays,
                                             const float* __restrict__ grad_weights,
                                             const float* __restrict__ grad_last,
                                             float* __restrict__ grad) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  const int i_s = (int)i_start[r], i_e = (int)i_end[r];
  float back_cum = grad_last[r] * alphainv_last[r];
  for (int i = i_e - 1; i >= i_s; --i) {
    const float gw = grad_weights[i];
    const float gt = __fmul_rn(gw, T[i]);
    grad[i] = (float)((double)gt - (double)back_cum / ((double)(1 - alpha[i]) + 1e-10));
    back_cum = __fadd_rn(back_cum, __fmul_rn(gw, weight[i]));
  }
}

}  // namespace

PW_API int pw_bev_pool_v2_grad(int c, int n_intervals, const float* out_grad, const float* depth,
                               const float* feat, const int* ranks_depth, const int* ranks_feat,
                               const int* ranks_bev, const int* interval_starts,
                               const int* interval_lengths, float* depth_grad, float* feat_grad,
                               void* stream) {
  PW_REQUIRE(c > 0 && n_intervals >= 0);
  if (n_intervals == 0) return 0;
  PW_REQUIRE(out_grad && depth && feat && ranks_depth && ranks_feat && ranks_bev &&
             interval_starts && interval_lengths && depth_grad && feat_grad);
  int blocks = (int)min((long long)148 * 8, ((long long)n_intervals * 32 + 255) / 256);
  bev_pool_v2_grad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      c, n_intervals, out_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_starts,
      interval_lengths, depth_grad, feat_grad);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_raw2alpha_backward(const float* exp_d, const float* grad_back, float interval,
                                 long long n, float* grad, void* stream) {
  PW_REQUIRE(n >= 0);
  if (n == 0) return 0;
  PW_REQUIRE(exp_d && grad_back && grad);
  int blocks = (int)min((long long)148 * 8, (n + 255) / 256);
  raw2alpha_backward_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(exp_d, grad_back, interval,
                                                                      n, grad);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_alpha2weight_backward(const float* alpha, const float* weight, const float* T,
                                    const float* alphainv_last, const long long* i_start,
                                    const long long* i_end, int n_rays, const float* grad_weights,
                                    const float* grad_last, float* grad, void* stream) {
  PW_REQUIRE(n_rays >= 0);
  if (n_rays == 0) return 0;
  PW_REQUIRE(alpha && weight && T && alphainv_last && i_start && i_end && grad_weights &&
             grad_last && grad);
  alpha2weight_backward_kernel<<<pw_ceil_div(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(
      alpha, weight, T, alphainv_last, i_start, i_end, n_rays, grad_weights, grad_last, grad);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}
