// Per-camera constant tables for the lift and the stereo cost volume.
// Replaces the torch.inverse / matmul prologue of get_lidar_coor
// (necks/view_transformer.py:141-150) and DepthNet.gen_grid (:552-566) with
// one tiny kernel each, so the module needs no host round trip.  3x3 inverses
// are fp64 cofactor inverses rounded to fp32 and products are fp32 fma chains
// in a fixed order -- the same contract as oracle/oracle_ref.c.
#include "common.cuh"
#include "../../include/preworld_b200.h"

namespace {

__device__ void inv3_f64(const float* m, float* out) {
  double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
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

// out[r][c] = sum_k s[r*4+k] * invk[k*3+c]  (rotation part of a 4x4 times 3x3)
__device__ void rot_times(const float* s44, const float* m33, float* out) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      float acc = s44[r * 4 + 0] * m33[0 * 3 + c];
      acc = fmaf(s44[r * 4 + 1], m33[1 * 3 + c], acc);
      out[r * 3 + c] = fmaf(s44[r * 4 + 2], m33[2 * 3 + c], acc);
    }
}

__global__ void lift_cam_kernel(int n, const float* __restrict__ sensor2ego,
                                const float* __restrict__ intrin,
                                const float* __restrict__ post_rot,
                                const float* __restrict__ post_tran, float* __restrict__ cam) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* o = cam + i * PW_LIFT_CAM_FLOATS;
  float invk[9];
  inv3_f64(post_rot + i * 9, o);
  for (int k = 0; k < 3; ++k) o[9 + k] = post_tran[i * 3 + k];
  inv3_f64(intrin + i * 9, invk);
  rot_times(sensor2ego + i * 16, invk, o + 12);
  for (int r = 0; r < 3; ++r) o[21 + r] = sensor2ego[i * 16 + r * 4 + 3];
}

__global__ void cv_cam_kernel(int n, const float* __restrict__ k2s, const float* __restrict__ intrin,
                              const float* __restrict__ post_rot,
                              const float* __restrict__ post_tran, float* __restrict__ cam) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* o = cam + i * PW_CV_CAM_FLOATS;
  float invk[9];
  inv3_f64(post_rot + i * 9, o);
  for (int k = 0; k < 3; ++k) o[9 + k] = post_tran[i * 3 + k];
  inv3_f64(intrin + i * 9, invk);
  rot_times(k2s + i * 16, invk, o + 12);
  for (int r = 0; r < 3; ++r) o[21 + r] = k2s[i * 16 + r * 4 + 3];
  for (int k = 0; k < 9; ++k) o[24 + k] = intrin[i * 9 + k];
  o[33] = post_rot[i * 9 + 0]; o[34] = post_rot[i * 9 + 1];
  o[35] = post_rot[i * 9 + 3]; o[36] = post_rot[i * 9 + 4];
  o[37] = post_tran[i * 3 + 0]; o[38] = post_tran[i * 3 + 1];
  for (int k = 39; k < PW_CV_CAM_FLOATS; ++k) o[k] = 0.f;
}

}  // namespace

PW_API int pw_lift_camera_params(int n, const float* sensor2ego, const float* intrin,
                                 const float* post_rot, const float* post_tran, float* cam,
                                 void* stream) {
  PW_REQUIRE(n > 0 && sensor2ego && intrin && post_rot && post_tran && cam);
  lift_cam_kernel<<<pw_ceil_div(n, 64), 64, 0, (cudaStream_t)stream>>>(n, sensor2ego, intrin,
                                                                      post_rot, post_tran, cam);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_cv_camera_params(int n, const float* k2s_sensor, const float* intrin,
                               const float* post_rot, const float* post_tran, float* cam,
                               void* stream) {
  PW_REQUIRE(n > 0 && k2s_sensor && intrin && post_rot && post_tran && cam);
  cv_cam_kernel<<<pw_ceil_div(n, 64), 64, 0, (cudaStream_t)stream>>>(n, k2s_sensor, intrin,
                                                                    post_rot, post_tran, cam);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}
