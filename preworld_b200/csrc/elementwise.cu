// Channels-last element-wise / reduction helpers of the image side and the
// 3-D encoder.  All kernels are HBM/L2-bound streaming kernels: float4
// accesses along the channel dimension, grid-stride loops.
#include "common.cuh"
#include "../../include/preworld_b200.h"

namespace {

constexpr int TPB = 256;

inline int grid_for(long long work, int per_block = TPB, int max_blocks = 148 * 16) {
  long long b = (work + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

// ---- NCHW -> NHWC (padded channels) --------------------------------------
__global__ void nchw_to_nhwc_pad_kernel(const float* __restrict__ x, float* __restrict__ y,
                                        int n, int c, long long hw, int c_pad,
                                        long long img_stride) {
  long long total = (long long)n * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long img = i / hw, p = i - img * hw;
    const float* src = x + img * img_stride + p;
    float* dst = y + i * c_pad;
    for (int ch = 0; ch < c_pad; ++ch) dst[ch] = ch < c ? __ldg(src + ch * hw) : 0.f;
  }
}

// ---- NCHW -> space-to-depth(2) NHWC ----------------------------------------------
// y[n, Y, X, (dy*2+dx)*4 + c] = x[n, c, 2Y+dy, 2X+dx]  (c < 3.. <=4), remaining
// channels up to c_pad zero.  Turns the stride-2 7x7 ResNet stem into a
// stride-1 4x4 conv with a 32-multiple Cin, i.e. a tensor-core conv.
__global__ void nchw_to_s2d_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int n,
                                        int c, int h, int w, int c_pad, long long img_stride) {
  const int oh = h >> 1, ow = w >> 1;
  const int q4 = c_pad >> 2;                            // float4 chunks per output pixel
  long long total = (long long)n * oh * ow * q4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % q4);
    long long t = i / q4;
    const int X = (int)(t % ow); t /= ow;
    const int Y = (int)(t % oh);
    const long long img = t / oh;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (q < 4) {                                        // chunk q = (dy, dx), lanes = channel
      const int dy = q >> 1, dx = q & 1;
      const float* src = x + img * img_stride + (long long)(2 * Y + dy) * w + 2 * X + dx;
      for (int ch = 0; ch < c && ch < 4; ++ch) v[ch] = __ldg(src + (long long)ch * h * w);
    }
    *reinterpret_cast<float4*>(y + (((img * oh + Y) * ow + X) * (long long)c_pad) + q * 4) =
        make_float4(v[0], v[1], v[2], v[3]);
  }
}

// ---- NHWC slice -> NCHW ----------------------------------------------------
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, int x_ld,
                                    float* __restrict__ y, int n, int c, long long hw) {
  __shared__ float tile[32][33];
  // blockIdx.x: pixel tile, blockIdx.y: channel tile, blockIdx.z: image
  long long p0 = (long long)blockIdx.x * 32;
  int c0 = blockIdx.y * 32;
  int img = blockIdx.z;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    long long p = p0 + r;
    int ch = c0 + tx;
    tile[r][tx] = (p < hw && ch < c) ? __ldg(x + ((long long)img * hw + p) * x_ld + ch) : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int ch = c0 + r;
    long long p = p0 + tx;
    if (p < hw && ch < c) y[((long long)img * c + ch) * hw + p] = tile[tx][r];
  }
}

// ---- 3x3 stride-2 pad-1 max pool ------------------------------------------
__global__ void maxpool3x3s2_kernel(const float* __restrict__ x, float* __restrict__ y, int n,
                                    int h, int w, int c4, int oh, int ow) {
  long long total = (long long)n * oh * ow * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int ch = (int)(i % c4);
    long long t = i / c4;
    int ox = (int)(t % ow); t /= ow;
    int oy = (int)(t % oh);
    int img = (int)(t / oh);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int dy = 0; dy < 3; ++dy) {
      int iy = oy * 2 - 1 + dy;
      if ((unsigned)iy >= (unsigned)h) continue;
      for (int dx = 0; dx < 3; ++dx) {
        int ix = ox * 2 - 1 + dx;
        if ((unsigned)ix >= (unsigned)w) continue;
        float4 v = pw_ldg4(x + ((((long long)img * h + iy) * w + ix) * c4 + ch) * 4);
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y);
        m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    }
    reinterpret_cast<float4*>(y)[i] = m;
  }
}

// ---- nearest upsample + add -------------------------------------------------
__global__ void upsample_nearest_add_kernel(const float* __restrict__ x, float* __restrict__ y,
                                            int n, int h, int w, int oh, int ow, int c4) {
  long long total = (long long)n * oh * ow * c4;
  // torch nearest: src = min(floor(dst * (in/out)), in-1), scale in float
  const float sy = (float)h / (float)oh, sx = (float)w / (float)ow;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int ch = (int)(i % c4);
    long long t = i / c4;
    int ox = (int)(t % ow); t /= ow;
    int oy = (int)(t % oh);
    int img = (int)(t / oh);
    int iy = min((int)floorf(oy * sy), h - 1);
    int ix = min((int)floorf(ox * sx), w - 1);
    float4 v = pw_ldg4(x + ((((long long)img * h + iy) * w + ix) * c4 + ch) * 4);
    float4 o = reinterpret_cast<float4*>(y)[i];
    o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
    reinterpret_cast<float4*>(y)[i] = o;
  }
}

// ---- per-(image, channel) gate ---------------------------------------------
__global__ void scale_channels_kernel(const float* __restrict__ x, int x_ld,
                                      const float* __restrict__ gate, float* __restrict__ y,
                                      int y_ld, int n, long long pixels, int c4) {
  long long total = (long long)n * pixels * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int ch = (int)(i % c4);
    long long row = i / c4;
    int img = (int)(row / pixels);
    float4 v = pw_ldg4(x + row * x_ld + ch * 4);
    float4 g = pw_ldg4(gate + (long long)img * c4 * 4 + ch * 4);
    v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
    *reinterpret_cast<float4*>(y + row * y_ld + ch * 4) = v;
  }
}

// ---- global average pool ----------------------------------------------------
// one block per (image, 32-channel group): 8 pixel lanes x 32 channels
__global__ void global_avgpool_kernel(const float* __restrict__ x, int x_ld, float* __restrict__ y,
                                      int n, long long pixels, int c) {
  __shared__ float part[8][32];
  int img = blockIdx.y;
  int ch = blockIdx.x * 32 + (threadIdx.x & 31);
  int lane_p = threadIdx.x >> 5;
  float s = 0.f;
  if (ch < c)
    for (long long p = lane_p; p < pixels; p += 8)
      s += __ldg(x + ((long long)img * pixels + p) * x_ld + ch);
  part[lane_p][threadIdx.x & 31] = s;
  __syncthreads();
  if (lane_p == 0 && ch < c) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x & 31];
    y[(long long)img * c + ch] = t / (float)pixels;
  }
}

__global__ void broadcast_channels_kernel(const float* __restrict__ v, float* __restrict__ y,
                                          int y_ld, int n, long long pixels, int c) {
  long long total = (long long)n * pixels * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int ch = (int)(i % c);
    long long row = i / c;
    int img = (int)(row / pixels);
    y[row * y_ld + ch] = __ldg(v + (long long)img * c + ch);
  }
}

// ---- softmax over depth bins ------------------------------------------------
__global__ void softmax_depth_kernel(const float* __restrict__ logits, int in_ld,
                                     float* __restrict__ prob_cl, float* __restrict__ prob_pl,
                                     int n, long long pixels, int d) {
  long long rows = (long long)n * pixels;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows;
       r += (long long)gridDim.x * blockDim.x) {
    const float* src = logits + r * in_ld;
    float m = -INFINITY;
    for (int k = 0; k < d; ++k) m = fmaxf(m, __ldg(src + k));
    float s = 0.f;
    for (int k = 0; k < d; ++k) s += expf(__ldg(src + k) - m);
    long long img = r / pixels, p = r - img * pixels;
    for (int k = 0; k < d; ++k) {
      float v = expf(__ldg(src + k) - m) / s;
      if (prob_cl) prob_cl[r * d + k] = v;
      if (prob_pl) prob_pl[(img * d + k) * pixels + p] = v;
    }
  }
}

// ---- trilinear upsample (align_corners=True) into a channel slice ----------
__global__ void upsample_trilinear_kernel(const float* __restrict__ x, int x_ld,
                                          float* __restrict__ y, int y_ld, int b, int iz, int iy,
                                          int ix, int c4, int oz, int oy, int ox) {
  long long total = (long long)b * oz * oy * ox * c4;
  // torch area_pixel_compute_scale(align_corners=True): (in-1)/(out-1)
  const float sz = oz > 1 ? (float)(iz - 1) / (float)(oz - 1) : 0.f;
  const float sy = oy > 1 ? (float)(iy - 1) / (float)(oy - 1) : 0.f;
  const float sx = ox > 1 ? (float)(ix - 1) / (float)(ox - 1) : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int ch = (int)(i % c4);
    long long t = i / c4;
    int x_o = (int)(t % ox); t /= ox;
    int y_o = (int)(t % oy); t /= oy;
    int z_o = (int)(t % oz);
    int bb = (int)(t / oz);
    float fz = sz * z_o, fy = sy * y_o, fx = sx * x_o;
    int z0 = (int)fz, y0 = (int)fy, x0 = (int)fx;
    int z1 = z0 + (z0 < iz - 1), y1 = y0 + (y0 < iy - 1), x1 = x0 + (x0 < ix - 1);
    float lz1 = fz - z0, ly1 = fy - y0, lx1 = fx - x0;
    float lz0 = 1.f - lz1, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    auto at = [&](int zz, int yy, int xx) {
      return pw_ldg4(x + ((((long long)bb * iz + zz) * iy + yy) * ix + xx) * x_ld + ch * 4);
    };
    float4 v000 = at(z0, y0, x0), v001 = at(z0, y0, x1), v010 = at(z0, y1, x0),
           v011 = at(z0, y1, x1), v100 = at(z1, y0, x0), v101 = at(z1, y0, x1),
           v110 = at(z1, y1, x0), v111 = at(z1, y1, x1);
    // same association as ATen's upsample_trilinear3d (t0*(h0*(w0*a+w1*b)+h1*(..))+t1*(..))
#define TRI(f)                                                                       \
    (lz0 * (ly0 * (lx0 * v000.f + lx1 * v001.f) + ly1 * (lx0 * v010.f + lx1 * v011.f)) + \
     lz1 * (ly0 * (lx0 * v100.f + lx1 * v101.f) + ly1 * (lx0 * v110.f + lx1 * v111.f)))
    float4 o = make_float4(TRI(x), TRI(y), TRI(z), TRI(w));
#undef TRI
    *reinterpret_cast<float4*>(
        y + ((((long long)bb * oz + z_o) * oy + y_o) * ox + x_o) * y_ld + ch * 4) = o;
  }
}

// y = up(x1) + up(x2): both inputs trilinearly upsampled (align_corners=True)
// to the output grid in one pass (LSSFPN3D after commuting its 1x1x1 conv with
// the interpolation; necks/lss_fpn.py:139-146).
__device__ __forceinline__ float4 trilerp4(const float* __restrict__ x, int x_ld, int bb, int iz,
                                           int iy, int ix, int oz, int oy, int ox, int z_o,
                                           int y_o, int x_o, int ch) {
  const float sz = oz > 1 ? (float)(iz - 1) / (float)(oz - 1) : 0.f;
  const float sy = oy > 1 ? (float)(iy - 1) / (float)(oy - 1) : 0.f;
  const float sx = ox > 1 ? (float)(ix - 1) / (float)(ox - 1) : 0.f;
  float fz = sz * z_o, fy = sy * y_o, fx = sx * x_o;
  int z0 = (int)fz, y0 = (int)fy, x0 = (int)fx;
  int z1 = z0 + (z0 < iz - 1), y1 = y0 + (y0 < iy - 1), x1 = x0 + (x0 < ix - 1);
  float lz1 = fz - z0, ly1 = fy - y0, lx1 = fx - x0;
  float lz0 = 1.f - lz1, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  auto at = [&](int zz, int yy, int xx) {
    return pw_ldg4(x + ((((long long)bb * iz + zz) * iy + yy) * ix + xx) * x_ld + ch * 4);
  };
  float4 v000 = at(z0, y0, x0), v001 = at(z0, y0, x1), v010 = at(z0, y1, x0),
         v011 = at(z0, y1, x1), v100 = at(z1, y0, x0), v101 = at(z1, y0, x1),
         v110 = at(z1, y1, x0), v111 = at(z1, y1, x1);
#define TRI(f)                                                                       \
  (lz0 * (ly0 * (lx0 * v000.f + lx1 * v001.f) + ly1 * (lx0 * v010.f + lx1 * v011.f)) + \
   lz1 * (ly0 * (lx0 * v100.f + lx1 * v101.f) + ly1 * (lx0 * v110.f + lx1 * v111.f)))
  float4 o = make_float4(TRI(x), TRI(y), TRI(z), TRI(w));
#undef TRI
  return o;
}

__global__ void upsample_trilinear2_kernel(const float* __restrict__ x1, int x1_ld, int z1, int y1,
                                           int w1, const float* __restrict__ x2, int x2_ld, int z2,
                                           int y2, int w2, float* __restrict__ y, int y_ld, int b,
                                           int c4, int oz, int oy, int ox) {
  // total < 2^31 (checked by the caller): 32-bit index arithmetic (four 64-bit divisions
  // per thread were a large part of this kernel's instructions)
  const unsigned total = (unsigned)b * oz * oy * ox * c4;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int ch = (int)(i % (unsigned)c4);
    unsigned t = i / (unsigned)c4;
    int x_o = (int)(t % (unsigned)ox); t /= (unsigned)ox;
    int y_o = (int)(t % (unsigned)oy); t /= (unsigned)oy;
    int z_o = (int)(t % (unsigned)oz);
    int bb = (int)(t / (unsigned)oz);
    const float4 a = trilerp4(x1, x1_ld, bb, z1, y1, w1, oz, oy, ox, z_o, y_o, x_o, ch);
    const float4 c = trilerp4(x2, x2_ld, bb, z2, y2, w2, oz, oy, ox, z_o, y_o, x_o, ch);
    __stcs(reinterpret_cast<float4*>(
               y + ((((long long)bb * oz + z_o) * oy + y_o) * ox + x_o) * y_ld + ch * 4),
           make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w));
  }
}

__global__ void copy_channels_kernel(const float* __restrict__ x, int x_ld, float* __restrict__ y,
                                     int y_ld, long long pixels, int c4) {
  long long total = pixels * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int ch = (int)(i % c4);
    long long row = i / c4;
    *reinterpret_cast<float4*>(y + row * y_ld + ch * 4) = pw_ldg4(x + row * x_ld + ch * 4);
  }
}

// ---- occupancy argmax, [Z,Y,X] voxel order -> [X,Y,Z] uint8 grid -----------
__global__ void argmax_zyx_to_xyz_kernel(const float* __restrict__ logits, int ld, int ncls,
                                         unsigned char* __restrict__ occ,
                                         unsigned char* __restrict__ geo, int free_idx,
                                         int geo_value, int gx, int gy, int gz) {
  long long total = (long long)gx * gy * gz;
  // iterate in OUTPUT order (x,y,z with z fastest) so byte stores coalesce
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int z = (int)(i % gz);
    long long t = i / gz;
    int yv = (int)(t % gy);
    int xv = (int)(t / gy);
    const float* src = logits + (((long long)z * gy + yv) * gx + xv) * ld;
    float best = __ldg(src);
    int arg = 0;
    for (int k = 1; k < ncls; ++k) {
      float v = __ldg(src + k);
      if (v > best) { best = v; arg = k; }
    }
    occ[i] = (unsigned char)arg;
    if (geo) geo[i] = (unsigned char)(arg != free_idx ? 0 : geo_value);
  }
}

__global__ void density_occ_kernel(const float* __restrict__ density, int dld,
                                   const float* __restrict__ sem, int ld, int ncls, float thr,
                                   int empty_idx, unsigned char* __restrict__ occ,
                                   unsigned char* __restrict__ geo, int gx, int gy, int gz) {
  long long total = (long long)gx * gy * gz;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int z = (int)(i % gz);
    long long t = i / gz;
    int yv = (int)(t % gy);
    int xv = (int)(t / gy);
    long long v = ((long long)z * gy + yv) * gx + xv;
    bool nonempty = __ldg(density + v * dld) > thr;
    int arg = empty_idx;
    if (nonempty) {
      const float* src = sem + v * ld;
      float best = __ldg(src);
      arg = 0;
      for (int k = 1; k < ncls; ++k) {
        float q = __ldg(src + k);
        if (q > best) { best = q; arg = k; }
      }
    }
    occ[i] = (unsigned char)arg;
    if (geo) geo[i] = nonempty ? 0 : (unsigned char)empty_idx;
  }
}

__global__ void zyx_to_xyz_kernel(const float* __restrict__ x, float* __restrict__ y, int b,
                                  int gz, int gy, int gx, int c4) {
  long long total = (long long)b * gz * gy * gx * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int ch = (int)(i % c4);
    long long t = i / c4;
    int z = (int)(t % gz); t /= gz;
    int yv = (int)(t % gy); t /= gy;
    int xv = (int)(t % gx);
    int bb = (int)(t / gx);
    reinterpret_cast<float4*>(y)[i] =
        pw_ldg4(x + (((((long long)bb * gz + z) * gy + yv) * gx + xv) * c4 + ch) * 4);
  }
}

}  // namespace

static long long g_launches = 0;
void pw_count_launch(int k) { __atomic_add_fetch(&g_launches, k, __ATOMIC_RELAXED); }
PW_API long long pw_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
PW_API int pw_abi_version(void) { return PW_ABI_VERSION; }

#define ST ((cudaStream_t)stream)

PW_API int pw_nchw_to_nhwc_pad(const float* x, long long img_stride, float* y, int n, int c,
                               int h, int w, int c_pad, void* stream) {
  PW_REQUIRE(x && y && n > 0 && c > 0 && c_pad >= c);
  long long hw = (long long)h * w;
  PW_REQUIRE(img_stride >= c * hw);
  nchw_to_nhwc_pad_kernel<<<grid_for(n * hw), TPB, 0, ST>>>(x, y, n, c, hw, c_pad, img_stride);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_nchw_to_s2d_nhwc(const float* x, long long img_stride, float* y, int n, int c,
                               int h, int w, int c_pad, void* stream) {
  PW_REQUIRE(x && y && n > 0 && c > 0 && c <= 4 && c_pad >= 16 && (c_pad & 3) == 0);
  PW_REQUIRE((h & 1) == 0 && (w & 1) == 0 && img_stride >= (long long)c * h * w);
  nchw_to_s2d_nhwc_kernel<<<grid_for((long long)n * (h / 2) * (w / 2) * (c_pad / 4)), TPB, 0, ST>>>(
      x, y, n, c, h, w, c_pad, img_stride);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_nhwc_to_nchw(const float* x, int x_ld, float* y, int n, int c, long long pixels,
                           void* stream) {
  PW_REQUIRE(x && y && n > 0 && c > 0 && pixels > 0 && x_ld >= c);
  dim3 grid(pw_ceil_div(pixels, 32), pw_ceil_div(c, 32), n);
  nhwc_to_nchw_kernel<<<grid, 256, 0, ST>>>(x, x_ld, y, n, c, pixels);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_maxpool3x3s2(const float* x, float* y, int n, int h, int w, int c, int oh, int ow,
                           void* stream) {
  PW_REQUIRE(x && y && (c & 3) == 0 && oh == (h + 2 - 3) / 2 + 1 && ow == (w + 2 - 3) / 2 + 1);
  maxpool3x3s2_kernel<<<grid_for((long long)n * oh * ow * (c / 4)), TPB, 0, ST>>>(
      x, y, n, h, w, c / 4, oh, ow);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_upsample_nearest_add(const float* x, float* y, int n, int h, int w, int oh, int ow,
                                   int c, void* stream) {
  PW_REQUIRE(x && y && (c & 3) == 0);
  upsample_nearest_add_kernel<<<grid_for((long long)n * oh * ow * (c / 4)), TPB, 0, ST>>>(
      x, y, n, h, w, oh, ow, c / 4);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_scale_channels(const float* x, int x_ld, const float* gate, float* y, int y_ld,
                             int n, long long pixels, int c, void* stream) {
  PW_REQUIRE(x && y && gate && (c & 3) == 0 && (x_ld & 3) == 0 && (y_ld & 3) == 0);
  scale_channels_kernel<<<grid_for(n * pixels * (c / 4)), TPB, 0, ST>>>(x, x_ld, gate, y, y_ld, n,
                                                                      pixels, c / 4);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_global_avgpool(const float* x, int x_ld, float* y, int n, long long pixels, int c,
                             void* stream) {
  PW_REQUIRE(x && y && n > 0 && pixels > 0);
  dim3 grid(pw_ceil_div(c, 32), n);
  global_avgpool_kernel<<<grid, 256, 0, ST>>>(x, x_ld, y, n, pixels, c);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_broadcast_channels(const float* v, float* y, int y_ld, int n, long long pixels,
                                 int c, void* stream) {
  PW_REQUIRE(v && y && y_ld >= c);
  broadcast_channels_kernel<<<grid_for(n * pixels * c), TPB, 0, ST>>>(v, y, y_ld, n, pixels, c);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_softmax_depth(const float* logits, int in_ld, float* prob_cl, float* prob_planar,
                            int n, long long pixels, int d, void* stream) {
  PW_REQUIRE(logits && in_ld >= d && (prob_cl || prob_planar));
  softmax_depth_kernel<<<grid_for(n * pixels, 128), 128, 0, ST>>>(logits, in_ld, prob_cl,
                                                                   prob_planar, n, pixels, d);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_upsample_trilinear(const float* x, int x_ld, float* y, int y_ld, int b, int z,
                                 int yy, int xx, int c, int oz, int oy, int ox, void* stream) {
  PW_REQUIRE(x && y && (c & 3) == 0 && (x_ld & 3) == 0 && (y_ld & 3) == 0);
  upsample_trilinear_kernel<<<grid_for((long long)b * oz * oy * ox * (c / 4)), TPB, 0, ST>>>(
      x, x_ld, y, y_ld, b, z, yy, xx, c / 4, oz, oy, ox);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_upsample_trilinear2(const float* x1, int x1_ld, int z1, int y1, int w1,
                                  const float* x2, int x2_ld, int z2, int y2, int w2, float* y,
                                  int y_ld, int b, int c, int oz, int oy, int ox, void* stream) {
  PW_REQUIRE(x1 && x2 && y && (c & 3) == 0 && (x1_ld & 3) == 0 && (x2_ld & 3) == 0 &&
             (y_ld & 3) == 0);
  PW_REQUIRE((long long)b * oz * oy * ox * (c / 4) < (1ll << 31));
  upsample_trilinear2_kernel<<<grid_for((long long)b * oz * oy * ox * (c / 4)), TPB, 0, ST>>>(
      x1, x1_ld, z1, y1, w1, x2, x2_ld, z2, y2, w2, y, y_ld, b, c / 4, oz, oy, ox);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

// ---- chains of small dense layers on a handful of row vectors ------------------
// DepthNet's camera-parameter branch (view_transformer.py:421-470,606-617): per image,
// Mlp(27 -> mid -> mid) and SELayer's reduce / expand on a [mid] vector, for the
// context and the depth gate -- eight 12-row GEMMs as separate launches were 0.13 ms
// of launch latency.  One CTA per (row, chain): the activations stay in shared
// memory, thread c owns output channel c (weights [cin, w_ld]: coalesced over c).
namespace {
struct ChainPack {
  pw_dense_chain c[PW_MAX_CHAINS];
};
constexpr int CHAIN_MAX_WIDTH = 1024;
constexpr int CHAIN_WARPS = 8;

__global__ void __launch_bounds__(CHAIN_WARPS * 32)
dense_chains_kernel(const ChainPack pack, const float* __restrict__ x, int x_ld, int rows) {
  // K is split over the 8 warps (a thread that walks all of K alone waits one L2 round
  // trip per weight: the chain was latency bound), lane = channel within a group of 32:
  // every weight load is one coalesced 128-byte row segment, the groups of a warp are
  // independent accumulation chains; the 8 partial sums are added in warp order.
  __shared__ float buf[2][CHAIN_MAX_WIDTH];
  __shared__ float part[CHAIN_WARPS][CHAIN_MAX_WIDTH];
  const pw_dense_chain& ch = pack.c[blockIdx.y];
  const int row = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cin0 = ch.layer[0].cin;
  for (int k = threadIdx.x; k < cin0; k += blockDim.x) buf[0][k] = x[(size_t)row * x_ld + k];
  __syncthreads();
  int cur = 0;
  for (int l = 0; l < ch.n_layers; ++l) {
    const pw_dense_layer& L = ch.layer[l];
    const bool last = l == ch.n_layers - 1;
    const int kper = (L.cin + CHAIN_WARPS - 1) / CHAIN_WARPS;
    const int k0 = warp * kper, k1 = min(L.cin, k0 + kper);
    for (int g0 = 0; g0 < L.cout; g0 += 32 * 8) {          // 8 groups of 32 channels at a time
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int k = k0; k < k1; ++k) {
        const float xv = buf[cur][k];
        const float* w = L.w + (size_t)k * L.w_ld + g0 + lane;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (g0 + j * 32 + lane < L.cout) acc[j] = fmaf(xv, __ldg(w + j * 32), acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (g0 + j * 32 + lane < L.cout) part[warp][g0 + j * 32 + lane] = acc[j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < L.cout; c += blockDim.x) {
      float acc = part[0][c];
#pragma unroll
      for (int w = 1; w < CHAIN_WARPS; ++w) acc += part[w][c];
      const float sc = L.scale ? __ldg(L.scale + c) : 1.f;
      const float bi = L.bias ? __ldg(L.bias + c) : 0.f;
      const float v = pw_activate(fmaf(acc, sc, bi), L.act);
      if (last) ch.out[(size_t)row * ch.out_ld + c] = v;
      else buf[cur ^ 1][c] = v;
    }
    // a layer's cin may exceed the previous cout by its padding to 4: zero those inputs
    if (!last)
      for (int c = L.cout + threadIdx.x; c < ch.layer[l + 1].cin; c += blockDim.x) buf[cur ^ 1][c] = 0.f;
    __syncthreads();
    cur ^= 1;
  }
}
}  // namespace

PW_API int pw_dense_chains(const pw_dense_chain* chains, int n_chains, const float* x, int x_ld,
                           int rows, void* stream) {
  PW_REQUIRE(chains && x && n_chains >= 1 && n_chains <= PW_MAX_CHAINS && rows >= 0);
  if (rows == 0) return 0;
  ChainPack pack;
  for (int i = 0; i < n_chains; ++i) {
    const pw_dense_chain& c = chains[i];
    PW_REQUIRE(c.n_layers >= 1 && c.n_layers <= PW_MAX_CHAIN_LAYERS && c.out);
    PW_REQUIRE(c.layer[0].cin == chains[0].layer[0].cin && c.layer[0].cin <= x_ld);
    for (int l = 0; l < c.n_layers; ++l) {
      PW_REQUIRE(c.layer[l].w && c.layer[l].cin > 0 && c.layer[l].cout > 0 &&
                 c.layer[l].cin <= CHAIN_MAX_WIDTH && c.layer[l].cout <= CHAIN_MAX_WIDTH &&
                 c.layer[l].w_ld >= c.layer[l].cout);
      PW_REQUIRE(l == 0 || (c.layer[l].cin >= c.layer[l - 1].cout &&
                            c.layer[l].cin <= c.layer[l - 1].cout + 3));
    }
    PW_REQUIRE(c.out_ld >= c.layer[c.n_layers - 1].cout);
    pack.c[i] = c;
  }
  dense_chains_kernel<<<dim3((unsigned)rows, (unsigned)n_chains), CHAIN_WARPS * 32, 0, ST>>>(pack, x, x_ld, rows);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_copy_channels(const float* x, int x_ld, float* y, int y_ld, long long pixels, int c,
                            void* stream) {
  PW_REQUIRE(x && y && (c & 3) == 0 && (x_ld & 3) == 0 && (y_ld & 3) == 0);
  copy_channels_kernel<<<grid_for(pixels * (c / 4)), TPB, 0, ST>>>(x, x_ld, y, y_ld, pixels,
                                                                    c / 4);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_argmax_zyx_to_xyz(const float* logits, int ld, int ncls, unsigned char* occ, int gx,
                                int gy, int gz, void* stream) {
  PW_REQUIRE(logits && occ && ncls > 0 && ld >= ncls);
  argmax_zyx_to_xyz_kernel<<<grid_for((long long)gx * gy * gz), TPB, 0, ST>>>(
      logits, ld, ncls, occ, nullptr, 0, 0, gx, gy, gz);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_argmax_geo_zyx_to_xyz(const float* logits, int ld, int ncls, int free_idx,
                                    int geo_value, unsigned char* occ, unsigned char* geo, int gx,
                                    int gy, int gz, void* stream) {
  PW_REQUIRE(logits && occ && geo && ncls > 0 && ld >= ncls);
  argmax_zyx_to_xyz_kernel<<<grid_for((long long)gx * gy * gz), TPB, 0, ST>>>(
      logits, ld, ncls, occ, geo, free_idx, geo_value, gx, gy, gz);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

// Strided rows -> contiguous rows on `stream` (cudaMemcpy2DAsync, direction from
// the pointers): ONE call moves a frame's images out of the loader's camera-major
// host batch.  Not counted as a kernel launch.
PW_API int pw_copy_rows(void* dst, long long dst_pitch, const void* src, long long src_pitch,
                        long long row_bytes, long long rows, void* stream) {
  PW_REQUIRE(dst && src && row_bytes > 0 && rows > 0 && row_bytes <= dst_pitch &&
             row_bytes <= src_pitch);
  cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch,
                                    (size_t)row_bytes, (size_t)rows, cudaMemcpyDefault, ST);
  return e == cudaSuccess ? 0 : (int)e;
}

PW_API int pw_density_occ_zyx_to_xyz(const float* density, int density_ld, const float* semantic,
                                     int ld, int ncls, float thr, int empty_idx,
                                     unsigned char* occ, unsigned char* geo, int gx, int gy,
                                     int gz, void* stream) {
  PW_REQUIRE(density && semantic && occ && ncls > 0 && ld >= ncls);
  density_occ_kernel<<<grid_for((long long)gx * gy * gz), TPB, 0, ST>>>(
      density, density_ld, semantic, ld, ncls, thr, empty_idx, occ, geo, gx, gy, gz);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_zyx_to_xyz(const float* x, float* y, int b, int gz, int gy, int gx, int c,
                         void* stream) {
  PW_REQUIRE(x && y && (c & 3) == 0);
  zyx_to_xyz_kernel<<<grid_for((long long)b * gz * gy * gx * (c / 4)), TPB, 0, ST>>>(
      x, y, b, gz, gy, gx, c / 4);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}
