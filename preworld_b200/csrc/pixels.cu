// Device-side pixel pipeline of PrepareImageInputs (reference
// mmdet3d/datasets/pipelines/loading.py:954-961 img_transform_core = PIL resize -> crop ->
// transpose -> rotate, :847-854 mmlabNormalize = mmcv imnormalize with to_rgb).  The decoded
// uint8 HWC image is uploaded as it is (a quarter of the fp32 CHW tensor at the network's
// resolution) and resampled here with PIL's own integer arithmetic:
//
//   Image.resize (default filter, antialiased bicubic): two separable passes over uint8,
//     out = clip8((2^21 + sum_k in[first + k] * tap[k]) >> 22), horizontal pass first, its
//     uint8 result the input of the vertical pass (host side: pixels.py:resample_tables).
//   crop / FLIP_LEFT_RIGHT / rotate (nearest, 16.16 fixed-point inverse map, zero fill):
//     pure index arithmetic -- the vertical pass is evaluated only at the pixels the final
//     view keeps, in the order the network reads them.
//   imnormalize: channel swap, fp32((double(x) - mean) * (1 / double(std))) -- what cv2 computes
//     for a float32 image and float64 scalars -- as a 3 x 256 table built on the host, CHW.
#include "common.cuh"

#include "../../include/preworld_b200.h"

namespace {

#define ST ((cudaStream_t)stream)
constexpr int PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ unsigned char clip8(int v) {
  v >>= PRECISION_BITS;
  return (unsigned char)min(max(v, 0), 255);
}

// horizontal pass: tmp[r, x, c] for source rows row0 .. row0+rows-1
__global__ void resample_rows_kernel(const unsigned char* __restrict__ src, long long pitch,
                                     int w, int row0, int rows, const int* __restrict__ first,
                                     const int* __restrict__ count, const int* __restrict__ taps,
                                     int ksize, int identity, unsigned char* __restrict__ tmp,
                                     int nw) {
  const long long total = (long long)rows * nw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % nw);
    const int r = (int)(i / nw);
    const unsigned char* row = src + (long long)(row0 + r) * pitch;
    unsigned char* o = tmp + i * 3;
    if (identity) {
      o[0] = row[x * 3]; o[1] = row[x * 3 + 1]; o[2] = row[x * 3 + 2];
      continue;
    }
    const int f = __ldg(first + x), n = __ldg(count + x);
    const int* k = taps + (long long)x * ksize;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    const unsigned char* p = row + f * 3;
    for (int j = 0; j < n; ++j) {
      const int kv = __ldg(k + j);
      s0 += p[j * 3] * kv;
      s1 += p[j * 3 + 1] * kv;
      s2 += p[j * 3 + 2] * kv;
    }
    o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
  }
}

struct ViewParams {
  const unsigned char* tmp;
  int row0, rows, nw, nh;
  const int* first;
  const int* count;
  const int* taps;
  int ksize, identity;
  int crop_x, crop_y, flip, rotated;
  int a[6];
  const float* lut;       // [3][256]: output channel c, 8-bit value v
  float* out;
  int fh, fw;
};

__global__ void resample_view_norm_kernel(const ViewParams p) {
  const int total = p.fh * p.fw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int x = i % p.fw, y = i / p.fw;
    // rotate: output pixel (x, y) <- pixel (gx, gy) of the flipped crop (nearest)
    int gx = x, gy = y;
    bool inside = true;
    if (p.rotated) {
      gx = (p.a[2] + y * p.a[1] + x * p.a[0]) >> 16;
      gy = (p.a[5] + y * p.a[4] + x * p.a[3]) >> 16;
      inside = gx >= 0 && gx < p.fw && gy >= 0 && gy < p.fh;
    }
    if (p.flip) gx = p.fw - 1 - gx;
    const int rx = gx + p.crop_x, ry = gy + p.crop_y;            // crop: zero outside
    inside = inside && rx >= 0 && rx < p.nw && ry >= 0 && ry < p.nh;
    int v[3] = {0, 0, 0};
    if (inside) {
      if (p.identity) {
        const unsigned char* q = p.tmp + ((long long)(ry - p.row0) * p.nw + rx) * 3;
        v[0] = q[0]; v[1] = q[1]; v[2] = q[2];
      } else {
        const int f = __ldg(p.first + ry), n = __ldg(p.count + ry);
        const int* k = p.taps + (long long)ry * p.ksize;
        int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
        const unsigned char* q = p.tmp + ((long long)(f - p.row0) * p.nw + rx) * 3;
        const long long step = (long long)p.nw * 3;
        for (int j = 0; j < n; ++j, q += step) {
          const int kv = __ldg(k + j);
          s0 += q[0] * kv;
          s1 += q[1] * kv;
          s2 += q[2] * kv;
        }
        v[0] = clip8(s0); v[1] = clip8(s1); v[2] = clip8(s2);
      }
    }
    // imnormalize(to_rgb=True): output channel c reads PIL channel 2 - c
#pragma unroll
    for (int c = 0; c < 3; ++c)
      p.out[(long long)c * total + i] = __ldg(p.lut + c * 256 + v[2 - c]);
  }
}

}  // namespace

PW_API int pw_resample_rows_u8(const unsigned char* src, long long src_pitch, int h, int w,
                               int row0, int rows, const int* first, const int* count,
                               const int* taps, int ksize, int identity, unsigned char* tmp,
                               int nw, void* stream) {
  PW_REQUIRE(src && tmp && h > 0 && w > 0 && rows > 0 && row0 >= 0 && row0 + rows <= h && nw > 0);
  PW_REQUIRE(identity ? nw == w : (first && count && taps && ksize > 0));
  PW_REQUIRE(src_pitch >= 3ll * w);
  const long long total = (long long)rows * nw;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  resample_rows_kernel<<<blocks, 256, 0, ST>>>(src, src_pitch, w, row0, rows, first, count, taps,
                                               ksize, identity, tmp, nw);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_resample_view_norm(const unsigned char* tmp, int row0, int rows, int nw, int nh,
                                 const int* first, const int* count, const int* taps, int ksize,
                                 int identity, int crop_x, int crop_y, int flip,
                                 const int* affine_fixed, const float* lut, float* out, int fh, int fw, void* stream) {
  PW_REQUIRE(tmp && lut && out && rows > 0 && nw > 0 && nh > 0 && fh > 0 && fw > 0);
  PW_REQUIRE(identity || (first && count && taps && ksize > 0));
  PW_REQUIRE((long long)fh * fw < (1ll << 31));
  ViewParams p;
  p.tmp = tmp; p.row0 = row0; p.rows = rows; p.nw = nw; p.nh = nh;
  p.first = first; p.count = count; p.taps = taps; p.ksize = ksize; p.identity = identity;
  p.crop_x = crop_x; p.crop_y = crop_y; p.flip = flip; p.rotated = affine_fixed != nullptr;
  for (int i = 0; i < 6; ++i) p.a[i] = affine_fixed ? affine_fixed[i] : 0;
  p.lut = lut; p.out = out; p.fh = fh; p.fw = fw;
  const int total = fh * fw;
  const int blocks = (total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16;
  resample_view_norm_kernel<<<blocks, 256, 0, ST>>>(p);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}
