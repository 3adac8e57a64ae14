// Voxel lift (LSS view transform): frustum geometry -> voxel rank -> pooled
// volume.  HBM-bound: the only large stream is the 4*Z*Y*X*C-byte output
// volume, which this file writes exactly once (every voxel, empty ones as
// zeros) in the channels-last layout the 3-D encoder consumes.
//
// Reference path replaced (getterupper/PreWorld @ 0b0e021):
//   necks/view_transformer.py:114-153  get_lidar_coor
//   necks/view_transformer.py:203-261  voxel_pooling_prepare_v2 (long-cast,
//                                      in-range mask, argsort, intervals)
//   ops/bev_pool_v2/bev_pool.py:17-41,86-92 + src/bev_pool_cuda.cu:21-48
//   (zero-fill of `out`, the pooling kernel, the permute(0,4,1,2,3) copy)
#include "common.cuh"
#include "../../include/preworld_b200.h"

namespace {

__device__ __forceinline__ float dot3(const float* m, float x, float y, float z) {
  float acc = m[0] * x;
  acc = fmaf(m[1], y, acc);
  return fmaf(m[2], z, acc);
}

// One thread per frustum point p = (((b*N+n)*D+d)*H+h)*W+w.
// Arithmetic order is pinned to oracle/oracle_ref.c:pw_ref_lift_ranks.
__global__ void lift_rank_kernel(const float* __restrict__ cam, const float* __restrict__ bda,
                                 const float* __restrict__ xs, const float* __restrict__ ys,
                                 const float* __restrict__ ds, float lx, float ly, float lz,
                                 float ix_, float iy_, float iz_, int B, int N, int D, int H,
                                 int W, int gx, int gy, int gz, int* __restrict__ rank,
                                 int* __restrict__ count) {
  long long total = (long long)B * N * D * H * W;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total;
       p += (long long)gridDim.x * blockDim.x) {
    int w = (int)(p % W);
    long long t = p / W;
    int h = (int)(t % H); t /= H;
    int d = (int)(t % D); t /= D;
    int bn = (int)t;
    int b = bn / N;
    const float* c = cam + (long long)bn * PW_LIFT_CAM_FLOATS;
    const float* bd = bda + b * 9;
    float px = __ldg(xs + w) - c[9], py = __ldg(ys + h) - c[10], pz = __ldg(ds + d) - c[11];
    float qx = dot3(c + 0, px, py, pz), qy = dot3(c + 3, px, py, pz), qz = dot3(c + 6, px, py, pz);
    qx = qx * qz;
    qy = qy * qz;
    float ex = dot3(c + 12, qx, qy, qz) + c[21];
    float ey = dot3(c + 15, qx, qy, qz) + c[22];
    float ez = dot3(c + 18, qx, qy, qz) + c[23];
    float fx = dot3(bd + 0, ex, ey, ez), fy = dot3(bd + 3, ex, ey, ez), fz = dot3(bd + 6, ex, ey, ez);
    // (coor - lower) / interval -> .long(): truncation toward zero keeps
    // points in (-1,0) voxel units in voxel 0 (view_transformer.py:226-236)
    float vx = __fdiv_rn(fx - lx, ix_), vy = __fdiv_rn(fy - ly, iy_), vz = __fdiv_rn(fz - lz, iz_);
    long long cx = (long long)vx, cy = (long long)vy, cz = (long long)vz;
    bool ok = cx >= 0 && cx < gx && cy >= 0 && cy < gy && cz >= 0 && cz < gz;
    int r = ok ? (int)((((long long)b * gz + cz) * gy + cy) * gx + cx) : -1;
    rank[p] = r;
    if (count != nullptr && ok) atomicAdd(count + r, 1);
  }
}

// start[v] = running offset of voxel v's point list.  The list placement is
// arbitrary (one atomic per warp) -- the pooling kernel orders the points of
// a voxel itself, so the result does not depend on it.
__global__ void lift_offsets_kernel(const int* __restrict__ count, int* __restrict__ start,
                                    int* __restrict__ cursor, int* __restrict__ fill, int V) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  int lane = threadIdx.x & 31;
  int c = v < V ? count[v] : 0;
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  int total = __shfl_sync(0xffffffffu, incl, 31);
  int base = 0;
  if (lane == 31 && total > 0) base = atomicAdd(cursor, total);
  base = __shfl_sync(0xffffffffu, base, 31);
  if (v < V) {
    start[v] = base + incl - c;
    fill[v] = 0;
  }
}

__global__ void lift_fill_kernel(const int* __restrict__ rank, const int* __restrict__ start,
                                 int* __restrict__ fill, int* __restrict__ list, long long P) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < P;
       p += (long long)gridDim.x * blockDim.x) {
    int r = rank[p];
    if (r >= 0) list[start[r] + atomicAdd(fill + r, 1)] = (int)p;
  }
}

// One warp per group of 32 consecutive voxels; lane = channel while pooling.
// Points of a voxel are consumed in ascending frustum index (== the stable
// sort order of the oracle), each as psum = fmaf(feat, depth, psum) exactly
// like bev_pool_cuda.cu:38-42 compiled with -fmad=true.
template <int CPL>   // channels per lane: C <= 32*CPL
__global__ void __launch_bounds__(256)
lift_pool_kernel(const float* __restrict__ depth, const float* __restrict__ feat, int feat_ld,
                 const int* __restrict__ count, const int* __restrict__ start,
                 const int* __restrict__ list, int C, int D, int HW, long long V,
                 float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long v0 = warp * 32; v0 < V; v0 += nwarps * 32) {
    long long vmine = v0 + lane;
    int cnt = vmine < V ? __ldg(count + vmine) : 0;
    int st = vmine < V ? __ldg(start + vmine) : 0;
    int nvox = (int)min((long long)32, V - v0);
    unsigned nonempty = __ballot_sync(0xffffffffu, cnt > 0);
    // zero the empty voxels of this group first (coalesced 128-byte rows)
    for (int j = 0; j < nvox; ++j) {
      if ((nonempty >> j) & 1u) continue;
      float* o = out + (v0 + j) * C;
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        int ch = lane + 32 * q;
        if (ch < C) o[ch] = 0.f;
      }
    }
    while (nonempty) {
      int j = __ffs(nonempty) - 1;
      nonempty &= nonempty - 1;
      int n = __shfl_sync(0xffffffffu, cnt, j);
      int s = __shfl_sync(0xffffffffu, st, j);
      float acc[CPL];
#pragma unroll
      for (int q = 0; q < CPL; ++q) acc[q] = 0.f;
      int last = -1;
      for (int it = 0; it < n; ++it) {
        // next point index greater than `last`
        int best = 0x7fffffff;
        for (int k = lane; k < n; k += 32) {
          int pnt = __ldg(list + s + k);
          if (pnt > last && pnt < best) best = pnt;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        last = best;
        // p = (bn*D + d)*HW + hw  ->  feature row bn*HW + hw
        int hw = best % HW;
        int bn = best / HW / D;
        float dv = __ldg(depth + best);
        const float* f = feat + ((long long)bn * HW + hw) * feat_ld;
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
          int ch = lane + 32 * q;
          if (ch < C) acc[q] = fmaf(__ldg(f + ch), dv, acc[q]);
        }
      }
      float* o = out + (v0 + j) * C;
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        int ch = lane + 32 * q;
        if (ch < C) o[ch] = acc[q];
      }
    }
  }
}

// Drop-in for the reference kernel: one thread per (interval, channel).
__global__ void bev_pool_v2_kernel(int c, int n_intervals, const float* __restrict__ depth,
                                   const float* __restrict__ feat,
                                   const int* __restrict__ ranks_depth,
                                   const int* __restrict__ ranks_feat,
                                   const int* __restrict__ ranks_bev,
                                   const int* __restrict__ interval_starts,
                                   const int* __restrict__ interval_lengths,
                                   float* __restrict__ out) {
  long long total = (long long)n_intervals * c;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int index = (int)(idx / c);
    int cur_c = (int)(idx - (long long)index * c);
    int s = __ldg(interval_starts + index);
    int len = __ldg(interval_lengths + index);
    float psum = 0.f;
    for (int i = 0; i < len; ++i)
      psum = fmaf(__ldg(feat + (long long)__ldg(ranks_feat + s + i) * c + cur_c),
                  __ldg(depth + __ldg(ranks_depth + s + i)), psum);
    out[(long long)__ldg(ranks_bev + s) * c + cur_c] = psum;
  }
}

struct LiftWs {
  int* rank;    // [P]
  int* count;   // [V]
  int* start;   // [V]
  int* fill;    // [V]
  int* list;    // [P]
  int* cursor;  // [1] (+pad)
};

inline long long align256(long long x) { return (x + 255) / 256 * 256; }

inline LiftWs carve(void* ws, long long P, long long V) {
  char* p = (char*)ws;
  LiftWs w;
  w.rank = (int*)p; p += align256(P * 4);
  w.count = (int*)p; p += align256(V * 4);
  w.cursor = (int*)p; p += 256;
  w.start = (int*)p; p += align256(V * 4);
  w.fill = (int*)p; p += align256(V * 4);
  w.list = (int*)p;
  return w;
}

}  // namespace

PW_API long long pw_lift_workspace_bytes(int b, int n, int d, int h, int w, int gx, int gy,
                                         int gz) {
  long long P = (long long)b * n * d * h * w, V = (long long)b * gx * gy * gz;
  return 2 * align256(P * 4) + 3 * align256(V * 4) + 256;
}

PW_API int pw_lift_ranks(const float* cam, const float* bda, const float* xs, const float* ys,
                         const float* ds, const float* lower, const float* interval, int b, int n,
                         int d, int h, int w, int gx, int gy, int gz, int* rank, void* stream) {
  PW_REQUIRE(cam && bda && xs && ys && ds && lower && interval && rank);
  long long P = (long long)b * n * d * h * w;
  PW_REQUIRE(P > 0 && P < (1ll << 31) && (long long)b * gx * gy * gz < (1ll << 31));
  int blocks = (int)min((long long)148 * 8, (P + 255) / 256);
  lift_rank_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      cam, bda, xs, ys, ds, lower[0], lower[1], lower[2], interval[0], interval[1], interval[2], b,
      n, d, h, w, gx, gy, gz, rank, nullptr);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_lift_fused(const float* depth, const float* feat, int feat_ld, const float* cam,
                         const float* bda, const float* xs, const float* ys, const float* ds,
                         const float* lower, const float* interval, int b, int n, int d, int h,
                         int w, int c, int gx, int gy, int gz, float* out, void* workspace,
                         void* stream) {
  PW_REQUIRE(depth && feat && cam && bda && xs && ys && ds && lower && interval && out && workspace);
  PW_REQUIRE(c > 0 && c <= 128 && feat_ld >= c);
  long long P = (long long)b * n * d * h * w, V = (long long)b * gx * gy * gz;
  PW_REQUIRE(P > 0 && P < (1ll << 31) && V > 0 && V < (1ll << 31));
  cudaStream_t st = (cudaStream_t)stream;
  LiftWs ws = carve(workspace, P, V);
  // count[V] and cursor are contiguous: one memset
  cudaError_t e = cudaMemsetAsync(ws.count, 0, align256(V * 4) + 256, st);
  if (e != cudaSuccess) return (int)e;
  int blocks = (int)min((long long)148 * 8, (P + 255) / 256);
  lift_rank_kernel<<<blocks, 256, 0, st>>>(cam, bda, xs, ys, ds, lower[0], lower[1], lower[2],
                                           interval[0], interval[1], interval[2], b, n, d, h, w,
                                           gx, gy, gz, ws.rank, ws.count);
  PW_LAUNCH_CHECK();
  lift_offsets_kernel<<<pw_ceil_div(V, 256), 256, 0, st>>>(ws.count, ws.start, ws.cursor, ws.fill,
                                                           (int)V);
  PW_LAUNCH_CHECK();
  lift_fill_kernel<<<blocks, 256, 0, st>>>(ws.rank, ws.start, ws.fill, ws.list, P);
  PW_LAUNCH_CHECK();
  // 8 warps/block, 32 voxels per warp-iteration; grid sized to the SM count
  int pblocks = (int)min((long long)148 * 8, (V + 255) / 256);
  int HW = h * w;
  if (c <= 32)
    lift_pool_kernel<1><<<pblocks, 256, 0, st>>>(depth, feat, feat_ld, ws.count, ws.start, ws.list,
                                                 c, d, HW, V, out);
  else if (c <= 64)
    lift_pool_kernel<2><<<pblocks, 256, 0, st>>>(depth, feat, feat_ld, ws.count, ws.start, ws.list,
                                                 c, d, HW, V, out);
  else
    lift_pool_kernel<4><<<pblocks, 256, 0, st>>>(depth, feat, feat_ld, ws.count, ws.start, ws.list,
                                                 c, d, HW, V, out);
  PW_LAUNCH_CHECK(); pw_count_launch(4);
  return 0;
}

PW_API int pw_bev_pool_v2(int c, int n_intervals, const float* depth, const float* feat,
                          const int* ranks_depth, const int* ranks_feat, const int* ranks_bev,
                          const int* interval_starts, const int* interval_lengths, float* out,
                          void* stream) {
  PW_REQUIRE(c > 0 && n_intervals >= 0);
  if (n_intervals == 0) return 0;
  PW_REQUIRE(depth && feat && ranks_depth && ranks_feat && ranks_bev && interval_starts &&
             interval_lengths && out);
  long long total = (long long)n_intervals * c;
  int blocks = (int)min((long long)148 * 16, (total + 255) / 256);
  bev_pool_v2_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      c, n_intervals, depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_starts,
      interval_lengths, out);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}
