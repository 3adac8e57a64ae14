// Swin image backbone (reference mmdet3d/models/backbones/swin.py:679-976, the backbone of
// the shipped config configs/preworld/nuscenes/bevstereo-occ.py:45-67): the three operators
// that are not a conv/linear.  Tokens stay in ONE layout for the whole backbone -- the
// channels-last image [B,H,W,C] (== the reference's [B, H*W, C]) -- so the reference's
// pad / roll / window_partition / window_reverse / crop copies (swin.py:371-440) never
// exist: the attention kernel resolves them as index arithmetic on its loads and stores.
//
//   pw_layernorm         nn.LayerNorm over channels (norm1/norm2/norm{i}/patch_embed.norm)
//   pw_patch_merge_ln    PatchMerging.forward up to the reduction (swin.py:185-206):
//                        2x2 gather + zero pad + LayerNorm(4C)
//   pw_window_attention  ShiftWindowMSA.forward + WindowMSA.forward without the two
//                        linears (swin.py:262-300, 364-427)
#include "common.cuh"

#include <math.h>
#include <stdlib.h>

#include "../../include/preworld_b200.h"

namespace {

#define ST ((cudaStream_t)stream)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- LayerNorm ---------------------------------------------------------------------
// One warp per row, the row in registers (NV float4 per lane), two-pass moments.
// A row is `nseg` segments of seg_c channels; MERGE: segment s = ky*2+kx of output token
// (Y,X) is input token (2Y+ky, 2X+kx) (zeros outside the map, swin.py:198-199).
struct LnParams {
  const float* x;
  float* y;
  const float* gamma;
  const float* beta;
  long long rows;
  int c, x_ld, y_ld;
  float eps;
  int h, w, oh, ow, seg_c;      // MERGE only
};

template <int NV, bool MERGE>
__global__ void __launch_bounds__(256) layernorm_kernel(const LnParams p) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int nvec = p.c >> 2;
  for (long long row = warp0; row < p.rows; row += nwarps) {
    const float* seg[4];
    if (MERGE) {
      const int X = (int)(row % p.ow);
      const long long t = row / p.ow;
      const int Y = (int)(t % p.oh);
      const long long b = t / p.oh;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int yy = 2 * Y + (s >> 1), xx = 2 * X + (s & 1);
        seg[s] = (yy < p.h && xx < p.w)
                     ? p.x + ((b * p.h + yy) * (long long)p.w + xx) * p.x_ld
                     : nullptr;
      }
    } else {
      seg[0] = p.x + row * (long long)p.x_ld;
    }
    float4 v[NV];
    float s1 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int e = i * 32 + lane;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < nvec) {
        const float* src;
        if (MERGE) {
          const int ch = e * 4, s = ch / p.seg_c;
          src = seg[s] ? seg[s] + (ch - s * p.seg_c) : nullptr;
        } else {
          src = seg[0] + e * 4;
        }
        if (src) v[i] = pw_ldg4(src);
        s1 += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    const float mean = warp_sum(s1) / (float)p.c;
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i * 32 + lane < nvec) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        s2 += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rstd = 1.f / sqrtf(warp_sum(s2) / (float)p.c + p.eps);
    float* yrow = p.y + row * (long long)p.y_ld;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int e = i * 32 + lane;
      if (e < nvec) {
        const float4 g = pw_ldg4(p.gamma + e * 4), bt = pw_ldg4(p.beta + e * 4);
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + bt.x;
        o.y = (v[i].y - mean) * rstd * g.y + bt.y;
        o.z = (v[i].z - mean) * rstd * g.z + bt.z;
        o.w = (v[i].w - mean) * rstd * g.w + bt.w;
        *reinterpret_cast<float4*>(yrow + e * 4) = o;
      }
    }
  }
}

template <bool MERGE>
int launch_layernorm(const LnParams& p, cudaStream_t s) {
  const int nvec = p.c / 4;
  const int wpb = 8;
  long long blocks = (p.rows + wpb - 1) / wpb;
  if (blocks > 148 * 8) blocks = 148 * 8;
  const dim3 g((unsigned)blocks), b(wpb * 32);
  if (nvec <= 32) layernorm_kernel<1, MERGE><<<g, b, 0, s>>>(p);
  else if (nvec <= 64) layernorm_kernel<2, MERGE><<<g, b, 0, s>>>(p);
  else if (nvec <= 128) layernorm_kernel<4, MERGE><<<g, b, 0, s>>>(p);
  else if (nvec <= 256) layernorm_kernel<8, MERGE><<<g, b, 0, s>>>(p);
  else if (nvec <= 512) layernorm_kernel<16, MERGE><<<g, b, 0, s>>>(p);
  else return PW_ERR_INVALID_ARGUMENT;
  return 0;
}

// ---- shifted-window attention --------------------------------------------------------
// CTA = one (window, head); thread t = query token t of the window (head dim 32: q and the
// output row live in registers), K and V of the window in shared memory, read as broadcast
// 128-bit loads.  One pass over the keys with a running maximum instead of a [N,N] score
// tile: 41 KB of smem per CTA at N = 144, five CTAs per SM.  Window token (iy,ix) of window (wy,wx) sits at (py,px) = (wy*ws+iy,
// wx*ws+ix) of the padded, rolled map; torch.roll(-shift) puts padded-map position
// ((py+shift)%Hp, (px+shift)%Wp) there.  Positions outside [H,W] are the zero padding the
// reference adds BEFORE the qkv linear (swin.py:371-374): their k / v rows are the qkv bias.
constexpr int HD = 32;

struct AttnParams {
  const float* qkv;        // [b,h,w,qkv_ld]: q | k | v, each c = heads*32 channels
  const float* qkv_bias;   // [3c] or null
  const float* table;      // [heads][(2ws-1)^2] relative position bias, head-major
  float* out;              // [b,h,w,out_ld]
  int b, h, w, c, heads, ws, shift, qkv_ld, out_ld;
  int hp, wp, nwx, nwin;   // padded map, windows per row / per image
  float scale;
};

// Packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2: two IEEE fp32 operations per lane and
// instruction -- the plain 3-register FFMA issues at half rate).  A 128-bit shared-memory
// load IS two packed pairs, so K / V rows need no packing moves.
typedef unsigned long long f2;
__device__ __forceinline__ f2 pack2(float x, float y) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void unpack2(f2 v, float& x, float& y) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// NQ = queries per thread (1 in production; 2 is an experiment kept for the record, see the
// launcher).
template <int MAXT, int MINB, int NQ>
__global__ void __launch_bounds__(MAXT, MINB) window_attention_kernel(const AttnParams p) {
  extern __shared__ __align__(16) float smem[];
  const int n = p.ws * p.ws;
  const int tw = 2 * p.ws - 1;
  float* ks = smem;                                   // [n][32]
  float* vs = ks + n * HD;                            // [n][32]
  float* tab = vs + n * HD;                           // [tw*tw]
  int* src = reinterpret_cast<int*>(tab + tw * tw);   // [n] pixel index or -1 (padding)
  int* koff = src + n;                                // [n] ky*tw + kx
  int* kid = koff + n;                                // [n] shift-mask region

  const int head = blockIdx.x % p.heads;
  const int win = (blockIdx.x / p.heads) % p.nwin;
  const int b = blockIdx.x / (p.heads * p.nwin);
  const int wy = win / p.nwx, wx = win % p.nwx;
  const int t = threadIdx.x;

  for (int tok = t; tok < n; tok += blockDim.x) {
    const int iy = tok / p.ws, ix = tok - iy * p.ws;
    const int py = wy * p.ws + iy, px = wx * p.ws + ix;
    int sy = py + p.shift, sx = px + p.shift;
    sy -= sy >= p.hp ? p.hp : 0;
    sx -= sx >= p.wp ? p.wp : 0;
    int id = 0;
    if (p.shift > 0) {
      // img_mask regions of swin.py:381-391 (slices 0:-ws, -ws:-shift, -shift:)
      const int hr = py < p.hp - p.ws ? 0 : (py < p.hp - p.shift ? 1 : 2);
      const int wr = px < p.wp - p.ws ? 0 : (px < p.wp - p.shift ? 1 : 2);
      id = hr * 3 + wr;
    }
    src[tok] = (sy < p.h && sx < p.w) ? (b * p.h + sy) * p.w + sx : -1;
    koff[tok] = iy * tw + ix;
    kid[tok] = id;
  }
  for (int i = t; i < tw * tw; i += blockDim.x) tab[i] = __ldg(p.table + head * tw * tw + i);
  __syncthreads();
  const int hc = head * HD;
  for (int i = t; i < n * (HD / 4); i += blockDim.x) {
    const int tok = i >> 3, j = (i & 7) * 4;
    const int s = src[tok];
    float4 kv, vv;
    if (s >= 0) {
      const float* row = p.qkv + (long long)s * p.qkv_ld + hc + j;
      kv = pw_ldg4(row + p.c);
      vv = pw_ldg4(row + 2 * p.c);
    } else if (p.qkv_bias) {
      kv = pw_ldg4(p.qkv_bias + p.c + hc + j);
      vv = pw_ldg4(p.qkv_bias + 2 * p.c + hc + j);
    } else {
      kv = vv = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    *reinterpret_cast<float4*>(ks + tok * HD + j) = kv;
    *reinterpret_cast<float4*>(vs + tok * HD + j) = vv;
  }
  __syncthreads();

  // this thread's queries: t, t + stride, ... (padding rows are cropped, swin.py:421-422)
  const int stride = (n + NQ - 1) / NQ;
  int my_src[NQ], my_id[NQ], my_base[NQ];
  bool any = false;
#pragma unroll
  for (int r = 0; r < NQ; ++r) {
    const int qi = t + r * stride;
    const bool ok = t < stride && qi < n;
    my_src[r] = ok ? src[qi] : -1;
    my_id[r] = ok ? kid[qi] : 0;
    my_base[r] = ok ? koff[qi] + (p.ws - 1) * tw + p.ws - 1 : (p.ws - 1) * tw + p.ws - 1;
    any = any || my_src[r] >= 0;
  }
  if (!any) return;

  f2 q[NQ][HD / 2];
#pragma unroll
  for (int r = 0; r < NQ; ++r) {
    const float* row = p.qkv + (long long)max(my_src[r], 0) * p.qkv_ld + hc;
#pragma unroll
    for (int j = 0; j < HD / 4; ++j) {
      const float4 v4 = pw_ldg4(row + j * 4);
      q[r][2 * j] = pack2(v4.x * p.scale, v4.y * p.scale);  // q = q * self.scale (swin.py:272)
      q[r][2 * j + 1] = pack2(v4.z * p.scale, v4.w * p.scale);
    }
  }
  // One pass over the keys with a running reference m (online softmax).  m is only moved
  // -- and the partial sums rescaled -- when a score exceeds it by more than RESCALE_SLACK:
  // the weights exp(s - m) may then reach e^8, far inside fp32 range for 144 keys, and the
  // common factor cancels in the final division (same relative rounding).  The rescale
  // branch (an expf and 16 packed multiplies, taken if ANY lane of the warp needs it) was
  // 9 % of the kernel's stall samples with a strict running maximum.
  constexpr float RESCALE_SLACK = 8.f;
  const bool masked = p.shift > 0;
  f2 o[NQ][HD / 2];
  float m[NQ], l[NQ];
#pragma unroll
  for (int r = 0; r < NQ; ++r) {
    m[r] = -INFINITY;
    l[r] = 0.f;
#pragma unroll
    for (int j = 0; j < HD / 2; ++j) o[r][j] = 0ull;
  }
#pragma unroll 2
  for (int k = 0; k < n; ++k) {
    // q . k: four independent chains per query (two packed accumulators each)
    f2 a0[NQ], a1[NQ];
#pragma unroll
    for (int r = 0; r < NQ; ++r) a0[r] = a1[r] = 0ull;
    const float* kr = ks + k * HD;
#pragma unroll
    for (int j = 0; j < HD / 4; ++j) {
      const ulonglong2 kv = *reinterpret_cast<const ulonglong2*>(kr + j * 4);
#pragma unroll
      for (int r = 0; r < NQ; ++r) {
        a0[r] = fma2(q[r][2 * j], kv.x, a0[r]);
        a1[r] = fma2(q[r][2 * j + 1], kv.y, a1[r]);
      }
    }
    const int ko = koff[k], ki = kid[k];
    f2 e2[NQ];
#pragma unroll
    for (int r = 0; r < NQ; ++r) {
      float s0, s1, s2, s3;
      unpack2(a0[r], s0, s1);
      unpack2(a1[r], s2, s3);
      float s = ((s0 + s1) + (s2 + s3)) + tab[my_base[r] - ko];
      if (masked && ki != my_id[r]) s += -100.f;
      if (s > m[r] + RESCALE_SLACK) {
        const float rs = expf(m[r] - s);          // first key: exp(-inf) = 0
        l[r] *= rs;
        const f2 r2 = pack2(rs, rs);
#pragma unroll
        for (int j = 0; j < HD / 2; ++j) o[r][j] = mul2(o[r][j], r2);
        m[r] = s;
      }
      const float e = expf(s - m[r]);
      l[r] += e;
      e2[r] = pack2(e, e);
    }
    const float* vr = vs + k * HD;
#pragma unroll
    for (int j = 0; j < HD / 4; ++j) {
      const ulonglong2 vv = *reinterpret_cast<const ulonglong2*>(vr + j * 4);
#pragma unroll
      for (int r = 0; r < NQ; ++r) {
        o[r][2 * j] = fma2(e2[r], vv.x, o[r][2 * j]);
        o[r][2 * j + 1] = fma2(e2[r], vv.y, o[r][2 * j + 1]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NQ; ++r) {
    if (my_src[r] < 0) continue;
    const float inv = 1.f / l[r];
    float* orow = p.out + (long long)my_src[r] * p.out_ld + hc;
#pragma unroll
    for (int j = 0; j < HD / 4; ++j) {
      float4 r4;
      unpack2(o[r][2 * j], r4.x, r4.y);
      unpack2(o[r][2 * j + 1], r4.z, r4.w);
      r4.x *= inv; r4.y *= inv; r4.z *= inv; r4.w *= inv;
      *reinterpret_cast<float4*>(orow + j * 4) = r4;
    }
  }
}


// ---- shifted-window attention on the (legacy) tensor-core path ---------------------------
// Same contract as window_attention_kernel.  CTA = one (window, head), warp = 16 query rows;
// S = Q K^T and O = P V are mma.sync m16n8k8 TF32 MMAs with the 3xTF32 split on both operands
// (hi = rna_tf32(x), lo = rna_tf32(x - hi); hi*hi + hi*lo + lo*hi, fp32 accumulation), K and V
// of the window pre-split once per CTA into shared memory (rows padded to 36 floats: the
// B-fragment loads of both GEMMs are bank-conflict free), the softmax on the accumulator
// fragments in registers.  Keys are walked in blocks of 72 (9 n-tiles) with a running row
// maximum (two blocks for a 12 x 12 window), which keeps the score fragment at 36 registers.
// The score fragment IS the A fragment of P V: accumulator element (row g, col 2t / 2t+1) of
// n-tile j becomes A element (row g, k-slot t / t+4) of k-step j when the 8 keys of the
// group are taken in the order 0,2,4,6,1,3,5,7 -- the V rows of the B fragment are fetched
// in the same order, no shuffles.
constexpr int MMA_KB = 9;                     // n-tiles (of 8 keys) per key block
constexpr int KV_LD = 36;                     // smem row pitch of the K / V tiles, floats

__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) window_attention_mma_kernel(const AttnParams p) {
  extern __shared__ __align__(16) float smem[];
  const int n = p.ws * p.ws;
  const int np = (n + 7) & ~7;                          // keys padded to whole n-tiles
  const int tw = 2 * p.ws - 1;
  uint32_t* khi = reinterpret_cast<uint32_t*>(smem);    // [np][36]
  uint32_t* klo = khi + np * KV_LD;
  uint32_t* vhi = klo + np * KV_LD;
  uint32_t* vlo = vhi + np * KV_LD;
  float* tab = reinterpret_cast<float*>(vlo + np * KV_LD);   // [tw*tw]
  int* src = reinterpret_cast<int*>(tab + tw * tw);     // [np] pixel index or -1 (padding)
  int* koff = src + np;                                 // [np]
  int* kid = koff + np;                                 // [np]

  const int head = blockIdx.x % p.heads;
  const int win = (blockIdx.x / p.heads) % p.nwin;
  const int b = blockIdx.x / (p.heads * p.nwin);
  const int wy = win / p.nwx, wx = win % p.nwx;
  const int tid = threadIdx.x;

  for (int tok = tid; tok < np; tok += blockDim.x) {
    int s_ = -1, ko = 0, id = -1;
    if (tok < n) {
      const int iy = tok / p.ws, ix = tok - iy * p.ws;
      const int py = wy * p.ws + iy, px = wx * p.ws + ix;
      int sy = py + p.shift, sx = px + p.shift;
      sy -= sy >= p.hp ? p.hp : 0;
      sx -= sx >= p.wp ? p.wp : 0;
      id = 0;
      if (p.shift > 0) {
        const int hr = py < p.hp - p.ws ? 0 : (py < p.hp - p.shift ? 1 : 2);
        const int wr = px < p.wp - p.ws ? 0 : (px < p.wp - p.shift ? 1 : 2);
        id = hr * 3 + wr;
      }
      s_ = (sy < p.h && sx < p.w) ? (b * p.h + sy) * p.w + sx : -1;
      ko = iy * tw + ix;
    }
    src[tok] = s_; koff[tok] = ko; kid[tok] = id;
  }
  for (int i = tid; i < tw * tw; i += blockDim.x) tab[i] = __ldg(p.table + head * tw * tw + i);
  __syncthreads();
  const int hc = head * HD;
  for (int i = tid; i < np * (HD / 4); i += blockDim.x) {
    const int tok = i >> 3, j = (i & 7) * 4;
    float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
    if (tok < n) {
      const int s_ = src[tok];
      if (s_ >= 0) {
        const float* row = p.qkv + (long long)s_ * p.qkv_ld + hc + j;
        kv = pw_ldg4(row + p.c);
        vv = pw_ldg4(row + 2 * p.c);
      } else if (p.qkv_bias) {
        kv = pw_ldg4(p.qkv_bias + p.c + hc + j);
        vv = pw_ldg4(p.qkv_bias + 2 * p.c + hc + j);
      }
    }
    const float kf[4] = {kv.x, kv.y, kv.z, kv.w}, vf[4] = {vv.x, vv.y, vv.z, vv.w};
    uint4 kh, kl, vh, vl;
    uint32_t* khp = &kh.x; uint32_t* klp = &kl.x; uint32_t* vhp = &vh.x; uint32_t* vlp = &vl.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      khp[e] = tf32_rna(kf[e]);
      klp[e] = tf32_rna(kf[e] - __uint_as_float(khp[e]));
      vhp[e] = tf32_rna(vf[e]);
      vlp[e] = tf32_rna(vf[e] - __uint_as_float(vhp[e]));
    }
    *reinterpret_cast<uint4*>(khi + tok * KV_LD + j) = kh;
    *reinterpret_cast<uint4*>(klo + tok * KV_LD + j) = kl;
    *reinterpret_cast<uint4*>(vhi + tok * KV_LD + j) = vh;
    *reinterpret_cast<uint4*>(vlo + tok * KV_LD + j) = vl;
  }
  __syncthreads();

  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = warp * 16 + g, q1 = q0 + 8;            // this thread's two query rows
  const int src0 = q0 < n ? src[q0] : -1, src1 = q1 < n ? src[q1] : -1;
  // a warp whose 16 rows are all padding (or beyond the window) has nothing to store
  if (__ballot_sync(0xffffffffu, src0 >= 0 || src1 >= 0) == 0u) return;
  const int base0 = (q0 < n ? koff[q0] : 0) + (p.ws - 1) * tw + p.ws - 1;
  const int base1 = (q1 < n ? koff[q1] : 0) + (p.ws - 1) * tw + p.ws - 1;
  const int id0 = q0 < n ? kid[q0] : -2, id1 = q1 < n ? kid[q1] : -2;
  const bool masked = p.shift > 0;

  // Q fragments (4 k-steps of 8 channels), scaled, split hi / lo
  uint32_t qh[4][4], ql[4][4];
  {
    const float* r0 = p.qkv + (long long)max(src0, 0) * p.qkv_ld + hc;
    const float* r1 = p.qkv + (long long)max(src1, 0) * p.qkv_ld + hc;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const float v[4] = {__ldg(r0 + ks * 8 + t) * p.scale, __ldg(r1 + ks * 8 + t) * p.scale,
                          __ldg(r0 + ks * 8 + t + 4) * p.scale,
                          __ldg(r1 + ks * 8 + t + 4) * p.scale};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        qh[ks][e] = tf32_rna(v[e]);
        ql[ks][e] = tf32_rna(v[e] - __uint_as_float(qh[ks][e]));
      }
    }
  }
  float o[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[nt][e] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  const int ntiles = np >> 3;
  for (int kb = 0; kb < ntiles; kb += MMA_KB) {
    const int nb = min(MMA_KB, ntiles - kb);             // n-tiles of this key block
    float sc[MMA_KB][4];
#pragma unroll
    for (int j = 0; j < MMA_KB; ++j) {
      sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
      if (j < nb) {
        const int key = (kb + j) * 8 + g;                // B fragment: n = g, k = t / t + 4
        const uint32_t* kh_ = khi + key * KV_LD + t;
        const uint32_t* kl_ = klo + key * KV_LD + t;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t bh0 = kh_[ks * 8], bh1 = kh_[ks * 8 + 4];
          const uint32_t bl0 = kl_[ks * 8], bl1 = kl_[ks * 8 + 4];
          mma_tf32(sc[j], ql[ks], bh0, bh1);
          mma_tf32(sc[j], qh[ks], bl0, bl1);
          mma_tf32(sc[j], qh[ks], bh0, bh1);
        }
      }
    }
    // relative position bias, shift mask, padded key columns; block row maxima
    float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < MMA_KB; ++j) {
      if (j < nb) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = (kb + j) * 8 + 2 * t + e;
          const int ko = koff[key], ki = kid[key];
          float s0 = sc[j][e] + tab[base0 - ko], s1 = sc[j][2 + e] + tab[base1 - ko];
          if (masked && ki != id0) s0 += -100.f;
          if (masked && ki != id1) s1 += -100.f;
          if (key >= n) { s0 = -INFINITY; s1 = -INFINITY; }
          sc[j][e] = s0; sc[j][2 + e] = s1;
          bm0 = fmaxf(bm0, s0); bm1 = fmaxf(bm1, s1);
        }
      }
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);
    const float r0 = expf(m0 - mn0), r1 = expf(m1 - mn1);       // first block: exp(-inf) = 0
    m0 = mn0; m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int j = 0; j < MMA_KB; ++j) {
      if (j < nb) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float e0 = expf(sc[j][e] - mn0), e1 = expf(sc[j][2 + e] - mn1);
          sc[j][e] = e0; sc[j][2 + e] = e1;
          rs0 += e0; rs1 += e1;
        }
      }
    }
    rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1);
    rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
    rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1);
    rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
    l0 = l0 * r0 + rs0;
    l1 = l1 * r1 + rs1;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      o[nt][0] *= r0; o[nt][1] *= r0; o[nt][2] *= r1; o[nt][3] *= r1;
    }
    // O += P V: k-step j = the 8 keys of n-tile j in the order 0,2,4,6,1,3,5,7
#pragma unroll
    for (int j = 0; j < MMA_KB; ++j) {
      if (j < nb) {
        uint32_t ph[4], pl[4];
        const float pv[4] = {sc[j][0], sc[j][2], sc[j][1], sc[j][3]};   // a0 a1 a2 a3
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          ph[e] = tf32_rna(pv[e]);
          pl[e] = tf32_rna(pv[e] - __uint_as_float(ph[e]));
        }
        const int key0 = (kb + j) * 8 + 2 * t;            // B fragment: k = t -> key 2t, t+4 -> 2t+1
        const uint32_t* vh_ = vhi + key0 * KV_LD + g;
        const uint32_t* vl_ = vlo + key0 * KV_LD + g;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const uint32_t bh0 = vh_[nt * 8], bh1 = vh_[KV_LD + nt * 8];
          const uint32_t bl0 = vl_[nt * 8], bl1 = vl_[KV_LD + nt * 8];
          mma_tf32(o[nt], pl, bh0, bh1);
          mma_tf32(o[nt], ph, bl0, bl1);
          mma_tf32(o[nt], ph, bh0, bh1);
        }
      }
    }
  }
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  if (src0 >= 0) {
    float* orow = p.out + (long long)src0 * p.out_ld + hc + 2 * t;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
      *reinterpret_cast<float2*>(orow + nt * 8) = make_float2(o[nt][0] * i0, o[nt][1] * i0);
  }
  if (src1 >= 0) {
    float* orow = p.out + (long long)src1 * p.out_ld + hc + 2 * t;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
      *reinterpret_cast<float2*>(orow + nt * 8) = make_float2(o[nt][2] * i1, o[nt][3] * i1);
  }
}

}  // namespace

PW_API int pw_layernorm(const float* x, int x_ld, const float* gamma, const float* beta,
                        float eps, float* y, int y_ld, long long rows, int c, void* stream) {
  PW_REQUIRE(x && y && gamma && beta && rows > 0 && c > 0 && c % 4 == 0 && c <= 2048);
  PW_REQUIRE(x_ld >= c && y_ld >= c && x_ld % 4 == 0 && y_ld % 4 == 0);
  LnParams p{x, y, gamma, beta, rows, c, x_ld, y_ld, eps, 0, 0, 0, 0, c};
  const int rc = launch_layernorm<false>(p, ST);
  if (rc) return rc;
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_patch_merge_ln(const float* x, int x_ld, int b, int h, int w, int c,
                             const float* gamma, const float* beta, float eps, float* y,
                             int y_ld, void* stream) {
  PW_REQUIRE(x && y && gamma && beta && b > 0 && h > 0 && w > 0 && c > 0 && c % 4 == 0);
  PW_REQUIRE(4 * c <= 2048 && x_ld >= c && x_ld % 4 == 0 && y_ld >= 4 * c && y_ld % 4 == 0);
  const int oh = (h + 1) / 2, ow = (w + 1) / 2;
  LnParams p{x, y, gamma, beta, (long long)b * oh * ow, 4 * c, x_ld, y_ld, eps, h, w, oh, ow, c};
  const int rc = launch_layernorm<true>(p, ST);
  if (rc) return rc;
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}

PW_API int pw_window_attention(const float* qkv, int qkv_ld, const float* qkv_bias,
                               const float* table, float* out, int out_ld, int b, int h, int w,
                               int c, int heads, int ws, int shift, float scale, void* stream) {
  PW_REQUIRE(qkv && table && out && b > 0 && h > 0 && w > 0 && heads > 0 && c == heads * HD);
  PW_REQUIRE(ws >= 1 && ws * ws <= 256 && shift >= 0 && shift < ws);
  PW_REQUIRE(qkv_ld >= 3 * c && qkv_ld % 4 == 0 && out_ld >= c && out_ld % 4 == 0);
  PW_REQUIRE((long long)b * h * w < (1ll << 31));
  AttnParams p;
  p.qkv = qkv; p.qkv_bias = qkv_bias; p.table = table; p.out = out;
  p.b = b; p.h = h; p.w = w; p.c = c; p.heads = heads; p.ws = ws; p.shift = shift;
  p.qkv_ld = qkv_ld; p.out_ld = out_ld;
  p.hp = (h + ws - 1) / ws * ws; p.wp = (w + ws - 1) / ws * ws;
  p.nwx = p.wp / ws; p.nwin = (p.hp / ws) * p.nwx;
  p.scale = scale;
  const int n = ws * ws, tw = 2 * ws - 1;
  const size_t smem = (size_t)(2 * n * HD + tw * tw) * 4 + (size_t)3 * n * 4;
  const long long blocks = (long long)b * p.nwin * heads;
  PW_REQUIRE(blocks < (1ll << 31));
  // Tensor-core path (mma.sync TF32, 3xTF32 split): PW_ATTN_MMA=0 selects the fp32 SIMT kernel
  // (read per call, not cached: the tests switch variants inside one process)
  const char* mma_s = getenv("PW_ATTN_MMA");
  const int mma_env = mma_s ? atoi(mma_s) : 1;
  if (mma_env != 0) {
    const int np = (n + 7) & ~7;
    const size_t smem_m = (size_t)(4 * np * KV_LD + tw * tw) * 4 + (size_t)3 * np * 4;
    const int warps = (n + 15) / 16;
    // windows up to 12 x 12 (9 warps): two CTAs per SM at 112 registers per thread
    // (PW_ATTN_MMA=2: one CTA per SM, 128 registers); larger windows: up to 16 warps
    auto km = warps * 32 <= 288 ? (mma_env == 2 ? window_attention_mma_kernel<288, 1>
                                                : window_attention_mma_kernel<288, 2>)
                                : window_attention_mma_kernel<512, 1>;
    cudaError_t em = cudaFuncSetAttribute(km, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          (int)cudaSharedmemCarveoutMaxShared);
    if (em == cudaSuccess)
      em = cudaFuncSetAttribute(km, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_m);
    if (em != cudaSuccess) return (int)em;
    km<<<(unsigned)blocks, warps * 32, smem_m, ST>>>(p);
    PW_LAUNCH_CHECK(); pw_count_launch(1);
    return 0;
  }
  // One query per thread.  PW_ATTN_NQ=2 selects two queries per thread (96 threads for the
  // 144 tokens of a 12 x 12 window, one LDS.128 per four FFMA2): bit-identical results,
  // measured 19 % SLOWER (1071 vs 897 us at stage 0 of Swin-B) -- a quarter of the lanes of
  // its third warp idle and 168 registers per thread leave 9 warps per SM.
  const char* nq_s = getenv("PW_ATTN_NQ");
  const int nq_env = nq_s ? atoi(nq_s) : 1;
  const int nq = (nq_env == 2 && n <= 192) ? 2 : 1;
  const int threads = ((n + nq - 1) / nq + 31) / 32 * 32;
  auto kern = nq == 2 ? window_attention_kernel<96, 3, 2>
                      : (threads <= 160 ? window_attention_kernel<160, 3, 1>
                                        : window_attention_kernel<256, 2, 1>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       (int)cudaSharedmemCarveoutMaxShared);
  if (e == cudaSuccess && smem > 48 * 1024)
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<(unsigned)blocks, threads, smem, ST>>>(p);
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}
