/*
 * preworld_b200 -- C ABI of the B200 (sm_100a) camera->voxel occupancy path.
 *
 * Drop-in boundary for getterupper/PreWorld @ 0b0e021.  Every entry point
 *   - takes raw DEVICE pointers + sizes, never allocates, never synchronises,
 *   - launches on the caller's stream (`stream` is a cudaStream_t; pass
 *     torch.cuda.current_stream().cuda_stream),
 *   - returns 0 on success, a cudaError_t value (>0) on a CUDA failure or
 *     PW_ERR_INVALID_ARGUMENT (-1) on a contract violation.
 * The reference's own FFI has no error checking and uses the legacy default
 * stream (mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:125-140).
 *
 * Layout convention: all feature maps are CHANNELS-LAST fp32
 * (images [N,H,W,C], volumes [B,Z,Y,X,C]) -- i.e. torch tensors of the
 * reference's logical shapes [N,C,H,W] / [B,C,Z,Y,X] carrying
 * torch.channels_last / channels_last_3d strides.  `*_ld` arguments are the
 * element stride between consecutive pixels/voxels (>= channel count), which
 * lets a kernel read or write a channel slice of a wider (concatenated)
 * tensor in place.
 *
 * Each declaration cites the reference interface it replaces.
 */
#ifndef PREWORLD_B200_H_
#define PREWORLD_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define PW_ABI_VERSION 1
int pw_abi_version(void);
/* Number of kernel launches issued through this library since load (the
 * `gpu_launches` figure of bench.py). */
long long pw_launch_count(void);

/* ------------------------------------------------------------------------
 * Convolution / linear (fp32, implicit GEMM, fused affine+residual+act).
 * Replaces torch.nn.Conv2d/Conv3d/Linear (+BatchNorm, +ReLU/Softplus/Sigmoid)
 * groups: mmdet ResNet; necks/fpn.py:154-203; necks/view_transformer.py:
 * 473-638; backbones/resnet.py:88-184; necks/lss_fpn.py:120-148;
 * detectors/preworld.py:72-105; heads/occupancy_head.py:81-177;
 * detectors/preworld_temporal_traj.py:119-150.
 *   y[m, co] = act( (sum_k x[..] * w[k, co]) * scale[co] + bias[co]
 *                   + residual[m, co] )
 * x: [n,d,h,w,in_ld] (first `cin` channels used); w: [kd*kh*kw*cin, w_ld]
 * (k ordered tap-major then cin; w_ld >= cout padded with zeros, w_ld%4==0);
 * scale/bias/residual may be NULL.  act: 0 none, 1 relu, 2 softplus
 * (threshold 20), 3 sigmoid, 4 gelu (erf form).  cin%4 == 0, in_ld%4 == 0, x and w 16-byte
 * aligned.
 * ---------------------------------------------------------------------- */
typedef struct pw_conv_desc {
  int n, d, h, w, cin, in_ld;
  int od, oh, ow, cout, out_ld, res_ld, w_ld;
  int kd, kh, kw;
  int sd, sh, sw;
  int pd, ph, pw;
  int dd, dh, dw;
  int act;
  int act_channels; /* >0: only output channels [0, act_channels) get `act`
                       (fuses a ReLU branch and a linear shortcut branch that
                       read the same input into one launch); 0: all */
} pw_conv_desc;

int pw_conv_fwd(const pw_conv_desc* desc, const float* x, const float* w,
                const float* scale, const float* bias, const float* residual,
                float* y, void* stream);

/* Same contract on the tcgen05 tensor cores (conv_umma.cu): TMA-staged
 * 128-pixel x 32-channel tiles, tcgen05.mma kind::tf32 with a 3-term hi/lo
 * split (fp32-level accuracy), accumulator in TMEM.  Requires cin % 32 == 0;
 * weights are passed TRANSPOSED and pre-split: wt_hi / wt_lo [cout, K]
 * (K-major, K = taps*cin tap-major), wt_hi = w with the low 13 mantissa bits
 * cleared, wt_lo = w - wt_hi.  pw_conv_umma_supported() returns 1 when this
 * descriptor can run on that path. */
int pw_conv_umma_supported(const pw_conv_desc* desc);
int pw_conv_umma_fwd(const pw_conv_desc* desc, const float* x,
                     const float* wt_hi, const float* wt_lo, const float* scale,
                     const float* bias, const float* residual, float* y,
                     void* stream);

/* Second-generation tensor-core conv (conv_halo.cu), same contract and weight
 * format as pw_conv_umma_fwd: the CTA's input box incl. halo is TMA-loaded once
 * per 32-channel chunk, the A operand is split to tf32 hi/lo in registers and
 * fed to tcgen05.mma from TENSOR MEMORY, hi|lo weights form one 2N-wide MMA.
 * pw_conv_halo_supported() returns 1 when the halo box fits shared memory
 * (all stride-1 convs of the path; 1x1 convs are flattened to a GEMM). */
int pw_conv_halo_supported(const pw_conv_desc* desc);
int pw_conv_halo_fwd(const pw_conv_desc* desc, const float* x,
                     const float* wt_hi, const float* wt_lo, const float* scale,
                     const float* bias, const float* residual, float* y,
                     void* stream);

/* The same kernel with the kw taps along x folded into the MMA's N dimension
 * (stride 1 along x, kw in 2..4): the main loop runs kd*kh taps over N =
 * kw*fold_n columns, P[row, kx, n] over INPUT columns, and the epilogue adds
 * the shifted partial rows, out[x] = sum_kx P[x + kx*dw, kx, :] -- every input
 * row is split and read from tensor memory once per (kz,ky) instead of once per
 * tap.  Weights: wf_hi/wf_lo [slabs*kw*fold_n, kd*kh*cin] (K-major), row
 * (slab*kw + kx)*fold_n + n = output channel slab*fold_n + n at x tap kx (zero
 * rows beyond cout), column (kz*kh + ky)*cin + ci; fold_n = pw_conv_fold_n(cout),
 * slabs = ceil(cout / fold_n).  Same results as pw_conv_halo_fwd up to the
 * summation order of the taps. */
int pw_conv_fold_n(int cout);
int pw_conv_fold_supported(const pw_conv_desc* desc);
int pw_conv_fold_fwd(const pw_conv_desc* desc, const float* x,
                     const float* wf_hi, const float* wf_lo, const float* scale,
                     const float* bias, const float* residual, float* y,
                     void* stream);

/* ------------------------------------------------------------------------
 * Fused per-voxel heads (pw_attr_mlp, pw_fusion_step,
 * pw_occhead_argmax).
 * ---------------------------------------------------------------------- */
/* Fused two-layer per-row MLP on the tensor cores (tcgen05 kind::tf32, 3xTF32
 * split, fp32-level accuracy):
 *     y[r, 0:n2] = act2(W2 . act1(W1 . x[r, 0:c1] + b1) + b2) [+ residual[r, 0:n2]]
 * The hidden row (`hidden` wide) never leaves the SM.  One call replaces
 *  - the forecasting step of detectors/preworld_temporal_traj.py:329-341,368
 *    (fusion_head = Linear(64,128)-Softplus-Linear(128,32) on cat([voxel, ego]) +
 *    residual): W1 = fusion_head[0].weight[:, :32], b1 = the per-sample bias
 *    fusion_head[0].weight[:, 32:] . plan_head(ego) + fusion_head[0].bias,
 *    residual = x;
 *  - the attribute projection of detectors/preworld.py:81-105,251-254
 *    (density / semantic / color MLPs, each Linear(32,64)-Softplus-Linear(64,k)):
 *    W1 = the three first layers stacked [192, 32], W2 = block diagonal [24, 192],
 *    act2 = Softplus on the first act2_channels = 2 (density) channels.
 * c1 == 32; hidden % 32 == 0, <= 256; n2 % 4 == 0, <= 32.  w1_hi / w1_lo
 * [hidden, c1] and w2_hi / w2_lo [n2p, hidden] (n2p = 16 if n2 <= 16 else 32, rows
 * >= n2 zero) are the K-major weights pre-split into rounded tf32 hi / lo parts;
 * b1 [hidden], b2 [n2] may be NULL.  act2 applies to channels [0, act2_channels).
 * x / residual / y are row arrays with pitches x_ld / res_ld / y_ld (multiples of
 * 4, 16-byte aligned); y may alias residual but not x. */
int pw_mlp2_supported(int c1, int hidden, int n2);
int pw_mlp2(const float* x, int x_ld, long long m, int c1, const float* w1_hi,
            const float* w1_lo, const float* b1, int hidden, int act1,
            const float* w2_hi, const float* w2_lo, const float* b2, int n2, int act2,
            int act2_channels, const float* residual, int res_ld, float* y, int y_ld,
            void* stream);
/* OccHead tail (heads/occupancy_head.py:95-105,147-162 + detectors/preworld.py:
 * 196-221), fused: feat [zyx voxels, feat_ld] = the 16-channel output of
 * occ_convs[0] (conv3^3 + BN + ReLU) in the library's [Z,Y,X] voxel order ->
 * relu(scale0 * (w0 . feat) + bias0) [mid = 8] -> w1 . h + bias1 [ncls] -> argmax
 * (first maximum wins) -> occ uint8 [X,Y,Z]; geo (may be NULL) = (class !=
 * free_idx) ? 0 : geo_value; logits (may be NULL) [zyx voxels, logits_ld] receives
 * the class logits.  w0 [mid, cin], w1 [ncls, mid] row-major. */
int pw_occhead_tail(const float* feat, int feat_ld, int cin, const float* w0,
                    const float* scale0, const float* bias0, int mid, const float* w1,
                    const float* bias1, int ncls, float* logits, int logits_ld,
                    unsigned char* occ, unsigned char* geo, int free_idx, int geo_value,
                    int gx, int gy, int gz, void* stream);

/* ------------------------------------------------------------------------
 * Image-side element-wise helpers (all channels-last).
 * ---------------------------------------------------------------------- */
/* imgs [n,c,h,w] (NCHW as delivered by the loader, loading.py:1124-1134;
 * consecutive images `img_stride` elements apart, so one frame of the
 * camera-major [B, N*T, 3, H, W] batch is read in place)
 * -> [n,h,w,c_pad] with zero padding channels. */
int pw_nchw_to_nhwc_pad(const float* x, long long img_stride, float* y, int n,
                        int c, int h, int w, int c_pad, void* stream);
/* Same input, space-to-depth by 2: y[n, Y, X, (dy*2+dx)*4 + c] =
 * x[n, c, 2Y+dy, 2X+dx] (c <= 4; h, w even; channels 16..c_pad-1 zero).  The
 * mmdet ResNet stem (7x7, stride 2, pad 3) becomes a stride-1 4x4 conv over
 * this tensor (taps a,b in -2..1; weight[o,(dy,dx,c),a,b] = W[o,c,2a+dy+3,
 * 2b+dx+3]), which has Cin % 32 == 0 and runs on the tensor-core kernel. */
int pw_nchw_to_s2d_nhwc(const float* x, long long img_stride, float* y, int n,
                        int c, int h, int w, int c_pad, void* stream);
/* channels-last -> NCHW copy: y[n,c,p] = x[n,p,c0+c]  (API edges only). */
int pw_nhwc_to_nchw(const float* x, int x_ld, float* y, int n, int c,
                    long long pixels, void* stream);
/* mmdet ResNet stem nn.MaxPool2d(kernel_size=3, stride=2, padding=1). */
int pw_maxpool3x3s2(const float* x, float* y, int n, int h, int w, int c,
                    int oh, int ow, void* stream);
/* necks/fpn.py:164-172: y[n,oh,ow,:] += x[n, floor(oh*h/OH), floor(ow*w/OW),:]
 * (F.interpolate mode='nearest' to the finer map's size, added in place). */
int pw_upsample_nearest_add(const float* x, float* y, int n, int h, int w,
                            int oh, int ow, int c, void* stream);
/* SELayer gate, view_transformer.py:440-470: y[n,p,c] = x[n,p,c]*gate[n,c]
 * (gate already passed through the sigmoid). */
int pw_scale_channels(const float* x, int x_ld, const float* gate, float* y,
                      int y_ld, int n, long long pixels, int c, void* stream);
/* nn.AdaptiveAvgPool2d((1,1)): y[n,c] = mean_p x[n,p,c]. */
int pw_global_avgpool(const float* x, int x_ld, float* y, int n,
                      long long pixels, int c, void* stream);
/* ASPP image-pool branch, view_transformer.py:407-410: bilinear upsampling of
 * a 1x1 map == broadcast.  y[n,p,c0:c0+c] = v[n,c]. */
int pw_broadcast_channels(const float* v, float* y, int y_ld, int n,
                          long long pixels, int c, void* stream);
/* Chains of small dense layers on a few row vectors, ONE launch: DepthNet's
 * camera-parameter branch (necks/view_transformer.py:421-470,606-617: Mlp(27->mid->mid)
 * + SELayer reduce / expand -> sigmoid gate, for the context and the depth branch).
 * Every chain reads the same input rows x [rows, x_ld]; layer l computes
 * act(scale * (in . w[:, c]) + bias) with w [cin, w_ld] (PackedConv's SIMT layout; a
 * layer's cin may exceed the previous cout by its padding to 4 -- those inputs are 0).
 * Widths <= 1024. */
#define PW_MAX_CHAINS 4
#define PW_MAX_CHAIN_LAYERS 4
typedef struct {
  const float* w;
  const float* scale;          /* may be NULL (1) */
  const float* bias;           /* may be NULL (0) */
  int cin, cout, w_ld, act;    /* act: PW_ACT_* */
} pw_dense_layer;
typedef struct {
  pw_dense_layer layer[PW_MAX_CHAIN_LAYERS];
  int n_layers;
  float* out;                  /* [rows, out_ld] */
  int out_ld;
} pw_dense_chain;
int pw_dense_chains(const pw_dense_chain* chains, int n_chains, const float* x,
                    int x_ld, int rows, void* stream);

/* view_transformer.py:801: depth = softmax over the D logits of every pixel.
 * logits [rows, in_ld] (first d channels) -> prob_cl [rows, d] (channels-last,
 * may be NULL) and prob_planar [n, d, pixels] (the reference's [B*N,D,H,W]
 * layout, rows = n*pixels; may be NULL). */
int pw_softmax_depth(const float* logits, int in_ld, float* prob_cl,
                     float* prob_planar, int n, long long pixels, int d,
                     void* stream);

/* ------------------------------------------------------------------------
 * Stereo cost volume, view_transformer.py:546-604 (gen_grid +
 * calculate_cost_volumn): for each (cam, depth bin, h4, w4) warp the previous
 * frame's stereo feature with the plane-sweep homography, bilinear sample
 * (zeros padding, align_corners=True), cost = sum_c |curr - warp|, + bias
 * where the warped last-group channel 0 is exactly 0, negate, softmax over D.
 *   curr/prev: [n,h,w,c] channels-last stereo features (c % 4 == 0);
 *   cam: [n, PW_CV_CAM_FLOATS] per-camera constants built by the host
 *   (inv(post_rot), post_tran, k2s_rot @ inv(K), k2s_tran, K, post_rot[:2,:2],
 *   post_tran[:2]);  frustum x/y/depth values; out: [n,h,w,out_ld] channels-last,
 *   channels [d, out_ld) written as zeros (padding for a 32-multiple Cin).
 * ---------------------------------------------------------------------- */
#define PW_CV_CAM_FLOATS 48
int pw_cost_volume(const float* curr, const float* prev, const float* cam,
                   const float* xs, const float* ys, const float* ds, float* out,
                   int out_ld, int n, int h, int w, int c, int d, float bias,
                   int img_h, int img_w, void* stream);

/* ------------------------------------------------------------------------
 * Voxel lift.
 * ---------------------------------------------------------------------- */
/* Drop-in for the reference FFI  bev_pool_v2(c, n_intervals, depth, feat,
 * ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths, out)
 * (mmdet3d/ops/bev_pool_v2/src/bev_pool.cpp:7-14,30-57; kernel
 * src/bev_pool_cuda.cu:21-48).  Same argument meaning; `out` must be
 * zero-initialised by the caller exactly as bev_pool.py:27 does. */
int pw_bev_pool_v2(int c, int n_intervals, const float* depth, const float* feat,
                   const int* ranks_depth, const int* ranks_feat,
                   const int* ranks_bev, const int* interval_starts,
                   const int* interval_lengths, float* out, void* stream);

/* Backward of bev_pool_v2: drop-in for the reference FFI  bev_pool_v2_grad(c,
 * n_intervals, out_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev,
 * interval_starts, interval_lengths, depth_grad, feat_grad)
 * (src/bev_pool.cpp:16-28,74-111; kernel src/bev_pool_cuda.cu:67-121).  As in
 * the reference the point lists are sorted by ranks_FEAT and the intervals
 * are runs of equal ranks_feat (bev_pool.py:47-57); depth_grad / feat_grad are
 * zero-initialised by the caller (bev_pool.py:67-68). */
int pw_bev_pool_v2_grad(int c, int n_intervals, const float* out_grad,
                        const float* depth, const float* feat,
                        const int* ranks_depth, const int* ranks_feat,
                        const int* ranks_bev, const int* interval_starts,
                        const int* interval_lengths, float* depth_grad,
                        float* feat_grad, void* stream);

/* Fused B200-native lift (ONE persistent kernel with grid barriers): geometry
 * -> voxel rank -> per-voxel point lists ->
 * dense pooled volume, replacing get_lidar_coor + voxel_pooling_prepare_v2 +
 * bev_pool_v2 + the zero-fill and the permute copy (view_transformer.py:
 * 114-153,176-261; bev_pool.py:27,91).
 *   cam: [b*n, PW_LIFT_CAM_FLOATS] = inv(post_rot)[9], post_tran[3],
 *        sensor2ego[:3,:3] @ inv(K) [9], sensor2ego[:3,3] [3]
 *   bda: [b,9]; xs[w], ys[h], ds[d]: frustum coordinates (device)
 *   lower[3], interval[3]: grid lower bound / voxel size -- HOST pointers
 *   depth: [b*n, d, h, w] probabilities (planar, as the reference)
 *   feat:  [b*n, h, w, feat_ld] channels-last context features (c used)
 *   out:   [b, gz, gy, gx, c] channels-last, EVERY voxel written (no memset)
 *   workspace: pw_lift_workspace_bytes(...) bytes of device scratch whose first
 *        256 bytes (grid-barrier / cursor words) must be ZERO before the first
 *        call; the kernel leaves them zero, so one workspace serves any number
 *        of calls on one stream.
 * Summation order inside a voxel is ascending frustum-point index -- the
 * order a stable sort gives the reference (its argsort is unstable,
 * view_transformer.py:246, so the reference's own order is unspecified). */
#define PW_LIFT_CAM_FLOATS 24
long long pw_lift_workspace_bytes(int b, int n, int d, int h, int w, int gx,
                                  int gy, int gz);
int pw_lift_fused(const float* depth, const float* feat, int feat_ld,
                  const float* cam, const float* bda, const float* xs,
                  const float* ys, const float* ds, const float* lower,
                  const float* interval, int b, int n, int d, int h, int w,
                  int c, int gx, int gy, int gz, float* out, void* workspace,
                  void* stream);
/* The same lift split in two for CONSTANT cameras -- the reference's
 * LSSViewTransformer(accelerate=True) (view_transformer.py:31-33,155-174,
 * 263-295: init_acceleration_v2 caches ranks/intervals once, every later
 * forward only calls bev_pool_v2):
 *   pw_lift_prepare builds the per-voxel point lists in `workspace`
 *   (same size / zeroed-control-words contract as pw_lift_fused);
 *   pw_lift_pool pools depth x feat with those lists: ONE pass, every output
 *   row written exactly once.  Results are identical to pw_lift_fused. */
int pw_lift_prepare(const float* cam, const float* bda, const float* xs,
                    const float* ys, const float* ds, const float* lower,
                    const float* interval, int b, int n, int d, int h, int w,
                    int gx, int gy, int gz, void* workspace, void* stream);
int pw_lift_pool(const float* depth, const float* feat, int feat_ld, int b,
                 int n, int d, int h, int w, int c, int gx, int gy, int gz,
                 float* out, void* workspace, void* stream);
/* Per-camera constant tables (device in, device out, no host round trip):
 * the torch.inverse/matmul prologues of get_lidar_coor
 * (view_transformer.py:141-150) and DepthNet.gen_grid (:552-566).
 * sensor2ego / k2s_sensor [n,4,4], intrin/post_rot [n,3,3], post_tran [n,3]
 * -> cam [n, PW_LIFT_CAM_FLOATS] resp. [n, PW_CV_CAM_FLOATS]. */
int pw_lift_camera_params(int n, const float* sensor2ego, const float* intrin,
                          const float* post_rot, const float* post_tran,
                          float* cam, void* stream);
int pw_cv_camera_params(int n, const float* k2s_sensor, const float* intrin,
                        const float* post_rot, const float* post_tran,
                        float* cam, void* stream);
/* Only the rank computation (voxel id or -1 per frustum point); exposed for
 * parity tests against voxel_pooling_prepare_v2. */
int pw_lift_ranks(const float* cam, const float* bda, const float* xs,
                  const float* ys, const float* ds, const float* lower,
                  const float* interval, int b, int n, int d, int h, int w,
                  int gx, int gy, int gz, int* rank, void* stream);

/* ------------------------------------------------------------------------
 * 3-D encoder helpers.
 * ---------------------------------------------------------------------- */
/* necks/lss_fpn.py:139-146: F.interpolate(scale_factor=s, mode='trilinear',
 * align_corners=True) of x [b,z,y,x,c] written into channel slice of
 * y [b,oz,oy,ox,y_ld]. */
int pw_upsample_trilinear(const float* x, int x_ld, float* y, int y_ld, int b,
                          int z, int yy, int xx, int c, int oz, int oy, int ox,
                          void* stream);
/* y = up(x1) + up(x2), both trilinear (align_corners=True) to [b,oz,oy,ox,c]:
 * LSSFPN3D with its 1x1x1 conv commuted in front of the interpolation
 * (lss_fpn.py:139-148), so the 224-channel concatenation never exists. */
int pw_upsample_trilinear2(const float* x1, int x1_ld, int z1, int y1, int w1,
                           const float* x2, int x2_ld, int z2, int y2, int w2,
                           float* y, int y_ld, int b, int c, int oz, int oy,
                           int ox, void* stream);
/* Strided channel-slice copy y[p, 0:c] = x[p, 0:c] (concatenation). */
int pw_copy_channels(const float* x, int x_ld, float* y, int y_ld,
                     long long pixels, int c, void* stream);
/* preworld.py:202-203 / occupancy argmax: logits [zyx voxels, ld] in the
 * library's [Z,Y,X] voxel order -> uint8 grid in the reference's [X,Y,Z] order
 * (first maximum wins, as torch.argmax). */
int pw_argmax_zyx_to_xyz(const float* logits, int ld, int ncls, unsigned char* occ,
                         int gx, int gy, int gz, void* stream);
/* preworld.py:201-221 (nuScenes branch): the same argmax plus the geometry grid
 * geo = (class != free_idx) ? 0 : geo_value, both in [X,Y,Z] order. */
int pw_argmax_geo_zyx_to_xyz(const float* logits, int ld, int ncls, int free_idx,
                             int geo_value, unsigned char* occ, unsigned char* geo,
                             int gx, int gy, int gz, void* stream);
/* Host plumbing of the public call (bevdet_occ.py:88-97 splits the loader's
 * camera-major image batch into frames): `rows` rows of `row_bytes` from a
 * pitched source into a pitched destination, asynchronously on `stream`
 * (host->device when src is pinned host memory, device->device otherwise). */
int pw_copy_rows(void* dst, long long dst_pitch, const void* src, long long src_pitch,
                 long long row_bytes, long long rows, void* stream);
/* preworld.py:173-194 density path: occ = density>thr ? argmax(semantic) : 17;
 * density [voxels] , semantic [voxels, ld]. */
int pw_density_occ_zyx_to_xyz(const float* density, int density_ld,
                              const float* semantic, int ld, int ncls, float thr,
                              int empty_idx, unsigned char* occ, unsigned char* geo,
                              int gx, int gy, int gz, void* stream);
/* [b,z,y,x,c] (library order) -> [b,x,y,z,c] (the reference's voxel_feats
 * order, preworld.py:169) for API edges. */
int pw_zyx_to_xyz(const float* x, float* y, int b, int gz, int gy, int gx, int c,
                  void* stream);

/* ------------------------------------------------------------------------
 * Volume rendering (nerf/nerf_head.py:32-55,165-269,332-353 + the CUDA ops
 * nerf/cuda/render_utils_kernel.cu:431-443,577-651 and
 * ub360_utils_kernel.cu:12-32), one warp per ray.
 * The three drop-ins keep the reference extension signatures
 * (render_utils.cpp:170-184, ub360_utils.cpp:20-22) on raw pointers.
 * ---------------------------------------------------------------------- */
int pw_raw2alpha(const float* density, float shift, float interval, long long n,
                 float* exp_d, float* alpha, void* stream);
int pw_alpha2weight(const float* alpha, const long long* ray_id, long long n_pts,
                    int n_rays, float* weight, float* T, float* alphainv_last,
                    long long* i_start, long long* i_end, void* stream);
int pw_cumdist_thres(const float* dist, float thres, int n_rays, int n_pts,
                     unsigned char* mask, void* stream);
/* Fused ray march: rays [r,16] (datasets/ray.py:49-56), volumes density
 * [voxels], semantic [voxels,17], color [voxels,3] with voxel strides given in
 * the descriptor (either voxel order works without a copy) -> per ray
 * render_depth, render_semantic[17], render_color[3], alphainv_last, and a
 * valid flag (0 < gt depth <= 52).  Invalid rays get zeros.  t_vals[n_steps]
 * are the ray parameters of sample_ray (nerf_head.py:36-44; 391 inner + 26
 * outer mid-points), computed once by the host with the reference's own
 * torch.linspace expressions. */
typedef struct pw_render_desc {
  float scene_center[3];
  float scene_radius[3];
  float xyz_min[3];
  float xyz_max[3];
  float bg_len;
  float act_shift;
  float interval;
  float step_size;
  float fast_color_thres;
  float radius;
  float max_depth;
  int world_len;
  int gx, gy, gz;
  int n_sem;
  long long vs_x, vs_y, vs_z; /* render_utils_cuda.raw2alpha_backward(exp_d, grad_back, interval) -> grad
 * (nerf/cuda/render_utils_kernel.cu:507-537, render_utils.cpp:54-60). */
int pw_raw2alpha_backward(const float* exp_d, const float* grad_back,
                          float interval, long long n, float* grad,
                          void* stream);
/* render_utils_cuda.alpha2weight_backward(alpha, weight, T, alphainv_last,
 * i_start, i_end, n_rays, grad_weights, grad_last) -> grad
 * (render_utils_kernel.cu:654-707); `grad` [n_pts] zero-initialised by the
 * caller (torch::zeros_like, :682). */
int pw_alpha2weight_backward(const float* alpha, const float* weight,
                             const float* T, const float* alphainv_last,
                             const long long* i_start, const long long* i_end,
                             int n_rays, const float* grad_weights,
                             const float* grad_last, float* grad, void* stream);

/* voxel strides (in voxels) of the volumes:
                                 library order [Z,Y,X] -> (1, gx, gx*gy);
                                 reference order [X,Y,Z] -> (gy*gz, gz, 1) */
} pw_render_desc;
int pw_render_rays(const pw_render_desc* desc, const float* rays, int n_rays,
                   const float* t_vals, int n_steps, const float* bda,
                   const float* density, int density_ld,
                   const float* semantic, int sem_ld, const float* color,
                   int col_ld, float* out_depth, float* out_sem, float* out_col,
                   float* out_last, unsigned char* out_valid, void* stream);

/* NerfHead.compute_loss's reductions over the renderings of pw_render_rays
 * (mmdet3d/models/nerf/nerf_head.py:271-291; silog_loss / l1_loss,
 * nerf/utils.py:71-87; nn.CrossEntropyLoss(weight, 'mean')), masked rays only.
 * sums [9] fp64 (zeroed inside): n, sum d, sum d^2 (d = log(depth + 1e-7) -
 * log(rays[:,2])), sum w[t] nll, sum w[t] (t = rays[:,3]), sum |color - rays[:,13:16]|
 * per channel, sum p log p + (1-p) log(1-p) of clamp(alphainv_last, 1e-6, 1-1e-6).
 * The five loss values follow on the host from these nine numbers. */
int pw_render_loss_sums(const float* rays, int n_rays, int n_sem, const float* depth,
                        const float* sem, const float* col, const float* last,
                        const unsigned char* valid, const float* class_weights,
                        double* sums, void* stream);

/* ------------------------------------------------------------------------
 * Evaluation: confusion matrices of Metric_mIoU.add_batch,
 * mmdet3d/datasets/occ_metrics.py:93-157, accumulated on the device.
 *   pred, gt: uint8 class grids (any order, n voxels); mask: uint8/bool or NULL
 *   (mask_camera / mask_lidar); hist [n_cl*n_cl] int64: hist[gt*n_cl+pred] += 1
 *   for masked voxels with gt < n_cl; occ_hist [4] int64: (gt != free)*2 +
 *   (pred != free) over all masked voxels.  Both accumulate (zero them once).
 * ---------------------------------------------------------------------- */
int pw_occ_confusion(const unsigned char* pred, const unsigned char* gt,
                     const unsigned char* mask, long long n, int n_cl,
                     int free_idx, long long* hist, long long* occ_hist,
                     void* stream);

/* ------------------------------------------------------------------------
 * Voxel SSC training losses: CE_ssc_loss, sem_scal_loss,
 * geo_scal_loss of mmdet3d/models/detectors/loss.py:20-113 as called from
 * preworld.py:151-154 -- one pass over the logits instead of ~60.
 * logits [n_vox, ld] channels-last (any voxel order, the same as target's),
 * target / camera_mask uint8 [n_vox] (camera_mask may be NULL; it applies to
 * the sem and geo terms only, as in the reference), class_weights [n_cls]
 * (the reference appends a 0 for the empty class, preworld.py:150).
 * pw_voxel_loss_stats fills stats[pw_voxel_loss_stats_size(n_cls)] (fp64 sums:
 * CE numerator/denominator, mask count, per class sum p / sum p*hit / count,
 * the five geo sums) and losses[3] = {ce, sem_scal, geo_scal} (unweighted).
 * pw_voxel_loss_grad writes d(w_ce*ce + w_sem*sem + w_geo*geo)/dlogits from the
 * same stats (closed form; the BCE clamp at log = -100 is not differentiated).
 */
int pw_voxel_loss_stats_size(int n_cls);
int pw_voxel_loss_stats(const float* logits, int ld, const unsigned char* target,
                        const unsigned char* camera_mask, long long n_vox, int n_cls,
                        int ignore_index, int empty_idx, const float* class_weights,
                        double* stats, float* losses, void* stream);
int pw_voxel_loss_grad(const float* logits, int ld, const unsigned char* target,
                       const unsigned char* camera_mask, long long n_vox, int n_cls,
                       int ignore_index, int empty_idx, const float* class_weights,
                       const double* stats, float w_ce, float w_sem, float w_geo,
                       float* grad_logits, int grad_ld, void* stream);

/* LSSViewTransformerBEVDepth.get_depth_loss + get_downsampled_gt_depth
 * (view_transformer.py:736-789, sid=False) fused: gt_depth [bn, H, W] lidar depth
 * maps (0 = no return); depth_pred = the D depth probabilities per feature cell,
 * addressed by element strides (image, bin, y, x) so any layout of [bn,D,h,w]
 * works; labels int32 [bn*h*w] receives the bin label of each cell (-1 =
 * background), sums[2] = {sum of BCE over foreground cells, foreground count},
 * loss[0] = weight * sums[0] / max(1, sums[1]).  pw_depth_loss_grad writes
 * d loss / d depth_pred as [bn*h*w, D] (torch's BCE backward formula). */
int pw_depth_loss(const float* gt_depth, int bn, int H, int W, int downsample,
                  const float* depth_pred, long long stride_img, long long stride_d,
                  long long stride_y, long long stride_x, int D, float depth_min,
                  float depth_step, float weight, int* labels, double* sums, float* loss,
                  void* stream);
int pw_depth_loss_grad(const int* labels, int bn, int h, int w, const float* depth_pred,
                       long long stride_img, long long stride_d, long long stride_y,
                       long long stride_x, int D, const double* sums, float weight,
                       float* grad, void* stream);

/* lovasz_softmax (mmdet3d/models/detectors/lovasz_softmax.py:157-239 with
 * classes='present', per_image=False; preworld.py:155 passes ignore = the
 * empty class and the camera mask).  x [n_vox, ld]: class probabilities
 * (is_logits = 0, as the reference function takes them) or logits (is_logits =
 * 1: softmax fused into the first pass).  Kept voxels: target != ignore_label
 * [and camera_mask != 0].  Per class the errors |[t==c] - p_c| are sorted
 * (one cub::DeviceRadixSort of (class, error) keys in the caller's workspace,
 * pw_lovasz_workspace_bytes) and one CTA per present class forms lovasz_grad and
 * the dot product in fp32.  loss[0] = mean over present classes; grad_probas
 * (NULL or [n_vox, n_cls], every element written) = d loss / d probabilities.
 * pw_softmax_backward turns that into d loss / d logits. */
long long pw_lovasz_workspace_bytes(long long n_vox, int n_cls);
int pw_lovasz_softmax(const float* x, int ld, int is_logits, const unsigned char* target,
                      const unsigned char* camera_mask, long long n_vox, int n_cls,
                      int ignore_label, void* workspace, long long workspace_bytes,
                      float* loss, float* grad_probas, void* stream);
int pw_softmax_backward(const float* logits, int ld, const float* grad_probas,
                        long long n_vox, int n_cls, float* grad_logits, void* stream);

/* CustomFocalLoss.forward (mmdet3d/models/loss_utils/focal_loss.py:162-273), the
 * voxel CE term PreWorld uses by default (use_focal_loss=True, preworld.py:43,
 * 146-148): sigmoid focal loss per (voxel, class) -- mmcv-full 1.6.0's
 * sigmoid_focal_loss op restated (third-party, absent from the reference tree;
 * the file's own py_sigmoid_focal_loss is the same function) -- times
 * class_weights[c] * radial[(v / depth) % hw] (the [H,W] centre-distance map of
 * focal_loss.py:197-203; NULL = 1), summed over classes, mean over kept voxels
 * (target != ignore_index [and camera_mask]), times loss_weight.
 * sums[2] = {sum, kept count}; pw_focal_loss_grad writes d loss / d logits
 * [n_vox, n_cls] (zero rows for dropped voxels). */
int pw_focal_loss(const float* logits, int ld, const unsigned char* target,
                  const unsigned char* camera_mask, long long n_vox, int n_cls,
                  int ignore_index, const float* class_weights, const float* radial, int hw,
                  int depth, float gamma, float alpha, float loss_weight, double* sums,
                  float* loss, void* stream);
int pw_focal_loss_grad(const float* logits, int ld, const unsigned char* target,
                       const unsigned char* camera_mask, long long n_vox, int n_cls,
                       int ignore_index, const float* class_weights, const float* radial,
                       int hw, int depth, float gamma, float alpha, float loss_weight,
                       const double* sums, float* grad, void* stream);

/* pts2ray / get_rays (mmdet3d/datasets/ray.py:34-56): n labelled pixels of ONE
 * camera -> rays [n, 16] = [x, y, depth, semantic, origin(3), direction(3), unit
 * view direction(3), rgb(3)], the record NerfHead consumes.  coor [n,2] pixel
 * xy, label_depth / label_seg [n], label_img [n,3], c2w [4,4] row-major
 * (camera -> key ego), cam_intrinsic [3,3] row-major; all fp32 device arrays. */
int pw_pts2ray(const float* coor, const float* label_depth, const float* label_seg,
               const float* label_img, const float* c2w, const float* cam_intrinsic,
               long long n, float* rays, void* stream);

/* ------------------------------------------------------------------------
 * Swin image backbone of the shipped config (mmdet3d/models/backbones/swin.py:
 * 679-976; configs/preworld/nuscenes/bevstereo-occ.py:45-67).  Tokens are the
 * channels-last image [b,h,w,C] (== the reference's [b, h*w, C]); the linears
 * (qkv, proj, FFN, reduction) are pw_conv_* 1x1 convs (act 4 = GELU).
 * ---------------------------------------------------------------------- */
/* torch.nn.LayerNorm over the last dim: y[r,:] = (x[r,:] - mean) * rsqrt(var +
 * eps) * gamma + beta (biased variance), rows [r, c], c % 4 == 0, c <= 2048. */
int pw_layernorm(const float* x, int x_ld, const float* gamma, const float* beta,
                 float eps, float* y, int y_ld, long long rows, int c, void* stream);
/* PatchMerging.forward up to its reduction (swin.py:185-206): y[b,Y,X, s*c + ch]
 * = LayerNorm_{4c}( x[b, 2Y + s/2, 2X + s%2, ch] ), zeros beyond an odd h / w
 * (F.pad, :198-199).  nn.Unfold orders the 4c channels (ch, ky, kx); this kernel
 * orders them (ky, kx, ch): gamma / beta and the columns of the reduction weight
 * are permuted accordingly by the caller.  y: [b,(h+1)/2,(w+1)/2,y_ld]. */
int pw_patch_merge_ln(const float* x, int x_ld, int b, int h, int w, int c,
                      const float* gamma, const float* beta, float eps, float* y,
                      int y_ld, void* stream);
/* ShiftWindowMSA.forward + WindowMSA.forward between the qkv and the proj linear
 * (swin.py:262-300, 364-427): zero padding to a multiple of ws, cyclic shift,
 * window partition, softmax(q*scale k^T + relative position bias + shift mask) v,
 * window reverse, shift back, crop -- as index arithmetic, no copies.
 * qkv [b,h,w,qkv_ld] = q|k|v (c channels each, head-major, head dim 32), qkv_bias
 * [3c] or NULL (k / v of the padding rows, which the reference pads BEFORE the
 * linear), table [heads][(2ws-1)^2] = relative_position_bias_table transposed
 * (entry (dy+ws-1)*(2ws-1) + dx+ws-1 for query - key offset (dy,dx), the layout
 * relative_position_index encodes, :246-252), out [b,h,w,out_ld] (c channels).
 * c == heads*32, ws*ws <= 256, 0 <= shift < ws. */
int pw_window_attention(const float* qkv, int qkv_ld, const float* qkv_bias,
                        const float* table, float* out, int out_ld, int b, int h, int w,
                        int c, int heads, int ws, int shift, float scale, void* stream);

/* ------------------------------------------------------------------------
 * Device-side pixel pipeline of PrepareImageInputs
 * (mmdet3d/datasets/pipelines/loading.py:954-961 img_transform_core: PIL resize,
 * crop, FLIP_LEFT_RIGHT, rotate; :847-854 mmlabNormalize).  Bit-exact with PIL's
 * 8-bit fixed-point resampling; the integer tap tables come from
 * preworld_b200/pixels.py:resample_tables.
 * ---------------------------------------------------------------------- */
/* Horizontal pass of Image.resize: tmp[r, x, c] = clip8((2^21 + sum_k src[row0 + r,
 * first[x] + k, c] * taps[x, k]) >> 22) for r < rows; src uint8 HWC (RGB) with a
 * row pitch in bytes, tmp uint8 [rows, nw, 3].  identity != 0: nw == w, plain copy
 * (PIL skips the pass when the width does not change). */
int pw_resample_rows_u8(const unsigned char* src, long long src_pitch, int h, int w,
                        int row0, int rows, const int* first, const int* count,
                        const int* taps, int ksize, int identity, unsigned char* tmp,
                        int nw, void* stream);
/* Vertical pass + crop + flip + nearest rotation + imnormalize, evaluated at the
 * pixels of the final [fh, fw] view: out[c, y, x] = lut[c][v[2 - c]] (lut [3][256] fp32 =
 * fp32((double(value) - mean[c]) / std[c]-style table of mmcv's imnormalize with to_rgb),
 * v = pixel (crop_x + gx', crop_y + gy) of the resized [nh, nw] image (0 outside: the
 * zero fill of Image.crop / Image.rotate), gx' = flip ? fw - 1 - gx : gx, (gx, gy) =
 * (x, y) or, with affine_fixed (HOST pointer to 6 ints, 16.16 fixed point, as PIL's
 * nearest-neighbour affine transform walks them; NULL = no rotation),
 * ((a2 + y a1 + x a0) >> 16, (a5 + y a4 + x a3) >> 16).  out fp32 [3, fh, fw]. */
int pw_resample_view_norm(const unsigned char* tmp, int row0, int rows, int nw, int nh,
                          const int* first, const int* count, const int* taps, int ksize,
                          int identity, int crop_x, int crop_y, int flip,
                          const int* affine_fixed, const float* lut, float* out, int fh,
                          int fw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PREWORLD_B200_H_ */
