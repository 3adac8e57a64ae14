"""Parity of every C-ABI kernel against the CPU oracle / a plain fp32 torch
reference of the same op, on seeded inputs small enough for the oracle to
finish in seconds.  Integer / index results must be bit-exact; float results
within the tolerance written at each assert."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import c_ref, torch_ref          # noqa: E402  (checker only)
from oracle.cases import RENDER_CASE, render_inputs   # noqa: E402
from preworld_b200 import _lib, ops          # noqa: E402
from preworld_b200 import synthetic as S     # noqa: E402

DEV = 'cuda'


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-6)


# --------------------------------------------------------------------- conv
CONV_CASES = [
    # (dims, n, cin, cout, spatial, k, stride, pad, dil, bias, bn, act, res)
    (2, 2, 3, 64, (32, 40), 7, 2, 3, 1, False, True, 'relu', False),   # stem
    (2, 3, 64, 256, (17, 23), 1, 1, 0, 1, False, True, 'relu', True),
    (2, 2, 128, 128, (16, 20), 3, 2, 1, 1, False, True, 'relu', False),
    (2, 2, 256, 96, (16, 44), 3, 1, 18, 18, False, True, 'relu', False),  # ASPP
    (2, 6, 88, 88, (16, 44), 3, 2, 1, 1, True, True, None, False),  # cvnet
    (2, 2, 344, 256, (8, 11), 1, 1, 0, 1, True, False, None, False),
    (2, 2, 1024, 512, (6, 9), 1, 1, 0, 1, False, True, 'relu', False),
    (3, 1, 32, 32, (8, 20, 24), 3, 1, 1, 1, False, True, 'relu', True),
    (3, 1, 64, 64, (8, 20, 20), 3, 2, 1, 1, False, True, None, False),
    (3, 2, 128, 128, (4, 10, 10), 3, 1, 1, 1, False, True, 'relu', True),
    (3, 1, 32, 16, (6, 10, 12), 3, 1, 1, 1, False, True, 'relu', False),
    (3, 1, 224, 32, (4, 8, 8), 1, 1, 0, 1, False, True, 'relu', False),
    (3, 1, 8, 18, (5, 7, 9), 1, 1, 0, 1, False, False, None, False),
    (3, 1, 32, 32, (5, 9, 11), 3, 1, 1, 1, True, False, 'relu', False),
]


@pytest.fixture(params=['simt', 'umma'])
def conv_path(request):
    """Run a test on the fp32 SIMT kernel and on the tcgen05 3xTF32 kernel
    (the latter is taken whenever cin % 32 == 0)."""
    old = ops.USE_UMMA
    ops.USE_UMMA = request.param == 'umma'
    yield request.param
    ops.USE_UMMA = old


@pytest.mark.parametrize('case', CONV_CASES)
def test_conv_matches_torch_fp32(case, conv_path):
    dims, n, cin, cout, sp, k, stride, pad, dil, bias, bn, act, res = case
    if conv_path == 'umma' and cin % 32:
        pytest.skip('tensor-core path needs cin % 32 == 0')
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    x = torch.randn(n, cin, *sp, generator=g)
    w = torch.randn(cout, cin, *([k] * dims), generator=g) / (cin * k ** dims) ** .5
    b = torch.randn(cout, generator=g) if bias else None
    conv = F.conv2d if dims == 2 else F.conv3d
    want = conv(x, w, b, stride, pad, dil)
    bn_t = None
    if bn:
        gamma, beta = torch.rand(cout, generator=g) + .5, torch.randn(cout, generator=g)
        mean, var = torch.randn(cout, generator=g) * .1, torch.rand(cout, generator=g) + .5
        want = F.batch_norm(want, mean, var, gamma, beta, False, 0., 1e-5)
        bn_t = (gamma.to(DEV), beta.to(DEV), mean.to(DEV), var.to(DEV), 1e-5)
    r = torch.randn(want.shape, generator=g) if res else None
    if res:
        want = want + r
    if act == 'relu':
        want = F.relu(want)
    pc = ops.PackedConv(w.to(DEV), b.to(DEV) if bias else None, bn_t,
                        stride=stride, padding=pad, dilation=dil)
    perm = (0, *range(2, 2 + dims), 1)
    x_cl = x.permute(*perm).contiguous()
    if cin % 4:
        x_cl = F.pad(x_cl, (0, 4 - cin % 4))
    r_cl = r.permute(*perm).contiguous().to(DEV) if res else None
    got = ops.conv(x_cl.to(DEV), pc, act, residual=r_cl)
    got = ops.to_logical(got).cpu()
    assert got.shape == want.shape
    # fp32 accumulation in a different order than the CPU reference; on the
    # tensor-core path the MMA's truncating accumulator leaves ~1e-6 rms / 6e-6
    # max at K = 4608 after the gain correction (tools/accuracy_probe.py)
    assert _rel(got, want) < 8e-6, _rel(got, want)


FOLD_CASES = [
    # (dims, n, cin, cout, spatial, k, pad, dil): stride-1 3-tap convs with
    # cout <= 32 -- the shapes pw_conv_fold_supported() takes by rule
    (3, 1, 32, 32, (7, 13, 45), 3, 1, 1),      # ragged x: 45 = 3*14 + 3
    (3, 2, 64, 32, (4, 9, 31), 3, 1, 1),       # two K chunks, batch 2
    (3, 1, 32, 16, (5, 9, 37), 3, 1, 1),       # fold_n = 16
    (2, 2, 64, 32, (19, 45), 3, 2, 2),         # dilated 2-D
    (2, 1, 96, 24, (11, 29), 3, 1, 1),         # cout not a multiple of 16
]


@pytest.mark.parametrize('case', FOLD_CASES)
def test_conv_fold_matches_torch_and_halo(case):
    """The x-tap-folded launch (pw_conv_fold_fwd: kw taps in the MMA's N
    dimension, shifted row sums in the epilogue) against torch fp32 and
    against the unfolded tensor-core kernel, with affine + residual + ReLU."""
    dims, n, cin, cout, sp, k, pad, dil = case
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    x = torch.randn(n, cin, *sp, generator=g)
    w = torch.randn(cout, cin, *([k] * dims), generator=g) / (cin * k ** dims) ** .5
    b = torch.randn(cout, generator=g)
    r = torch.randn(n, cout, *sp, generator=g)
    conv = F.conv2d if dims == 2 else F.conv3d
    want = F.relu(conv(x, w, b, 1, pad, dil) + r)
    pc = ops.PackedConv(w.to(DEV), b.to(DEV), None, stride=1, padding=pad,
                        dilation=dil)
    assert pc.wf_hi is not None
    perm = (0, *range(2, 2 + dims), 1)
    x_cl = x.permute(*perm).contiguous().to(DEV)
    r_cl = r.permute(*perm).contiguous().to(DEV)
    got = {}
    for fold in (True, False):
        old = ops.USE_FOLD, ops.FOLD_AUTO_COUT
        ops.USE_FOLD, ops.FOLD_AUTO_COUT = fold, 128     # (auto rule: cout <= 16)
        n0 = _lib.launch_count()
        try:
            got[fold] = ops.to_logical(ops.conv(x_cl, pc, 'relu', residual=r_cl)).cpu()
        finally:
            ops.USE_FOLD, ops.FOLD_AUTO_COUT = old
        assert _lib.launch_count() == n0 + 1
    assert _rel(got[True], want) < 8e-6, _rel(got[True], want)
    assert _rel(got[True], got[False]) < 8e-6
    # the folded path really was the one taken
    sp3 = (1,) * (3 - dims) + tuple(sp)
    k3 = (1,) * (3 - dims) + (k,) * dims
    p3 = (0,) * (3 - dims) + (pad,) * dims
    d3 = (1,) * (3 - dims) + (dil,) * dims
    d = ops.ConvDesc(n=n, d=sp3[0], h=sp3[1], w=sp3[2], cin=cin, in_ld=cin,
                     od=sp3[0], oh=sp3[1], ow=sp3[2], cout=cout, out_ld=cout,
                     res_ld=cout, w_ld=pc.w_ld, kd=k3[0], kh=k3[1], kw=k3[2],
                     sd=1, sh=1, sw=1, pd=p3[0], ph=p3[1], pw=p3[2],
                     dd=d3[0], dh=d3[1], dw=d3[2], act=0, act_channels=0)
    import ctypes
    assert _lib.lib().pw_conv_fold_supported(ctypes.byref(d)) == 1


def test_conv_channel_slices_and_split_activation(conv_path):
    """Input / output / residual as channel slices of wider buffers, and the
    fused [relu | linear] double-branch launch used by BasicBlock3D."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 6, 9, 10, 48, generator=g).to(DEV)
    w = (torch.randn(64, 32, 3, 3, 3, generator=g) / 30).to(DEV)
    pc = ops.PackedConv(w, None, None, stride=1, padding=1)
    out = torch.zeros(1, 6, 9, 10, 80, device=DEV)
    ops.conv(x[..., 8:40], pc, 'relu', out=out[..., 16:], act_channels=32)
    # reference on the CPU: torch's CUDA convs default to TF32
    want = F.conv3d(x[..., 8:40].permute(0, 4, 1, 2, 3).cpu(), w.cpu(), None,
                    1, 1)
    want = torch.cat([F.relu(want[:, :32]), want[:, 32:]], 1)
    assert _rel(ops.to_logical(out[..., 16:]).cpu(), want) < 2e-5
    assert (out[..., :16] == 0).all()


def test_linear_and_activations():
    g = torch.Generator().manual_seed(6)
    x = (torch.randn(1000, 32, generator=g) * 4).to(DEV)
    w, b = torch.randn(17, 32, generator=g).to(DEV), torch.randn(17, generator=g).to(DEV)
    pc = ops.PackedConv(w, b)
    for act, fn in (('softplus', F.softplus), ('sigmoid', torch.sigmoid),
                    (None, lambda t: t)):
        got = ops.linear(x, pc, act)
        want = fn(F.linear(x.cpu(), w.cpu(), b.cpu()))
        assert got.shape == (1000, 17)
        assert _rel(got.cpu(), want) < 1e-5


# --------------------------------------------------------------- elementwise
def test_image_side_elementwise():
    g = torch.Generator().manual_seed(7)
    img = torch.randn(5, 3, 18, 26, generator=g)
    frames = img.to(DEV)[1::2]                          # strided images
    y = ops.nchw_to_nhwc(frames, 4)
    assert torch.equal(y[..., :3].cpu(), img[1::2].permute(0, 2, 3, 1))
    assert (y[..., 3] == 0).all()
    x = torch.randn(2, 64, 17, 23, generator=g)
    x_cl = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    assert torch.equal(ops.to_logical(ops.maxpool3x3s2(x_cl)).cpu(),
                       F.max_pool2d(x, 3, 2, 1))
    assert torch.equal(ops.nhwc_to_nchw(x_cl[..., 8:40]).cpu(), x[:, 8:40])
    lo = torch.randn(2, 64, 9, 12, generator=g)
    want = x + F.interpolate(lo, size=(17, 23), mode='nearest')
    y = x_cl.clone()
    ops.upsample_nearest_add_(y, lo.permute(0, 2, 3, 1).contiguous().to(DEV))
    assert torch.equal(ops.to_logical(y).cpu(), want)
    gate = torch.rand(2, 64, generator=g)
    got = ops.scale_channels(x_cl, gate.to(DEV))
    assert torch.equal(ops.to_logical(got).cpu(), x * gate[:, :, None, None])
    got = ops.global_avgpool(x_cl).cpu()
    assert _rel(got, x.mean((2, 3))) < 1e-6
    out = torch.zeros(2, 17, 23, 80, device=DEV)
    ops.broadcast_channels_(out[..., 16:], gate.to(DEV))
    assert torch.equal(out[..., 16:].cpu(),
                       gate[:, None, None, :].expand(2, 17, 23, 64))
    logits = torch.randn(3, 16, 11, 120, generator=g) * 3
    got = ops.softmax_depth(logits.to(DEV), 88).cpu()
    want = logits[..., :88].permute(0, 3, 1, 2).softmax(1)
    assert got.shape == (3, 88, 16, 11) and _rel(got, want) < 1e-6


def test_voxel_side_elementwise():
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 64, 4, 10, 12, generator=g)
    for s in (2, 4):
        want = F.interpolate(x, scale_factor=s, mode='trilinear',
                             align_corners=True)
        out = torch.zeros(2, 4 * s, 10 * s, 12 * s, 96, device=DEV)
        ops.upsample_trilinear_(out[..., 32:],
                                x.permute(0, 2, 3, 4, 1).contiguous().to(DEV))
        assert _rel(ops.to_logical(out[..., 32:]).cpu(), want) < 1e-6
        assert (out[..., :32] == 0).all()
    logits = torch.randn(1, 5, 7, 9, 18, generator=g)       # [1,Z,Y,X,18]
    logits[0, 0, 0, 0, 3] = logits[0, 0, 0, 0, 11] = 9.0    # tie -> first
    occ = ops.argmax_zyx_to_xyz(logits.to(DEV)).cpu()
    want = logits[0].argmax(-1).permute(2, 1, 0)            # [X,Y,Z]
    assert occ.dtype == torch.uint8 and torch.equal(occ.long(), want)
    assert occ[0, 0, 0] == 3
    # preworld.py:205-221: geo_occ = num_classes-1 where the class is 17 else 0
    logits[0, 1, 2, 3, 17] = 50.0
    both = ops.argmax_geo_zyx_to_xyz(logits.to(DEV), 17, 17).cpu()
    want = logits[0].argmax(-1).permute(2, 1, 0)
    want_geo = torch.ones_like(want) * 17
    want_geo[want != 17] = 0
    assert both.shape == (2, 9, 7, 5) and both.dtype == torch.uint8
    assert torch.equal(both[0].long(), want) and torch.equal(both[1].long(), want_geo)
    assert both[1, 3, 2, 1] == 17 and (both[1] == 17).sum() >= 1
    dens = torch.rand(1, 5, 7, 9, 2, generator=g) * 17
    occ, geo = ops.density_occ_zyx_to_xyz(dens.to(DEV)[..., 0:1],
                                          logits.to(DEV)[..., :17], 8.5, 17)
    ne = (dens[0, ..., 0] > 8.5).permute(2, 1, 0)
    want = torch.where(ne, logits[0, ..., :17].argmax(-1).permute(2, 1, 0), 17)
    assert torch.equal(occ.cpu().long(), want)
    assert torch.equal(geo.cpu().long(), torch.where(ne, 0, 17))
    v = torch.randn(2, 5, 7, 9, 32, generator=g)
    assert torch.equal(ops.zyx_to_xyz(v.to(DEV)).cpu(),
                       v.permute(0, 3, 2, 1, 4))


# ----------------------------------------------------------------------- lift
def _lift_setup(batch=1, input_size=(64, 176), seed=0, grid=None):
    from oracle.cases import TINY_GRID
    geo = torch_ref.LiftGeometry(grid or TINY_GRID, input_size, 16, 32)
    inputs = S.make_img_inputs(batch, input_size, seed=seed)
    pi = torch_ref.prepare_inputs(inputs)
    s2k, intr, pr, pt, bda = pi[1][0], pi[3][0], pi[4][0], pi[5][0], pi[6]
    return geo, s2k, intr, pr, pt, bda


def test_bev_pool_v2_drop_in_kat():
    """The reference's own known-answer test (bev_pool.py:145-176) through
    the C ABI."""
    depth = torch.tensor([0.3, 0.4, 0.2, 0.1, 0.7, 0.6, 0.8, 0.9], device=DEV)
    feat = torch.ones(4, 2, device=DEV)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
    out = torch.zeros(4, 2, device=DEV)
    ops.bev_pool_v2_(depth, feat, i32([0, 4, 1, 6]), i32([0, 0, 1, 2]),
                     i32([0, 0, 1, 1]), i32([0, 2]), i32([2, 2]), out)
    assert abs(out.sum().item() - 4.4) < 1e-6
    assert torch.allclose(out.cpu(), torch.tensor(
        [[1., 1.], [1.2, 1.2], [0, 0], [0, 0]]), atol=1e-6)
    ops.bev_pool_v2_(depth, feat, i32([]), i32([]), i32([]), i32([]), i32([]),
                     out)                      # empty: no-op


def test_bev_pool_v2_grad_kat_and_oracle():
    """Backward drop-in (bev_pool.py:43-83, src/bev_pool_cuda.cu:67-121): the
    reference's own gradient KAT (bev_pool.py:170-176) through the C ABI, then
    a random case bit-exact against the C oracle."""
    def intervals(rf):
        kept = np.ones(len(rf), bool)
        kept[1:] = rf[1:] != rf[:-1]
        st = np.where(kept)[0].astype(np.int32)
        return st, np.diff(np.append(st, len(rf))).astype(np.int32)

    def run(out_grad, depth, feat, rd, rf, rb):
        order = np.argsort(rf, kind='stable')
        rd, rf, rb = rd[order], rf[order], rb[order]
        st, ln = intervals(rf)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
        dg = torch.zeros(depth.size, device=DEV)
        fg = torch.zeros(feat.shape, device=DEV)
        ops.bev_pool_v2_grad_(t(out_grad), t(depth), t(feat), t(rd), t(rf),
                              t(rb), t(st), t(ln), dg, fg)
        want = c_ref.bev_pool_v2_bwd(out_grad, depth, feat, rd, rf, rb, st, ln)
        return dg.cpu().numpy(), fg.cpu().numpy(), want

    depth = np.array([0.3, 0.4, 0.2, 0.1, 0.7, 0.6, 0.8, 0.9], np.float32)
    dg, fg, _ = run(np.ones((4, 2), np.float32), depth, np.ones((4, 2), np.float32),
                    np.array([0, 4, 1, 6], np.int32), np.array([0, 0, 1, 2], np.int32),
                    np.array([0, 0, 1, 1], np.int32))
    np.testing.assert_allclose(dg, [2, 2, 0, 0, 2, 0, 2, 0], atol=1e-6)
    np.testing.assert_allclose(fg.ravel(), [1, 1, .4, .4, .8, .8, 0, 0], atol=1e-6)

    rng = np.random.default_rng(0)
    n_pts, n_depth, n_feat, n_bev, c = 20000, 30000, 900, 5000, 80
    rd = rng.choice(n_depth, n_pts, replace=False).astype(np.int32)
    rf = rng.integers(0, n_feat, n_pts).astype(np.int32)
    rb = rng.integers(0, n_bev, n_pts).astype(np.int32)
    dg, fg, (wdg, wfg) = run(rng.standard_normal((n_bev, c)).astype(np.float32),
                             rng.random(n_depth).astype(np.float32),
                             rng.standard_normal((n_feat, c)).astype(np.float32),
                             rd, rf, rb)
    assert np.array_equal(dg, wdg)           # same sequential fmaf chains
    assert np.array_equal(fg, wfg)


def test_render_backward_primitives_match_c_oracle():
    """raw2alpha_backward / alpha2weight_backward (render_utils_kernel.cu:
    507-517, 654-676) against the C restatements."""
    g = torch.Generator().manual_seed(5)
    dens = torch.rand(6000, generator=g) * 40 - 5
    gb = torch.randn(6000, generator=g)
    e, a = ops.raw2alpha(dens.to(DEV), -13.8155, 0.5)
    got = ops.raw2alpha_backward(e, gb.to(DEV), 0.5).cpu().numpy()
    want = c_ref.raw2alpha_bwd(e.cpu().numpy(), gb.numpy(), 0.5)
    np.testing.assert_allclose(got, want, rtol=3e-6, atol=1e-12)   # powf ulp
    n_rays = 500
    counts = torch.randint(0, 60, (n_rays,), generator=g)
    counts[::13] = 0
    ray_id = torch.repeat_interleave(torch.arange(n_rays), counts)
    alpha = torch.rand(len(ray_id), generator=g) ** 2
    w, T, last, i_s, i_e = ops.alpha2weight(alpha.to(DEV), ray_id.to(DEV), n_rays)
    gw = torch.randn(len(ray_id), generator=g)
    gl = torch.randn(n_rays, generator=g)
    got = ops.alpha2weight_backward(alpha.to(DEV), w, T, last, i_s, i_e,
                                    gw.to(DEV), gl.to(DEV)).cpu().numpy()
    want = c_ref.alpha2weight_bwd(alpha.numpy(), w.cpu().numpy(), T.cpu().numpy(),
                                  last.cpu().numpy(), i_s.cpu().numpy(),
                                  i_e.cpu().numpy(), gw.numpy(), gl.numpy())
    assert np.array_equal(got, want)


@pytest.mark.parametrize('batch', [1, 2])
def test_lift_matches_oracle_bit_exact(batch):
    geo, s2k, intr, pr, pt, bda = _lift_setup(batch, seed=3)
    B, N = s2k.shape[:2]
    D, H, W = geo.frustum.shape[:3]
    xs, ys, ds = geo.frustum[0, 0, :, 0], geo.frustum[0, :, 0, 1], geo.frustum[:, 0, 0, 2]
    grid = tuple(int(v) for v in geo.grid_size)
    # camera tables: device kernel == C oracle, bit for bit
    cam_ref = c_ref.lift_camera_params(s2k.numpy(), intr.numpy(), pr.numpy(),
                                       pt.numpy())
    cam = ops.lift_camera_params(s2k.to(DEV), intr.to(DEV), pr.to(DEV),
                                 pt.to(DEV))
    assert np.array_equal(cam.cpu().numpy(), cam_ref)
    # voxel ranks: bit exact vs the C oracle
    rank_ref = c_ref.lift_ranks(B, N, xs.numpy(), ys.numpy(), ds.numpy(),
                                cam_ref, bda.numpy(), geo.lower.numpy(),
                                geo.interval.numpy(), grid)
    rank = ops.lift_ranks(cam, bda.reshape(B, 9).to(DEV), xs.to(DEV),
                          ys.to(DEV), ds.to(DEV), geo.lower.tolist(),
                          geo.interval.tolist(), B, N, grid)
    assert np.array_equal(rank.cpu().numpy(), rank_ref)
    assert 0.05 < (rank_ref >= 0).mean() < 0.9
    # ... and they agree with the reference's torch geometry except for
    # points within fp32 rounding of a voxel face
    coor = torch_ref.get_lidar_coor(geo, s2k, intr, pr, pt, bda)
    c = ((coor - geo.lower) / geo.interval).long().view(-1, 3)
    ok = ((c >= 0) & (c < geo.grid_size.long())).all(1)
    b_idx = torch.arange(B).repeat_interleave(N * D * H * W)
    r_t = torch.where(ok, ((b_idx * grid[2] + c[:, 2]) * grid[1] + c[:, 1])
                      * grid[0] + c[:, 0], -1)
    assert (r_t.numpy() != rank_ref).mean() < 1e-4
    # pooled volume: same points, same order, same fmaf -> bit exact
    g = torch.Generator().manual_seed(1)
    depth = torch.rand(B * N, D, H, W, generator=g).softmax(1)
    feat = torch.randn(B * N, H, W, 40, generator=g)
    valid = np.where(rank_ref >= 0)[0]
    order = valid[np.argsort(rank_ref[valid], kind='stable')]
    rb = rank_ref[order].astype(np.int32)
    rd = order.astype(np.int32)
    hw = H * W
    rf = ((order // (D * hw)) * hw + order % hw).astype(np.int32)
    kept = np.ones(len(rb), bool)
    kept[1:] = rb[1:] != rb[:-1]
    st = np.where(kept)[0].astype(np.int32)
    ln = np.diff(np.append(st, len(rb))).astype(np.int32)
    nvox = B * grid[0] * grid[1] * grid[2]
    want = c_ref.bev_pool_v2_fwd(depth.numpy().ravel(),
                                 feat[..., 4:36].reshape(-1, 32).numpy(), rd,
                                 rf, rb, st, ln, nvox)
    got = ops.lift_fused(depth.to(DEV), feat.to(DEV)[..., 4:36], cam,
                         bda.reshape(B, 9).to(DEV), xs.to(DEV), ys.to(DEV),
                         ds.to(DEV), geo.lower.tolist(), geo.interval.tolist(),
                         B, N, grid)
    assert got.shape == (B, grid[2], grid[1], grid[0], 32)
    assert np.array_equal(got.cpu().numpy().reshape(nvox, 32), want)
    assert ln.max() > 1                       # multi-point voxels exercised
    # the drop-in kernel on the same ranks/intervals: also bit exact
    out = torch.zeros(nvox, 32, device=DEV)
    t = lambda a: torch.from_numpy(a).to(DEV)
    ops.bev_pool_v2_(depth.to(DEV), feat[..., 4:36].contiguous().to(DEV),
                     t(rd), t(rf), t(rb), t(st), t(ln), out)
    assert np.array_equal(out.cpu().numpy(), want)
    # ... and so is the REFERENCE's own CUDA kernel (bev_pool_cuda.cu:21-48 compiled
    # unmodified for sm_100a by oracle/build_ref.sh): a second, GPU-side oracle
    from oracle import gpu_ref
    if gpu_ref.available():
        out_r = torch.zeros(nvox, 32, device=DEV)
        gpu_ref.bev_pool_v2_forward(depth.to(DEV), feat[..., 4:36].contiguous().to(DEV), out_r,
                                    t(rd), t(rf), t(rb), t(ln), t(st))
        torch.cuda.synchronize()
        assert np.array_equal(out_r.cpu().numpy(), want)
    # linearity in the features (size-independent property)
    got2 = ops.lift_fused(depth.to(DEV), (feat * 2).to(DEV)[..., 4:36], cam,
                          bda.reshape(B, 9).to(DEV), xs.to(DEV), ys.to(DEV),
                          ds.to(DEV), geo.lower.tolist(),
                          geo.interval.tolist(), B, N, grid)
    assert torch.equal(got2, got * 2)


def test_lift_prepare_pool_equals_fused():
    """accelerate=True split (pw_lift_prepare once + pw_lift_pool per call,
    view_transformer.py:155-174,263-295) gives the bits of pw_lift_fused, also
    on a second depth/feature pair pooled with the same lists."""
    geo, s2k, intr, pr, pt, bda = _lift_setup(1, seed=5)
    B, N = s2k.shape[:2]
    D, H, W = geo.frustum.shape[:3]
    xs, ys, ds = geo.frustum[0, 0, :, 0], geo.frustum[0, :, 0, 1], geo.frustum[:, 0, 0, 2]
    grid = tuple(int(v) for v in geo.grid_size)
    cam = ops.lift_camera_params(s2k.to(DEV), intr.to(DEV), pr.to(DEV), pt.to(DEV))
    args = (cam, bda.reshape(B, 9).to(DEV), xs.to(DEV), ys.to(DEV), ds.to(DEV),
            geo.lower.tolist(), geo.interval.tolist(), B, N, grid)
    ws = ops.lift_prepare(*args)
    g = torch.Generator().manual_seed(2)
    for _ in range(2):
        depth = torch.rand(B * N, D, H, W, generator=g).softmax(1).to(DEV)
        feat = torch.randn(B * N, H, W, 40, generator=g).to(DEV)[..., 4:36]
        want = ops.lift_fused(depth, feat, *args)
        got = ops.lift_pool(depth, feat, ws, B, N, grid)
        assert torch.equal(got, want)
        assert (want != 0).any()


def test_lift_all_points_outside_grid():
    geo, s2k, intr, pr, pt, bda = _lift_setup(1)
    far = dict(geo.grid_config, x=[500, 516, 0.4])
    geo2 = torch_ref.LiftGeometry(far, (64, 176), 16, 32)
    D, H, W = geo2.frustum.shape[:3]
    xs, ys, ds = geo2.frustum[0, 0, :, 0], geo2.frustum[0, :, 0, 1], geo2.frustum[:, 0, 0, 2]
    cam = ops.lift_camera_params(s2k.to(DEV), intr.to(DEV), pr.to(DEV), pt.to(DEV))
    depth = torch.rand(6, D, H, W).to(DEV)
    feat = torch.randn(6, H, W, 32).to(DEV)
    grid = tuple(int(v) for v in geo2.grid_size)
    got = ops.lift_fused(depth, feat, cam, bda.reshape(1, 9).to(DEV),
                         xs.to(DEV), ys.to(DEV), ds.to(DEV),
                         geo2.lower.tolist(), geo2.interval.tolist(), 1, 6, grid)
    assert (got == 0).all()                   # view_transformer.py:238-246


# ---------------------------------------------------------------- cost volume
def test_cost_volume_matches_oracle():
    torch.manual_seed(0)
    inputs = S.make_img_inputs(1, (64, 176), seed=9)
    pi = torch_ref.prepare_inputs(inputs)
    geo = torch_ref.LiftGeometry(
        {'x': [-8, 8, .4], 'y': [-8, 8, .4], 'z': [-1, 5.4, .4],
         'depth': [1., 45., .5]}, (64, 176), 16, 32)
    g = torch.Generator().manual_seed(2)
    curr = F.relu(torch.randn(6, 256, 16, 44, generator=g))
    prev = F.relu(torch.randn(6, 256, 16, 44, generator=g))   # exact zeros
    metas = dict(k2s_sensor=pi[7][0], intrins=pi[3][0], post_rots=pi[4][0],
                 post_trans=pi[5][0], frustum=geo.cv_frustum,
                 cv_feat_list=[prev, curr])
    want = torch_ref.calculate_cost_volume(metas, 5.0)        # [6,88,16,44]
    fr = geo.cv_frustum
    cam = ops.cv_camera_params(pi[7][0].to(DEV), pi[3][0].to(DEV),
                               pi[4][0].to(DEV), pi[5][0].to(DEV))
    cl = lambda t: t.permute(0, 2, 3, 1).contiguous().to(DEV)
    got = ops.cost_volume(cl(curr), cl(prev), cam, fr[0, 0, :, 0].to(DEV),
                          fr[0, :, 0, 1].contiguous().to(DEV),
                          fr[:, 0, 0, 2].contiguous().to(DEV), 5.0, (64, 176))
    got = got.permute(0, 3, 1, 2).cpu()
    assert got.shape == want.shape
    # softmax probabilities; the sampling grid differs from torch's matmul
    # chain by fp32 rounding -> bilinear weights by ~1e-5
    assert (got - want).abs().max().item() < 2e-4
    assert torch.allclose(got.sum(1), torch.ones(6, 16, 44), atol=1e-5)


# ---------------------------------------------------------------------- render
def test_render_primitives_match_c_oracle():
    g = torch.Generator().manual_seed(3)
    dens = torch.rand(5000, generator=g) * 40 - 5
    e, a = ops.raw2alpha(dens.to(DEV), -13.8155, 0.5)
    want = c_ref.raw2alpha(dens.numpy(), -13.8155, 0.5)
    # 1 - (1+e)^-0.5 cancels for tiny e: one ulp of powf (CUDA vs glibc) is
    # 6e-8 absolute
    np.testing.assert_allclose(a.cpu().numpy(), want, rtol=2e-6, atol=1.2e-7)
    n_rays = 300
    counts = torch.randint(0, 40, (n_rays,), generator=g)
    counts[::17] = 0
    ray_id = torch.repeat_interleave(torch.arange(n_rays), counts)
    alpha = torch.rand(len(ray_id), generator=g) ** 3
    w, T, last, i_s, i_e = ops.alpha2weight(alpha.to(DEV), ray_id.to(DEV), n_rays)
    rw, rT, rlast, ris, rie = c_ref.alpha2weight(alpha.numpy(), ray_id.numpy(),
                                                 n_rays, full=True)
    assert np.array_equal(w.cpu().numpy(), rw)           # same rounding chain
    assert np.array_equal(T.cpu().numpy(), rT)
    assert np.array_equal(last.cpu().numpy(), rlast)
    assert np.array_equal(i_s.cpu().numpy(), ris)
    assert np.array_equal(i_e.cpu().numpy(), rie)
    assert (rie - ris < counts.numpy()).any()            # early stops occurred
    dist = torch.rand(64, 416, generator=g) * 0.01
    m = ops.cumdist_thres(dist.to(DEV), 0.0049)
    assert np.array_equal(m.cpu().numpy(),
                          c_ref.cumdist_thres(dist.numpy(), 0.0049))
    # empty inputs
    w, T, last, i_s, i_e = ops.alpha2weight(alpha[:0].to(DEV),
                                            ray_id[:0].to(DEV), 4)
    assert last.tolist() == [1, 1, 1, 1] and i_e.tolist() == [0, 0, 0, 0]


@pytest.mark.parametrize('library_order', [False, True])
def test_render_rays_matches_reference_fixture(golden_dir, library_order):
    import os
    from preworld_b200.plugin.heads import NerfHead
    fx = np.load(os.path.join(golden_dir, 'render.npz'))
    rays, bda, density, semantic, color = render_inputs(RENDER_CASE)
    head = NerfHead([-40., -40., -1., 40., 40., 5.4], 0.4).to(DEV)
    if library_order:
        attr = torch.zeros(16, 200, 200, 24)
        attr[..., 0] = density[0].permute(2, 1, 0)
        attr[..., 2:19] = semantic[0].permute(2, 1, 0, 3)
        attr[..., 19:22] = color[0].permute(2, 1, 0, 3)
        attr = attr.to(DEV)
        out = head.render(attr[..., 0:1], attr[..., 2:19], attr[..., 19:22],
                          rays[0].to(DEV), bda[0].to(DEV), library_order=True)
    else:
        out = head.render(density[0].to(DEV), semantic[0].to(DEV),
                          color[0].to(DEV), rays[0].to(DEV), bda[0].to(DEV))
    mask = out['ray_mask'].cpu().numpy()
    assert np.array_equal(mask, fx['ray_mask'])
    for k in ('render_depth', 'render_semantic', 'render_color',
              'alphainv_last'):
        got = out[k].cpu().numpy()[mask]
        scale = max(1.0, float(np.abs(fx[k]).max()))
        err = np.abs(got - fx[k]).max() / scale
        # tolerance: 1e-3 relative (north_star); observed ~1e-5 (sample
        # positions differ from torch's by fp32 rounding of the norm)
        assert err < 1e-3, (k, err)
        assert np.median(np.abs(got - fx[k])) / scale < 1e-5, k
    assert (out['render_depth'].cpu().numpy()[~mask] == 0).all()


def test_device_miou_matches_oracle_bit_exact():
    """preworld_b200.metrics.Metric_mIoU (pw_occ_confusion) vs the numpy
    restatement of occ_metrics.py:93-185: integer confusion matrices identical,
    over several samples, with and without the camera mask, 255 = unlabelled."""
    from oracle.metrics_ref import MetricRef
    from preworld_b200.metrics import Metric_mIoU
    rng = np.random.default_rng(1)
    for use_mask in (False, True):
        ref = MetricRef(use_image_mask=use_mask)
        dev = Metric_mIoU(use_image_mask=use_mask)
        for _ in range(3):
            pred = rng.integers(0, 18, (200, 200, 16)).astype(np.uint8)
            gt = rng.integers(0, 18, (200, 200, 16)).astype(np.uint8)
            gt[rng.random(gt.shape) < 0.05] = 255
            mask = rng.random(gt.shape) < 0.6
            ref.add_batch(pred, gt, None, mask)
            dev.add_batch(torch.from_numpy(pred).to(DEV), torch.from_numpy(gt).to(DEV),
                          None, torch.from_numpy(mask).to(DEV))
        assert np.array_equal(dev.hist, ref.hist)
        assert np.array_equal(dev.occ_hist, ref.occ_hist)
        assert dev.count_miou()[3] == ref.count_miou()[1]
        assert dev.count_iou()[3] == ref.count_iou()[1]


def test_resnet_stem_space_to_depth_matches_conv7x7():
    """The tensor-core stem (space-to-depth(2) image, 4x4 stride-1 conv with
    re-indexed weights) reproduces conv1(7x7, s2, p3) + bn1 + relu + maxpool of
    mmdet ResNet (call site detectors/bevdet.py:577-588)."""
    from preworld_b200.plugin.image import ResNet
    torch.manual_seed(0)
    bb = ResNet(depth=50, out_indices=(0, 2, 3)).eval()
    S.lively_init_(bb, 3)
    img = torch.randn(3, 3, 64, 96)
    with torch.no_grad():
        want = F.max_pool2d(F.relu(bb.bn1(bb.conv1(img))), 3, 2, 1)
    bb = bb.to(DEV)
    assert bb.packs()['stem_s2d'] is not None
    assert bb.stem_input_shape(64, 96) == (32, 48, 32)
    with torch.no_grad():
        got = ops.to_logical(bb.run_stem(img.to(DEV))).cpu()
    assert got.shape == want.shape
    err = (got - want).abs().max().item() / want.abs().max().item()
    assert err < 2e-5, err
    # odd image sizes fall back to the plain NHWC stem
    img2 = torch.randn(1, 3, 33, 47)
    with torch.no_grad():
        want2 = F.max_pool2d(F.relu(bb.cpu().bn1(bb.conv1(img2))), 3, 2, 1)
        got2 = ops.to_logical(bb.to(DEV).run_stem(img2.to(DEV))).cpu()
    assert (got2 - want2).abs().max().item() / want2.abs().max().item() < 2e-5


def test_device_temporal_miou_matches_oracle():
    """Metric_mIoU_Temporal (occ_metrics.py:413-596): per-horizon matrices."""
    from oracle.metrics_ref import MetricTemporalRef
    from preworld_b200.metrics import Metric_mIoU_Temporal
    rng = np.random.default_rng(2)
    ref = MetricTemporalRef(use_image_mask=True)
    dev = Metric_mIoU_Temporal(use_image_mask=True)
    for _ in range(2):
        preds = [rng.integers(0, 18, (40, 40, 16)).astype(np.uint8) for _ in range(4)]
        gts, ml, mc = {}, {}, {}
        for idx in (0, 2, 4, 6):
            g = rng.integers(0, 18, (40, 40, 16)).astype(np.uint8)
            g[rng.random(g.shape) < 0.05] = 255
            gts[idx], ml[idx] = g, rng.random(g.shape) < 0.5
            mc[idx] = rng.random(g.shape) < 0.6
        ref.add_batch(preds, gts, ml, mc)
        t = lambda a: torch.from_numpy(a).to(DEV)
        dev.add_batch([t(p) for p in preds], {k: t(v) for k, v in gts.items()},
                      {k: t(v) for k, v in ml.items()},
                      {k: t(v) for k, v in mc.items()})
    for k, idx in enumerate((0, 2, 4, 6)):
        assert np.array_equal(getattr(dev, f'hist_{k}s'), ref.m[idx].hist)
        assert np.array_equal(getattr(dev, f'occ_hist_{k}s'), ref.m[idx].occ_hist)
    assert dev.count_miou()[1] == ref.count_miou()
    assert dev.count_iou() == ref.count_iou()


def test_copy_rows_splits_the_camera_major_batch():
    """ops.copy_rows_: one pitched async copy takes frame f's images out of the
    loader's camera-major [B, N*T, C, H, W] batch (bevdet_occ.py:88-97), from
    pinned host memory and from device memory."""
    g = torch.Generator().manual_seed(3)
    bn, nf = 6, 3
    raw = torch.randn(1, bn * nf, 3, 16, 44, generator=g)
    for src_t in (raw.pin_memory(), raw.to(DEV)):
        src = src_t.view(bn, nf, 3, 16, 44)
        for f in range(nf):
            dst = torch.zeros(bn, 3, 16, 44, device=DEV)
            ops.copy_rows_(dst[2:6], src[2:6, f])
            torch.cuda.synchronize()
            want = raw.view(bn, nf, 3, 16, 44)[:, f]
            assert torch.equal(dst[2:6].cpu(), want[2:6]) and (dst[:2] == 0).all()
    side = torch.cuda.Stream()
    dst = torch.zeros(bn, 3, 16, 44, device=DEV)
    ops.copy_rows_(dst[:1], raw.pin_memory().view(bn, nf, 3, 16, 44)[:1, 1], stream=side)
    side.synchronize()
    assert torch.equal(dst[0].cpu(), raw.view(bn, nf, 3, 16, 44)[0, 1])


@pytest.mark.parametrize('seed,use_mask', [(0, False), (1, True), (2, True), (3, False)])
def test_voxel_losses_match_oracle(seed, use_mask):
    """csrc/losses.cu (one pass: CE + sem_scal + geo_scal of loss.py:20-113, and
    their closed-form gradients) against the CPU restatement and its autograd.
    fp64 accumulation on the device vs fp32 torch sums: 2e-5 relative."""
    from oracle import loss_ref
    from preworld_b200 import losses
    pred, target, cam, cw = loss_ref.seeded_case(seed)
    cwz = torch.cat([cw, torch.zeros(1)])
    m = cam if use_mask else None
    p_ref = pred.clone().requires_grad_(True)
    want = dict(ce=loss_ref.ce_ssc_loss(p_ref, target, cwz, 255),
                sem=loss_ref.sem_scal_loss(p_ref, target, 255, m),
                geo=loss_ref.geo_scal_loss(p_ref, target, 255, 17, m))
    # channels-last logits, as OccHead produces them
    p_dev = pred.to(DEV).permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3) \
        .requires_grad_(True)
    got = losses.voxel_loss_terms(p_dev, target.to(DEV), cwz.to(DEV), 255, 17,
                                  None if m is None else m.to(DEV))
    for k in want:
        assert abs(float(got[k]) - float(want[k])) <= 2e-5 * abs(float(want[k])), k
    wts = dict(ce=0.7, sem=1.3, geo=2.0)
    sum(wts[k] * want[k] for k in want).backward()
    sum(wts[k] * got[k] for k in got).backward()
    g_ref, g_dev = p_ref.grad, p_dev.grad.cpu()
    assert (g_ref - g_dev).abs().max() <= 1e-4 * g_ref.abs().max()
    # the reference-named wrappers and the NCDHW-contiguous input route
    lv = losses.loss_voxel(pred.to(DEV), target.to(DEV), cw.to(DEV), 17,
                           None if m is None else m.to(DEV), 1.0, 1.0, 1.0,
                           use_focal_loss=False)
    assert abs(float(lv['loss_voxel_sem']) - float(want['sem'])) <= 2e-5 * abs(float(want['sem']))
    assert abs(float(lv['loss_voxel_ce']) - float(want['ce'])) <= 2e-5 * abs(float(want['ce']))
    assert abs(float(lv['loss_voxel_geo']) - float(want['geo'])) <= 2e-5 * abs(float(want['geo']))


def test_voxel_losses_full_grid_and_golden(golden_dir):
    """200x200x16 grid (BASELINE size): device losses == oracle; and the
    committed reference values (tests/golden/voxel_losses.json) directly."""
    import json
    import os
    from oracle import loss_ref
    from preworld_b200 import losses
    gold = json.load(open(os.path.join(golden_dir, 'voxel_losses.json')))
    pred, target, cam, cw = loss_ref.seeded_case(0)
    cwz = torch.cat([cw, torch.zeros(1)])
    got = losses.voxel_loss_terms(pred.to(DEV), target.to(DEV), cwz.to(DEV), 255, 17,
                                  cam.to(DEV))
    want = gold['seed0_mask1']
    for k in want:
        assert abs(float(got[k]) - want[k]) <= 2e-5 * abs(want[k]), k
    pred, target, cam, cw = loss_ref.seeded_case(5, shape=(1, 18, 200, 200, 16))
    cwz = torch.cat([cw, torch.zeros(1)])
    got = losses.voxel_loss_terms(pred.to(DEV), target.to(DEV), cwz.to(DEV), 255, 17,
                                  cam.to(DEV))
    assert abs(float(got['ce']) - float(loss_ref.ce_ssc_loss(pred, target, cwz, 255))) < 1e-4
    assert abs(float(got['sem']) - float(loss_ref.sem_scal_loss(pred, target, 255, cam))) < 1e-4
    assert abs(float(got['geo']) - float(loss_ref.geo_scal_loss(pred, target, 255, 17, cam))) < 1e-4


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_depth_loss_matches_oracle(seed, golden_dir):
    """pw_depth_loss / pw_depth_loss_grad (view_transformer.py:736-789 fused)
    against the CPU restatement: bin labels bit-exact, loss 1e-5 relative,
    gradient vs autograd 1e-4 of its maximum; both memory layouts of the
    prediction tensor; and the reference's own value from the golden file."""
    import json
    import os
    from oracle import loss_ref
    from preworld_b200 import losses
    cfg = [1.0, 45.0, 0.5]
    gt, preds = loss_ref.seeded_depth_case(seed)
    p_ref = preds.clone().requires_grad_(True)
    want = loss_ref.depth_loss(gt, p_ref, 16, cfg, 88, 3.0)
    want.backward()
    want_lab = loss_ref.downsampled_gt_depth(gt, 16, cfg, 88)
    gold = json.load(open(os.path.join(golden_dir, 'voxel_losses.json')))[f'depth_seed{seed}']
    for channels_last in (False, True):
        p_dev = preds.to(DEV)
        if channels_last:
            p_dev = p_dev.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        p_dev = p_dev.requires_grad_(True)
        got = losses.get_depth_loss(gt.to(DEV), p_dev, 16, cfg, 3.0)
        assert abs(float(got) - float(want)) <= 1e-5 * float(want)
        assert abs(float(got) - gold['loss']) <= 1e-5 * gold['loss']
        got.backward()
        assert (p_dev.grad.cpu() - p_ref.grad).abs().max() <= 1e-4 * p_ref.grad.abs().max()
    _, labels, sums = ops.depth_loss(gt.reshape(6, 64, 176).to(DEV), preds.to(DEV), 16,
                                     cfg[0], cfg[2], 3.0)
    lab_ref = torch.where(want_lab.sum(1) > 0, want_lab.argmax(1), torch.full((1,), -1))
    assert torch.equal(labels.cpu().long(), lab_ref)
    assert int(sums[1].item()) == gold['n_fg']


@pytest.mark.parametrize('seed,use_mask', [(0, True), (1, False), (2, True)])
def test_lovasz_softmax_matches_oracle(seed, use_mask, golden_dir):
    """pw_lovasz_softmax (keys -> segmented sort -> per-class Jaccard scan) against
    the CPU restatement of lovasz_softmax.py:157-239 and the reference's own
    values; gradient w.r.t. the logits against autograd (random data: no ties)."""
    import json
    import os
    from oracle import loss_ref
    from preworld_b200 import losses
    gold = json.load(open(os.path.join(golden_dir, 'voxel_losses.json')))
    pred, target, cam, cw = loss_ref.seeded_case(seed)
    m = cam if use_mask else None
    p_ref = pred.clone().requires_grad_(True)
    want = loss_ref.lovasz_softmax(torch.softmax(p_ref, 1), target, 17, m)
    want.backward()
    md = None if m is None else m.to(DEV)
    # (a) logits in, fused softmax
    p_dev = pred.to(DEV).requires_grad_(True)
    got = losses.lovasz_softmax(p_dev, target.to(DEV), ignore=17, camera_mask=md,
                                from_logits=True)
    assert abs(float(got) - float(want)) <= 1e-5 * float(want)
    assert abs(float(got) - gold[f'lovasz_seed{seed}_mask{int(use_mask)}']) <= 1e-5 * float(want)
    got.backward()
    assert (p_dev.grad.cpu() - p_ref.grad).abs().max() <= 1e-4 * p_ref.grad.abs().max()
    # (b) probabilities in, as the reference function is called (preworld.py:155)
    p2 = pred.to(DEV).requires_grad_(True)
    got2 = losses.lovasz_softmax(torch.softmax(p2, 1), target.to(DEV), ignore=17,
                                 camera_mask=md)
    assert abs(float(got2) - float(want)) <= 1e-5 * float(want)
    got2.backward()
    assert (p2.grad.cpu() - p_ref.grad).abs().max() <= 1e-4 * p_ref.grad.abs().max()
    # all four terms through loss_voxel
    lv = losses.loss_voxel(pred.to(DEV), target.to(DEV), cw.to(DEV), 17, md)
    assert abs(float(lv['loss_voxel_lovasz']) - float(want)) <= 1e-5 * float(want)


def test_lovasz_softmax_full_grid():
    """200x200x16x18 (BASELINE grid): same value as the CPU restatement."""
    from oracle import loss_ref
    from preworld_b200 import losses
    pred, target, cam, _ = loss_ref.seeded_case(5, shape=(1, 18, 200, 200, 16))
    want = loss_ref.lovasz_softmax(torch.softmax(pred, 1), target, 17, cam)
    got = losses.lovasz_softmax(pred.to(DEV), target.to(DEV), ignore=17,
                                camera_mask=cam.to(DEV), from_logits=True)
    assert abs(float(got) - float(want)) <= 2e-5 * float(want)


@pytest.mark.parametrize('seed,use_mask', [(0, True), (1, False), (2, True)])
def test_custom_focal_loss_matches_oracle(seed, use_mask, golden_dir):
    """pw_focal_loss / pw_focal_loss_grad (CustomFocalLoss, PreWorld's default CE
    term) against the CPU restatement, its autograd and the reference's value;
    built through the LOSSES registry as preworld.py:117 does."""
    import json
    import os
    from oracle import loss_ref
    from preworld_b200 import losses
    from preworld_b200.plugin import builder
    gold = json.load(open(os.path.join(golden_dir, 'voxel_losses.json')))
    pred, target, cam, cw = loss_ref.seeded_case(seed)
    cwz = torch.cat([cw, torch.zeros(1)])
    m = cam if use_mask else None
    p_ref = pred.clone().requires_grad_(True)
    want = loss_ref.custom_focal_loss(p_ref, target, cwz, 255, m)
    want.backward()
    fl = builder.build_loss(dict(type='CustomFocalLoss'))
    assert isinstance(fl, losses.CustomFocalLoss)
    p_dev = pred.to(DEV).requires_grad_(True)
    got = fl(p_dev, target.to(DEV), cwz.to(DEV), 255,
             camera_mask=None if m is None else m.to(DEV))
    assert abs(float(got) - float(want)) <= 2e-5 * float(want)
    assert abs(float(got) - gold[f'focal_seed{seed}_mask{int(use_mask)}']) <= 2e-5 * float(want)
    got.backward()
    assert (p_dev.grad.cpu() - p_ref.grad).abs().max() <= 1e-4 * p_ref.grad.abs().max()
    # loss_voxel takes the focal term by default (use_focal_loss=True, preworld.py:43)
    lv = losses.loss_voxel(pred.to(DEV), target.to(DEV), cw.to(DEV), 17,
                           None if m is None else m.to(DEV))
    assert abs(float(lv['loss_voxel_ce']) - float(want)) <= 2e-5 * float(want)


def test_loss_edge_cases():
    """Degenerate inputs of the training losses, as the reference treats them:
    no lidar return at all -> depth loss exactly 0 with zero gradient
    (view_transformer.py:788: sum / max(1, #fg)); every voxel dropped -> Lovasz 0
    with zero gradient (lovasz_softmax.py:182-184); a target with one single class
    -> the per-class terms of sem_scal reduce to that class (loss.py:56-79); an
    all-ignored target -> focal loss over zero voxels is NaN like a mean of nothing."""
    from oracle import loss_ref
    from preworld_b200 import losses
    g = torch.Generator().manual_seed(11)
    # depth loss, empty ground truth
    gt = torch.zeros(1, 2, 32, 48)
    preds = torch.softmax(torch.randn(2, 88, 2, 3, generator=g), 1).to(DEV).requires_grad_(True)
    loss = losses.get_depth_loss(gt.to(DEV), preds, 16, [1.0, 45.0, 0.5], 3.0)
    assert float(loss) == 0.0
    loss.backward()
    assert (preds.grad == 0).all()
    # Lovasz, nothing kept (every label is the ignored empty class)
    pred = torch.randn(1, 18, 4, 5, 3, generator=g)
    target = torch.full((1, 4, 5, 3), 17)
    p = pred.to(DEV).requires_grad_(True)
    lv = losses.lovasz_softmax(p, target.to(DEV), ignore=17, from_logits=True)
    assert float(lv) == 0.0
    lv.backward()
    assert (p.grad == 0).all()
    # one class only (+ some ignored voxels): device == oracle
    target = torch.full((1, 4, 5, 3), 6)
    target[0, 0, :2] = 255
    cw = torch.ones(18)
    got = losses.voxel_loss_terms(pred.to(DEV), target.to(DEV), cw.to(DEV), 255, 17)
    assert abs(float(got['sem']) - float(loss_ref.sem_scal_loss(pred, target, 255))) < 2e-5
    assert abs(float(got['ce']) - float(loss_ref.ce_ssc_loss(pred, target, cw, 255))) < 2e-5
    want_l = loss_ref.lovasz_softmax(torch.softmax(pred, 1), target, 17)
    got_l = losses.lovasz_softmax(pred.to(DEV), target.to(DEV), ignore=17, from_logits=True)
    assert abs(float(got_l) - float(want_l)) <= 1e-5 * float(want_l)
    # focal loss over zero kept voxels
    fl = losses.CustomFocalLoss()
    nan = fl(pred.to(DEV), torch.full((1, 4, 5, 3), 255).to(DEV), cw.to(DEV), 255)
    assert torch.isnan(nan)


def test_pts2ray_matches_reference_fixture(golden_dir):
    """pw_pts2ray / rays.generate_rays (datasets/ray.py:34-119) against the
    reference's own output (tests/golden/rays.npz): pixel, label, origin and
    direction columns bit-exact (the kernel rounds like the torch expression: no
    FMA), unit view directions to 1 ulp-level tolerance; sampling weights equal to
    the oracle's; weighted sampling draws the requested number of distinct rays."""
    from oracle import ray_ref
    from preworld_b200 import rays as R
    coors, depths, segs, imgs, c2ws, Ks, time_ids, dyn = ray_ref.seeded_case(0)
    dev = lambda ts: [t.to(DEV) for t in ts]
    got = R.generate_rays(dev(coors), dev(depths), dev(segs), dev(imgs), dev(c2ws), dev(Ks),
                          max_ray_nums=0, time_ids=time_ids, dynamic_class=dyn, use_wrs=False)
    gold = torch.from_numpy(np.load(os.path.join(golden_dir, 'rays.npz'))['rays'])
    assert got.shape == gold.shape
    g = got.cpu()
    assert torch.equal(g[:, :10], gold[:, :10]) and torch.equal(g[:, 13:], gold[:, 13:])
    assert (g[:, 10:13] - gold[:, 10:13]).abs().max() <= 2e-7
    per_cam = [R.pts2ray(*[t.to(DEV) for t in (coors[i], depths[i], segs[i], imgs[i], c2ws[i], Ks[i])])
               for t_ in time_ids for i in time_ids[t_]]
    ids = [t_ for t_ in time_ids for _ in time_ids[t_]]
    w_dev, _ = R.ray_weights(per_cam, ids, dyn)
    w_ref = ray_ref.ray_weights([r.cpu() for r in per_cam], ids, dyn)
    assert (torch.cat(w_dev).cpu() - torch.cat(w_ref)).abs().max() <= 1e-6
    pick = R.generate_rays(dev(coors), dev(depths), dev(segs), dev(imgs), dev(c2ws), dev(Ks),
                           max_ray_nums=1000, time_ids=time_ids, dynamic_class=dyn)
    assert pick.shape == (1000, 16)
    assert len({tuple(r) for r in pick[:, [0, 1, 4, 5, 7]].cpu().tolist()}) > 990
    assert R.pts2ray(*[t.to(DEV) for t in (coors[0][:0], depths[0][:0], segs[0][:0],
                                           imgs[0][:0], c2ws[0], Ks[0])]).shape == (0, 16)


# --------------------------------------------------- fused per-voxel heads
@pytest.mark.parametrize('m,hidden,n2,res,act2c', [
    (1000, 128, 32, True, 0),      # forecasting step shape, ragged last tile
    (128 * 149 + 5, 192, 22, False, 2),   # attribute projection shape, > 148 tiles
    (77, 32, 4, False, 0),         # one partial tile, one hidden chunk, n2p = 16
])
def test_mlp2_matches_torch(m, hidden, n2, res, act2c):
    """pw_mlp2 (fused Linear-Softplus-Linear on the tensor cores, hidden row on
    chip) vs the same two layers in torch fp32 on the CPU."""
    g = torch.Generator().manual_seed(m + hidden)
    x = torch.randn(m, 32, generator=g)
    w1 = torch.randn(hidden, 32, generator=g) / 32 ** .5
    b1 = torch.randn(hidden, generator=g)
    w2 = torch.randn(n2, hidden, generator=g) / hidden ** .5
    b2 = torch.randn(n2, generator=g)
    h = F.softplus(F.linear(x, w1, b1))
    want = F.linear(h, w2, b2)
    if act2c:
        want[:, :act2c] = F.softplus(want[:, :act2c])
    pm = ops.PackedMlp2(w1.to(DEV), b1.to(DEV), w2.to(DEV), b2.to(DEV),
                        act1='softplus', act2='softplus' if act2c else None,
                        act2_channels=act2c)
    xd = x.to(DEV)
    got = ops.mlp2(xd, pm, residual=xd if res else None)
    if res:
        want = want + x
    got = got.cpu()[:, :n2]
    assert _rel(got, want) < 5e-6, _rel(got, want)
    # per-call bias override (the ego bias of the forecasting step)
    b1b = torch.randn(hidden, generator=g)
    want2 = F.linear(F.softplus(F.linear(x, w1, b1b)), w2, b2)
    if act2c:
        want2[:, :act2c] = F.softplus(want2[:, :act2c])
    got2 = ops.mlp2(xd, pm, bias1=b1b.to(DEV)).cpu()[:, :n2]
    assert _rel(got2, want2) < 5e-6


@pytest.mark.parametrize('grid', [(40, 40, 16), (37, 9, 5), (200, 3, 16)])
def test_occhead_tail_matches_torch(grid):
    """pw_occhead_tail: 16 -> 8 (BN, ReLU) -> 18 -> argmax (+ geo), transposed to
    the [X,Y,Z] grid, vs torch."""
    gx, gy, gz = grid
    g = torch.Generator().manual_seed(gx * 7 + gz)
    feat = torch.relu(torch.randn(1, gz, gy, gx, 16, generator=g))
    w0 = torch.randn(8, 16, generator=g) / 4
    s0, b0 = torch.rand(8, generator=g) + .5, torch.randn(8, generator=g) * .1
    w1 = torch.randn(18, 8, generator=g) / 8 ** .5
    h = torch.relu(F.linear(feat, w0) * s0 + b0)
    logits = F.linear(h, w1)                                  # [1,Z,Y,X,18]
    both, lg = ops.occhead_tail(feat.to(DEV), w0.to(DEV), s0.to(DEV), b0.to(DEV),
                                w1.to(DEV), None, 17, 17, want_logits=True)
    assert _rel(lg.cpu(), logits) < 2e-6
    # argmax of the kernel's own logits, exactly; the reference argmax wherever
    # the top-2 margin is above the fp32 noise
    own = lg.cpu()[0].argmax(-1).permute(2, 1, 0)             # [X,Y,Z]
    occ = both[0].cpu().long()
    assert torch.equal(occ, own)
    top2 = logits[0].topk(2, -1).values
    clear = ((top2[..., 0] - top2[..., 1]) > 1e-4).permute(2, 1, 0)
    assert torch.equal(occ[clear], logits[0].argmax(-1).permute(2, 1, 0)[clear])
    geo = both[1].cpu().long()
    assert torch.equal(geo, torch.where(occ == 17, 17, 0))


def test_lift_bin_overflow_falls_back_bit_exact():
    """pw_lift_fused bins the kept points by voxel range (1024 voxels per bin, 4096
    entries); when a bin overflows -- here a coarse 8x8x16 grid over +-64 m: ONE bin for every
    kept point -- the call is taken by the persistent fallback kernel.  Both give
    the bits of the C oracle, and the workspace is left clean for the next call
    (a second, non-overflowing lift with the same workspace cache)."""
    geo, s2k, intr, pr, pt, bda = _lift_setup(1, seed=7)
    coarse = dict(geo.grid_config, x=[-64, 64, 16.0], y=[-64, 64, 16.0], z=[-20, 44, 4.0])
    geo2 = torch_ref.LiftGeometry(coarse, (64, 176), 16, 32)
    B, N = s2k.shape[:2]
    D, H, W = geo2.frustum.shape[:3]
    xs, ys, ds = geo2.frustum[0, 0, :, 0], geo2.frustum[0, :, 0, 1], geo2.frustum[:, 0, 0, 2]
    g = torch.Generator().manual_seed(1)
    depth = torch.rand(B * N, D, H, W, generator=g).softmax(1)
    feat = torch.randn(B * N, H, W, 32, generator=g)
    cam = ops.lift_camera_params(s2k.to(DEV), intr.to(DEV), pr.to(DEV), pt.to(DEV))
    cam_ref = cam.cpu().numpy()
    for gg in (geo2, geo):                      # overflow first, then the normal path
        grid = tuple(int(v) for v in gg.grid_size)
        rank_ref = c_ref.lift_ranks(B, N, xs.numpy(), ys.numpy(), ds.numpy(), cam_ref,
                                    bda.numpy(), gg.lower.numpy(), gg.interval.numpy(), grid)
        valid = np.where(rank_ref >= 0)[0]
        if gg is geo2:
            assert len(valid) > 4096 and grid[0] * grid[1] * grid[2] <= 1024
        order = valid[np.argsort(rank_ref[valid], kind='stable')]
        rb = rank_ref[order].astype(np.int32)
        hw = H * W
        rf = ((order // (D * hw)) * hw + order % hw).astype(np.int32)
        kept = np.ones(len(rb), bool)
        kept[1:] = rb[1:] != rb[:-1]
        starts = np.where(kept)[0].astype(np.int32)
        lengths = np.diff(np.append(starts, len(rb))).astype(np.int32)
        want = c_ref.bev_pool_v2_fwd(depth.numpy().ravel(), feat.numpy().reshape(-1, 32),
                                     order.astype(np.int32), rf, rb, starts, lengths,
                                     B * grid[0] * grid[1] * grid[2])
        got = ops.lift_fused(depth.to(DEV), feat.to(DEV), cam, bda.reshape(B, 9).to(DEV),
                             xs.to(DEV), ys.to(DEV), ds.to(DEV), gg.lower.tolist(),
                             gg.interval.tolist(), B, N, grid)
        assert np.array_equal(got.cpu().numpy().reshape(-1, 32), want)


def test_render_loss_matches_oracle():
    """NerfHead.compute_loss (nerf_head.py:271-291) on the device -- one reduction kernel
    (nine fp64 sums) -- against the CPU restatement on the masked rays; 1e-5 relative
    (fp32 log / exp per ray, fp64 accumulation)."""
    from preworld_b200.plugin.heads import NerfHead
    head = NerfHead([-40., -40., -1., 40., 40., 5.4], 0.4).to(DEV)
    g = torch.Generator().manual_seed(3)
    n = 4097
    rays = torch.zeros(n, 16)
    rays[:, 2] = torch.rand(n, generator=g) * 60                       # some > 52: masked out
    rays[:, 3] = torch.randint(0, 17, (n,), generator=g).float()
    rays[:, 13:16] = torch.randn(n, 3, generator=g)
    valid = (rays[:, 2] > 0) & (rays[:, 2] <= 52)
    res = dict(render_depth=torch.rand(n, generator=g) * 50 + 0.5,
               render_semantic=torch.randn(n, 17, generator=g) * 2,
               render_color=torch.randn(n, 3, generator=g),
               alphainv_last=torch.rand(n, generator=g), ray_mask=valid)
    want = torch_ref.nerf_compute_loss({k: v[valid] for k, v in res.items() if k != 'ray_mask'},
                                       rays[valid, 2], rays[valid, 3], rays[valid, 13:16],
                                       head.class_weights)
    got = head.compute_loss({k: v.to(DEV) for k, v in res.items()}, rays.to(DEV))
    assert sorted(got) == sorted(want)
    for k in want:
        assert abs(float(got[k]) - float(want[k])) <= 1e-5 * abs(float(want[k])), (k, got[k], want[k])
    # temporal keys (compute_loss_temporal, :301-329)
    got_t = head.compute_loss({k: v.to(DEV) for k, v in res.items()}, rays.to(DEV), interval=2)
    assert sorted(got_t) == sorted(k + '_2s' for k in want)
    # no valid ray: nothing to average (nan, as the reference's empty mean)
    none = head.compute_loss({k: v.to(DEV) for k, v in res.items() if k != 'ray_mask'} |
                             {'ray_mask': torch.zeros(n, dtype=torch.bool, device=DEV)}, rays.to(DEV))
    assert all(torch.isnan(v) for v in none.values())


def test_dense_chains_match_torch():
    """pw_dense_chains (DepthNet's Mlp + SE gate vectors, view_transformer.py:421-470,
    606-617): two 4-layer chains on 12 rows in one launch vs torch fp32, incl. the
    27 -> 28 input padding and a folded input affine (BatchNorm1d in front of fc1)."""
    g = torch.Generator().manual_seed(11)
    rows, cin, mid = 12, 27, 256
    x = torch.randn(rows, cin, generator=g)
    s_in, t_in = torch.rand(cin, generator=g) + 0.5, torch.randn(cin, generator=g) * 0.1
    acts = ['relu', None, 'relu', 'sigmoid']
    chains, refs = [], []
    for c in range(2):
        ws = [torch.randn(mid, cin if i == 0 else mid, generator=g) / (cin if i == 0 else mid) ** .5
              for i in range(4)]
        bs = [torch.randn(mid, generator=g) * 0.1 for _ in range(4)]
        layers = []
        y = x * s_in + t_in
        for i in range(4):
            kw = dict(in_scale=s_in.to(DEV), in_shift=t_in.to(DEV)) if i == 0 else {}
            layers.append((ops.PackedConv(ws[i].to(DEV), bs[i].to(DEV), None, **kw), acts[i]))
            y = y @ ws[i].t() + bs[i]
            y = torch.relu(y) if acts[i] == 'relu' else (torch.sigmoid(y) if acts[i] == 'sigmoid' else y)
        chains.append(layers)
        refs.append(y)
    xp = F.pad(x, (0, 1)).to(DEV).contiguous()
    outs = ops.dense_chains(xp, chains)
    for got, want in zip(outs, refs):
        assert got.shape == want.shape
        assert (got.cpu() - want).abs().max().item() < 2e-6
