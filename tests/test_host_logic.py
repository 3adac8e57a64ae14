"""Host-side logic of the plugin surface (runs without a GPU): registry,
config loading, state_dict contract, pose chain, frustum / geometry constants,
weight packing, and that the product refuses to run without CUDA."""
import os

import numpy as np
import pytest
import torch

from oracle import torch_ref
from oracle.cases import CASES, build_case_inputs, model_cfg_for
from preworld_b200 import (Config, ConfigDict, build_model, model_cfg, ops)
from preworld_b200 import plugin
from preworld_b200 import synthetic as S


def test_registry_surface():
    # one registry under five names (mmdet3d/models/builder.py:16-28)
    assert plugin.BACKBONES is plugin.MODELS is plugin.NECKS is plugin.HEADS \
        is plugin.DETECTORS
    for name in ('ResNet', 'CustomFPN', 'LSSViewTransformerBEVStereo',
                 'CustomResNet3D', 'LSSFPN3D', 'OccHead', 'NerfHead',
                 'BEVStereo4DOCC', 'PreWorld', 'PreWorld4DTraj',
                 'CrossEntropyLoss', 'CustomFocalLoss', 'SwinTransformer',
                 'FPN_LSS'):
        assert name in plugin.MODELS, name
    with pytest.raises(KeyError):
        plugin.build_backbone(dict(type='VoVNet'))
    with pytest.raises(TypeError):
        plugin.build_neck(dict(out_channels=3))
    m = plugin.build_neck(dict(type='LSSFPN3D', in_channels=224,
                               out_channels=32))
    assert sorted(m.state_dict())[0] == 'conv.bn.bias'


def test_reference_config_files_load_unchanged(reference_root):
    """configs/preworld/** parse with our loader and equal model_cfg()."""
    def norm(o):
        if isinstance(o, dict):
            return {k: norm(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return [norm(v) for v in o]
        return o
    files = {
        'finetune': 'nuscenes/preworld-7frame-finetune.py',
        'pretrain': 'nuscenes/preworld-7frame-pretrain.py',
        'finetune-traj': 'nuscenes-temporal/preworld-7frame-finetune-traj.py',
        'pretrain-traj': 'nuscenes-temporal/preworld-7frame-pretrain-traj.py'}
    for variant, f in files.items():
        cfg = Config.fromfile(os.path.join(reference_root, 'configs',
                                           'preworld', f))
        assert norm(cfg['model']) == norm(model_cfg(variant, 'swin')), variant
        assert cfg.model.type in plugin.DETECTORS


@pytest.mark.parametrize('variant', ['finetune', 'pretrain', 'finetune-traj'])
def test_state_dict_keys_equal_the_reference(variant, reference_root):
    """Reference checkpoints must load by key: same keys, same shapes."""
    import warnings
    warnings.filterwarnings('ignore')
    from oracle import ref_shim
    builder = ref_shim.load_all()
    cfg = model_cfg(variant, 'r50', (64, 176))
    ref = builder.build_model(ConfigDict(cfg)).state_dict()
    mine = build_model(cfg).state_dict()
    assert set(mine) == set(ref)
    assert all(mine[k].shape == ref[k].shape for k in ref)
    # and a reference state_dict loads strictly
    build_model(cfg).load_state_dict(ref, strict=True)


def test_key_addressed_init_is_construction_order_independent():
    cfg = model_cfg('finetune', 'r50', (64, 176))
    a = S.lively_init_(build_model(cfg), 3).state_dict()
    b = S.lively_init_(build_model(cfg), 3).state_dict()
    c = S.lively_init_(build_model(cfg), 4).state_dict()
    k = 'img_bev_encoder_backbone.layers.1.0.conv1.conv.weight'
    assert torch.equal(a[k], b[k]) and not torch.equal(a[k], c[k])


def test_prepare_inputs_matches_oracle():
    """bevdet_occ.py:88-139 -- fp64 pose chain, frame split."""
    case = CASES['tiny_finetune']
    model = build_model(model_cfg_for(case)).eval()
    inputs, _ = build_case_inputs(case, batch=2)
    got = model.prepare_inputs(inputs, stereo=True)
    want = torch_ref.prepare_inputs(inputs)
    assert len(got) == len(want) == 8
    for g, w in zip(got, want):
        if isinstance(w, list):
            assert len(g) == len(w)
            for gi, wi in zip(g, w):
                assert (gi is None) == (wi is None)
                if wi is not None:
                    assert torch.equal(gi, wi)
        else:
            assert torch.equal(g, w)
    assert got[7][-1] is None and len(got[0]) == 3


def test_view_transformer_constants_match_oracle():
    case = CASES['full_finetune']
    cfg = model_cfg_for(case)
    vt = plugin.build_neck(cfg['img_view_transformer'])
    geo = torch_ref.LiftGeometry(cfg['img_view_transformer']['grid_config'],
                                 (256, 704), 16, 32)
    assert vt.D == geo.D == 88
    assert torch.equal(vt.frustum, geo.frustum)
    assert torch.equal(vt.cv_frustum, geo.cv_frustum)
    assert torch.equal(vt.grid_lower_bound, geo.lower)
    assert torch.equal(vt.grid_interval, geo.interval)
    assert [int(v) for v in vt.grid_size] == [200, 200, 16]
    inputs, _ = build_case_inputs(case)
    pi = torch_ref.prepare_inputs(inputs)
    got = vt.get_mlp_input(pi[1][0], pi[2][0], pi[3][0], pi[4][0], pi[5][0],
                           pi[6])
    want = torch_ref.get_mlp_input(pi[1][0], pi[3][0], pi[4][0], pi[5][0],
                                   pi[6])
    assert got.shape == (1, 6, 27) and torch.equal(got, want)


def test_nerf_head_buffers_and_ray_parameters():
    from preworld_b200.plugin.heads import NerfHead
    h = NerfHead([-40., -40., -1., 40., 40., 5.4], 0.4, scene_center=[0, 0, 2.2])
    ng = torch_ref.NerfGeometry([-40., -40., -1., 40., 40., 5.4])
    assert torch.equal(h.scene_center, ng.scene_center)
    assert torch.equal(h.xyz_min, ng.xyz_min) and torch.equal(h.xyz_max, ng.xyz_max)
    assert abs(float(h.act_shift) - ng.act_shift) < 1e-7
    t = h.ray_parameters(torch.device('cpu'))
    assert t.numel() == 417                       # 391 inner + 26 outer
    rays = torch.zeros(4, 16)
    rays[:, 7] = 1.0
    _, _, t_ref = torch_ref.sample_ray(ng, rays[:, 4:7], rays[:, 7:10],
                                       torch.eye(3))
    assert torch.equal(t, t_ref)


def test_weight_packing():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(18, 8, 1, 1, 1, generator=g)
    pc = ops.PackedConv(w)
    assert pc.w.shape == (8, 20) and pc.cout == 18 and pc.wt_hi is None
    assert torch.equal(pc.w[:, :18], w[:, :, 0, 0, 0].t())
    assert (pc.w[:, 18:] == 0).all()
    w = torch.randn(16, 32, 3, 3, 3, generator=g)
    gamma, beta = torch.rand(16, generator=g) + .5, torch.randn(16, generator=g)
    mean, var = torch.randn(16, generator=g), torch.rand(16, generator=g) + .5
    pc = ops.PackedConv(w, None, (gamma, beta, mean, var, 1e-5), padding=1)
    assert pc.w.shape == (27 * 32, 16) and pc.pad == (1, 1, 1)
    s = gamma / torch.sqrt(var + 1e-5)
    assert torch.allclose(pc.scale, s) and torch.allclose(pc.bias, beta - mean * s)
    # tap-major then cin; tensor-core copy is the transpose, split hi + lo
    assert torch.equal(pc.w[(1 * 9 + 2 * 3 + 0) * 32 + 5], w[:, 5, 1, 2, 0])
    assert pc.wt_hi.shape == (16, 27 * 32)
    # round-to-nearest split (ops.split_tf32): both parts are tf32 values, hi is
    # the nearest one (|lo| <= half a tf32 ulp of hi), hi + lo misses w only by
    # the rounding of lo (<= 2^-22 |w|)
    wt = pc.w.t()
    assert (pc.wt_hi.view(torch.int32) & 0x1FFF == 0).all()
    assert (pc.wt_lo.view(torch.int32) & 0x1FFF == 0).all()
    assert ((pc.wt_hi + pc.wt_lo - wt).abs() <= wt.abs() * 2 ** -22).all()
    assert (pc.wt_lo.abs() <= pc.wt_hi.abs() * 2 ** -11 * 1.001 + 1e-30).all()
    # spatial permutation used by OccHead on [Z,Y,X] memory
    pr = ops.PackedConv(w, padding=1, spatial_perm=(2, 1, 0))
    assert torch.equal(pr.w[(0 * 9 + 2 * 3 + 1) * 32 + 5], w[:, 5, 1, 2, 0])
    # 2-D conv: depth stride/pad neutral; BN1d folded into a linear
    p2 = ops.PackedConv(torch.randn(8, 4, 3, 3, generator=g), stride=2,
                        padding=1, dilation=1)
    assert p2.k == (1, 3, 3) and p2.stride == (1, 2, 2) and p2.pad == (0, 1, 1)
    lw, lb = torch.randn(6, 27, generator=g), torch.randn(6, generator=g)
    sc, sh = torch.rand(27, generator=g), torch.randn(27, generator=g)
    pl = ops.PackedConv(lw, lb, in_scale=sc, in_shift=sh)
    x = torch.randn(5, 27, generator=g)
    want = torch.nn.functional.linear(x * sc + sh, lw, lb)
    got = torch.nn.functional.pad(x, (0, 1)) @ pl.w[:, :6] + pl.bias
    assert torch.allclose(got, want, atol=1e-5)


def test_fold_weight_layout_reproduces_the_conv():
    """PackedConv.set_fold_weights (the layout include/preworld_b200.h documents
    for pw_conv_fold_fwd): evaluate the folded formulation in plain torch --
    P[row, (slab,kx,n)] = sum over (kz,ky,ci) of x * Wf over INPUT columns, then
    out[x] = sum_kx P[x + kx*dw] -- and compare with conv3d.  Pure host logic:
    what the kernel's main loop and shifted-row epilogue compute."""
    g = torch.Generator().manual_seed(4)
    for cout, dil in ((32, 1), (16, 1), (40, 2)):
        cin, kd, kh, kw = 32, 3, 3, 3
        w = torch.randn(cout, cin, kd, kh, kw, generator=g)
        pc = ops.PackedConv(w, padding=dil, dilation=dil)
        fold_n = 16 if cout <= 16 else 32
        slabs = -(-cout // fold_n)
        wf = pc.wf_hi + pc.wf_lo
        assert wf.shape == (slabs * kw * fold_n, kd * kh * cin)
        assert (pc.wf_hi.view(torch.int32) & 0x1FFF == 0).all()
        x = torch.randn(1, cin, 4, 5, 9, generator=g)
        want = torch.nn.functional.conv3d(x, w, None, 1, dil, dil)
        xp = torch.nn.functional.pad(x, (dil,) * 6)[0]            # [ci, Z+2d, Y+2d, X+2d]
        Z, Y, X = x.shape[2:]
        # K index = (kz*kh + ky)*cin + ci over the (kz,ky)-shifted input columns
        cols = torch.stack([xp[:, kz * dil:kz * dil + Z, ky * dil:ky * dil + Y, :]
                            for kz in range(kd) for ky in range(kh)], 0)   # [T',ci,Z,Y,Xin]
        A = cols.permute(2, 3, 4, 0, 1).reshape(Z, Y, X + 2 * dil, kd * kh * cin)
        P = (A.double() @ wf.double().t()).view(Z, Y, X + 2 * dil, slabs, kw, fold_n)
        out = sum(P[:, :, kx * dil:kx * dil + X, :, kx] for kx in range(kw))  # [Z,Y,X,slab,n]
        got = out.reshape(Z, Y, X, slabs * fold_n)[..., :cout].permute(3, 0, 1, 2)
        assert torch.allclose(got.float(), want[0], atol=2e-4), (cout, dil)
        # rows beyond cout are zero
        if slabs * fold_n > cout:
            pad_rows = wf.view(slabs, kw, fold_n, -1)[-1, :, cout - (slabs - 1) * fold_n:]
            assert (pad_rows == 0).all()


def test_cl_layout_helpers():
    x = torch.zeros(2, 5, 7, 16)
    assert ops.cl_ld(x) == 16 and ops.cl_ld(x[..., 4:12]) == 16
    assert ops.cl_ld(x[:, :, :, None, :].squeeze(3)[:, None][:, 0]) == 16
    assert ops.cl_ld(torch.zeros(10, 24)[:, None, 2:19]) == 24
    with pytest.raises(ValueError):
        ops.cl_ld(x[:, ::2])
    lg = ops.to_logical(x)
    assert lg.shape == (2, 16, 5, 7) and ops.from_logical(lg).shape == x.shape
    with pytest.raises(ValueError):
        ops.from_logical(torch.zeros(2, 16, 5, 7))      # NCHW-contiguous


def test_no_cpu_fallback():
    """The product path fails loudly without CUDA tensors / in train mode."""
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.conv(torch.zeros(1, 4, 4, 32), ops.PackedConv(torch.zeros(8, 32)))
    m = plugin.build_neck(dict(type='LSSFPN3D', in_channels=224,
                               out_channels=32))
    m.train()
    with pytest.raises(RuntimeError, match='forward-only'):
        m.packs()
    det = build_model(model_cfg('finetune', 'r50', (64, 176)))
    with pytest.raises(NotImplementedError):
        det(return_loss=True)


def test_collapse_axes_drops_taps_that_only_see_padding():
    """PackedConv.collapse_axes: ASPP's dilation-18 3x3 conv on a 16-row map equals the 1x3
    conv of its centre kernel row (the other rows only read zero padding); a dilation that
    still reaches valid rows is left alone."""
    import torch.nn.functional as F
    from preworld_b200 import ops
    g = torch.Generator().manual_seed(0)
    w = torch.randn(12, 8, 3, 3, generator=g)
    x = torch.randn(2, 8, 16, 44, generator=g)
    pc = ops.PackedConv(w, None, None, stride=1, padding=18, dilation=18)
    col = pc.collapse_axes((1, 16, 44))
    assert col is not pc and col.k == (1, 1, 3) and col.pad == (0, 0, 18) and col.dil == (1, 1, 18)
    assert pc.collapse_axes((1, 16, 44)) is col                      # cached
    want = F.conv2d(x, w, padding=18, dilation=18)
    got = F.conv2d(x, w[:, :, 1:2, :], padding=(0, 18), dilation=(1, 18))
    assert torch.allclose(got, want, atol=1e-5)
    # packed weights of the collapsed conv are the centre row, tap-major
    assert torch.equal(col.w.view(3, 8, 12), w[:, :, 1, :].permute(2, 1, 0).contiguous())
    assert pc.collapse_axes((1, 32, 88)) is pc                       # 18 < 32: rows are reachable
    both = ops.PackedConv(w, None, None, stride=1, padding=18, dilation=18).collapse_axes((1, 16, 16))
    assert both.k == (1, 1, 1) and both.pad == (0, 0, 0)
    strided = ops.PackedConv(w, None, None, stride=2, padding=18, dilation=18)
    assert strided.collapse_axes((1, 16, 44)) is strided


def test_swin_init_weights_loads_an_official_checkpoint(tmp_path):
    """SwinTransformer(pretrained=<official-style checkpoint>).init_weights(): keys renamed,
    downsample tensors re-ordered to nn.Unfold's channel order (swin.py:25-73, 861-925), and a
    relative position table of another window size is resampled."""
    from oracle import swin_ref
    from preworld_b200.plugin.swin import swin_convert
    cfg = dict(swin_ref.TINY_SWIN, with_cp=False)
    donor = plugin.build_backbone(dict(type='SwinTransformer', **cfg))
    swin_ref.seeded_init_(donor, 21)
    # our keys -> official names (the inverse of the key part of swin_convert)
    official = {}
    for k, v in donor.state_dict().items():
        if 'relative_position_index' in k:
            continue
        k = k.replace('stages', 'layers', 1).replace('attn.w_msa.', 'attn.') \
            .replace('ffn.layers.0.0.', 'mlp.fc1.').replace('ffn.layers.1.', 'mlp.fc2.') \
            .replace('patch_embed.projection', 'patch_embed.proj')
        official[k] = v.clone()
    official['head.weight'] = torch.zeros(3, 3)
    path = str(tmp_path / 'swin_official.pth')
    torch.save({'model': official}, path)
    bb = plugin.build_backbone(dict(type='SwinTransformer', pretrained=path, **cfg))
    bb.init_weights()
    want = swin_convert(official)
    got = bb.state_dict()
    assert set(want) <= set(got) and 'head.weight' not in want
    for k, v in want.items():
        assert torch.equal(got[k], v), k
    # the reduction columns really moved (official order != nn.Unfold order)
    k = 'stages.0.downsample.reduction.weight'
    assert not torch.equal(got[k], donor.state_dict()[k])
    # a checkpoint trained with window 4 (7x7 table) -> window 6 (11x11 table)
    small = dict(cfg, window_size=4)
    donor4 = plugin.build_backbone(dict(type='SwinTransformer', **small))
    swin_ref.seeded_init_(donor4, 22)
    torch.save({'state_dict': {k: v for k, v in donor4.state_dict().items()
                               if 'relative_position_index' not in k}}, path)
    bb = plugin.build_backbone(dict(type='SwinTransformer', pretrained=path,
                                    **dict(cfg, pretrain_style='mmcls')))
    bb.init_weights()
    t = bb.state_dict()['stages.1.blocks.0.attn.w_msa.relative_position_bias_table']
    assert t.shape == (121, 2) and torch.isfinite(t).all() and t.abs().sum() > 0
