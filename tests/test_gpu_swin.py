"""Parity of the Swin image side (the shipped config's backbone + neck): each kernel of
csrc/swin.cu against the plain-torch oracle (oracle/swin_ref.py, pinned bit-for-bit to the
reference's own swin.py / lss_fpn.py by tests/test_oracle.py), the whole backbone + neck
against the golden fixture the REFERENCE files produced, all through the C ABI."""
from unittest import mock

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import swin_ref                    # noqa: E402  (checker only)
from preworld_b200 import ops, plugin          # noqa: E402
from preworld_b200.plugin import swin as pswin  # noqa: E402

DEV = 'cuda'


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-6)


@pytest.mark.parametrize('rows,c,ld', [(37, 32, 32), (1000, 128, 384), (5, 2048, 2048),
                                       (64, 1024, 1024), (3, 36, 40)])
def test_layernorm_matches_torch(rows, c, ld):
    g = torch.Generator(device=DEV).manual_seed(rows + c)
    buf = torch.randn((rows, ld), device=DEV, generator=g) * 3 + 1
    x = buf[:, :c]
    gamma = torch.rand(c, device=DEV, generator=g) + 0.5
    beta = torch.randn(c, device=DEV, generator=g)
    y = ops.layernorm(x, gamma, beta, 1e-5)
    ref = F.layer_norm(x.double(), (c,), gamma.double(), beta.double(), 1e-5)
    assert (y.double() - ref).abs().max().item() <= 4e-6          # fp32 rounding of |y| <~ 10
    ops.layernorm(x, gamma, beta, 1e-5, out=x)                    # in place, channel slice
    assert torch.equal(buf[:, :c], y)


@pytest.mark.parametrize('b,h,w,c', [(2, 8, 12, 32), (1, 7, 5, 64), (1, 3, 3, 512)])
def test_patch_merge_ln_matches_oracle(b, h, w, c):
    g = torch.Generator(device=DEV).manual_seed(h * w + c)
    x = torch.randn((b, h, w, c), device=DEV, generator=g)
    down = pswin.PatchMerging(c, 2 * c, 2, dict(type='LN')).to(DEV)
    swin_ref.seeded_init_(down, 5)
    with torch.no_grad():
        ref = swin_ref.patch_merging(x.double(), down.double())
        p = down.float().pack()
        y = pswin.PatchMerging.run(p, x)
    assert y.shape == ref.shape
    # 3xTF32 reduction over K = 4C (up to 2048): accumulator noise ~2e-9 K (profiles/r02_accuracy.md)
    assert _rel(y.double(), ref) <= 2e-6 + 2.5e-9 * 4 * c


ATTN_CASES = [
    # (b, h, w, heads, ws, shift)
    (2, 12, 12, 1, 6, 0),          # exact windows
    (2, 12, 18, 2, 6, 3),          # shifted, exact windows
    (1, 16, 44, 4, 12, 6),         # Swin-B stage 3 at 256x704: padding + shift
    (1, 2, 6, 8, 6, 3),            # map smaller than a window (tiny stage 4)
    (3, 7, 9, 2, 7, 3),            # window 7 (Swin-T/S)
    (1, 13, 31, 1, 12, 0),         # padding without shift
    (1, 20, 33, 2, 16, 8),         # window 16: 256 tokens, 16 warps, four key blocks (9+9+9+5 tiles)
    (2, 9, 14, 1, 9, 4),           # window 9: 81 tokens -> 88 padded key columns, two key blocks
]


# the production kernel (mma.sync TF32, 3xTF32) and the fp32 SIMT kernels it replaced
ATTN_VARIANTS = {'mma': {'PW_ATTN_MMA': '1'}, 'mma_1cta': {'PW_ATTN_MMA': '2'},
                 'simt': {'PW_ATTN_MMA': '0', 'PW_ATTN_NQ': '1'},
                 'simt_2q': {'PW_ATTN_MMA': '0', 'PW_ATTN_NQ': '2'}}


@pytest.mark.parametrize('variant', sorted(ATTN_VARIANTS))
@pytest.mark.parametrize('b,h,w,heads,ws,shift', ATTN_CASES)
def test_window_attention_matches_oracle(b, h, w, heads, ws, shift, variant, monkeypatch):
    for k, v in ATTN_VARIANTS[variant].items():
        monkeypatch.setenv(k, v)
    c = heads * 32
    attn = pswin.ShiftWindowMSA(c, heads, ws, shift).to(DEV)
    swin_ref.seeded_init_(attn, 9)
    g = torch.Generator(device=DEV).manual_seed(h * 31 + w)
    x = torch.randn((b, h, w, c), device=DEV, generator=g)
    with torch.no_grad():
        ref = swin_ref.shifted_window_msa(x.double(), attn.double())
        attn.float()
        p = attn.w_msa.pack()
        qkv = ops.conv(x, p['qkv'])
        a = ops.window_attention(qkv, p['qkv_bias'], p['table'], heads, ws, shift,
                                 attn.w_msa.scale)
        y = ops.conv(a, p['proj'])
    assert _rel(y.double(), ref) <= 5e-6


def test_gelu_epilogue_matches_torch():
    g = torch.Generator(device=DEV).manual_seed(4)
    x = torch.randn((2, 9, 11, 64), device=DEV, generator=g)
    lin = torch.nn.Linear(64, 256).to(DEV)
    y = ops.conv(x, ops.PackedConv(lin.weight, lin.bias), 'gelu')
    with torch.no_grad():
        ref = F.gelu(F.linear(x.double(), lin.weight.double(), lin.bias.double()))
    assert _rel(y.double(), ref) <= 3e-6


def _tiny_modules():
    bb = plugin.build_backbone(dict(type='SwinTransformer', with_cp=False,
                                    **swin_ref.TINY_SWIN)).eval()
    neck = plugin.build_neck(dict(type='FPN_LSS', **swin_ref.TINY_NECK)).eval()
    swin_ref.seeded_init_(bb, 11)
    swin_ref.seeded_init_(neck, 12)
    return bb.to(DEV), neck.to(DEV)


def test_backbone_and_neck_match_the_reference_fixture(golden_dir):
    """SwinTransformer.forward + FPN_LSS.forward (swin.py:927-970, lss_fpn.py:83-99)
    against outputs of the reference's own files on the same seeded weights / input.
    Tolerance: 2e-5 of the tensor's maximum (3xTF32 linears, fp32 elsewhere; 8 blocks)."""
    gold = np.load(f'{golden_dir}/tiny_swin.npz')
    bb, neck = _tiny_modules()
    x = swin_ref.tiny_input(DEV)
    with torch.no_grad():
        outs = bb(x)
        n = neck(outs[1:])
    for name, o in zip(('stereo', 'out2', 'out3'), outs):
        ref = torch.from_numpy(gold[name]).to(DEV)
        assert o.shape == ref.shape, name
        assert _rel(o, ref) <= 2e-5, (name, _rel(o, ref))
    ref = torch.from_numpy(gold['neck']).to(DEV)
    assert n.shape == ref.shape and _rel(n, ref) <= 2e-5, _rel(n, ref)


def test_stage_entry_points_equal_forward():
    """run_stem / run_layer(0) / run_from_layer(1) -- the surface the detectors batch the
    frames over (stereo reference frame: stage 0 only, bevdet.py:589-604) -- give forward's
    bits."""
    bb, _ = _tiny_modules()
    x = swin_ref.tiny_input(DEV)
    with torch.no_grad():
        outs = bb(x)
        l0 = bb.run_layer(0, bb.run_stem(x))
        rest = bb.run_from_layer(1, l0)
    assert bb.stage0_is_stereo
    assert torch.equal(ops.to_logical(l0), outs[0])
    assert len(rest) == 2 and all(torch.equal(a, b) for a, b in zip(rest, outs[1:]))


def test_shipped_config_routes_agree():
    """The shipped model dict (Swin image side, reduced depth / window) through the
    frame-batched route, the CUDA-graph route and the frame-by-frame route of the reference
    (bevdet_occ.py:219-240): identical occupancy grids.  Parity of this model with the
    verbatim reference: tests/test_gpu_model.py, case tiny_swin_finetune."""
    from oracle.cases import CASES, build_case_inputs, model_cfg_for
    from preworld_b200 import build_model, synthetic as S
    case = CASES['tiny_swin_finetune']
    model = build_model(model_cfg_for(case)).eval()
    S.lively_init_(model, case['seed'])
    model = model.cuda()
    inputs, _ = build_case_inputs(case)
    dev = tuple(t.cuda() for t in inputs)
    with torch.no_grad():
        a = model.simple_test(None, None, img=dev)
        with mock.patch.object(type(model.img_backbone), 'stage0_is_stereo',
                               new_callable=mock.PropertyMock, return_value=False):
            b = model.simple_test(None, None, img=dev)      # frame-by-frame route
        model.enable_cuda_graph()
        host = tuple(t.cpu().pin_memory() for t in inputs)
        c = model(return_loss=False, img_inputs=[host], img_metas=[None])
        c2 = model(return_loss=False, img_inputs=[host], img_metas=[None])
    for k in ('semantic_occ', 'geo_occ'):
        assert (a[k][0] == b[k][0]).all(), k
        assert (a[k][0] == c[k][0]).all() and (a[k][0] == c2[k][0]).all(), k



def test_fpn_lss_optional_branches_match_oracle():
    """lateral + input_conv + extra upsampling (up2) of FPN_LSS against the restatement
    (pinned to the reference class by tests/test_oracle.py)."""
    cfg = dict(in_channels=128 + 256, out_channels=32, scale_factor=2, input_feature_index=(0, 1),
               extra_upsample=2, lateral=128, use_input_conv=True)
    neck = plugin.build_neck(dict(type='FPN_LSS', **cfg)).eval()
    swin_ref.seeded_init_(neck, 4)
    neck = neck.to(DEV)
    g = torch.Generator(device=DEV).manual_seed(2)
    feats = [torch.randn((2, 6, 10, 128), device=DEV, generator=g).permute(0, 3, 1, 2),
             torch.randn((2, 3, 5, 256), device=DEV, generator=g).permute(0, 3, 1, 2)]
    with torch.no_grad():
        got = neck(feats)
        want = swin_ref.neck_forward(neck.double(), [f.double() for f in feats])
    assert got.shape == want.shape == (2, 32, 12, 20)
    assert _rel(got.double(), want) <= 1e-5
