"""TEST INFRASTRUCTURE ONLY -- the seeded parity cases shared by
oracle/make_golden.py (build container) and tests/ (container + GPU box).

Each case names a model dict (preworld_b200.configs.model_cfg -- equal to the
reference's config files, see tests/test_configs.py), an input size, a weight
seed and an input seed.  'tiny_*' cases run through the CPU oracle in about a
second; 'full_finetune' is BASELINE.json configs[0]/[1] (6x3x256x704 ->
200x200x16).
"""
import numpy as np
import torch

from preworld_b200 import configs as C
from preworld_b200 import synthetic as S

TINY_GRID = C.grid_config(x=(-8, 8, 0.4), y=(-8, 8, 0.4))

CASES = {
    'tiny_finetune': dict(variant='finetune', backbone='r50',
                          input_size=(64, 176), grid=TINY_GRID, seed=0,
                          input_seed=0, detector='PreWorld'),
    'tiny_pretrain': dict(variant='pretrain', backbone='r50',
                          input_size=(64, 176), grid=TINY_GRID, seed=3,
                          input_seed=4, detector='PreWorld'),
    'tiny_traj': dict(variant='finetune-traj', backbone='r50',
                      input_size=(64, 176), grid=TINY_GRID, seed=5,
                      input_seed=6, detector='PreWorld4DTraj'),
    'tiny_pretrain_traj': dict(variant='pretrain-traj', backbone='r50',
                               input_size=(64, 176), grid=TINY_GRID, seed=7,
                               input_seed=8, detector='PreWorld4DTraj'),
    # the SHIPPED image side (SwinTransformer + FPN_LSS, bevstereo-occ.py:45-74) at reduced
    # depth / window; width 128 as Swin-B (the stereo feature must be 128 channels wide)
    'tiny_swin_finetune': dict(variant='finetune', backbone='swin',
                               input_size=(64, 192), grid=TINY_GRID, seed=9,
                               input_seed=10, detector='PreWorld'),
    'full_finetune': dict(variant='finetune', backbone='r50',
                          input_size=(256, 704), grid=None, seed=0,
                          input_seed=0, detector='PreWorld'),
}

RENDER_CASE = dict(num_rays=768, seed=11)


def model_cfg_for(case, reference_root=None):
    cfg = C.model_cfg(case['variant'], case['backbone'], case['input_size'],
                      case['grid'])
    if case['backbone'] == 'swin':
        cfg['img_backbone'].update(depths=[2, 2, 2, 2], window_size=6, with_cp=False)
        cfg['img_neck'].update(out_channels=128)
        cfg['img_view_transformer'].update(in_channels=128,
                                           input_size=tuple(case['input_size']))
    return cfg


def build_case_inputs(case, batch=1):
    inputs = S.make_img_inputs(batch, case['input_size'],
                               seed=case['input_seed'])
    extra = {}
    if case['detector'] == 'PreWorld4DTraj':
        ego = S.make_ego_states(batch, seed=case['input_seed'] + 100)
        # kwargs['temporal_ego_states'][0][0] is the [B,1,21] tensor the
        # reference reads (preworld_temporal_traj.py:229,331)
        extra['temporal_ego_states'] = [[ego]]
    return inputs, extra


def stage_sample(t, n=4096):
    """Deterministic strided sub-sample of a stage tensor."""
    flat = t.detach().reshape(-1).float().cpu()
    step = max(1, flat.numel() // n)
    return flat[::step][:n].clone()


def stage_stats(t):
    t = t.detach().float()
    return (t.mean().item(), t.abs().mean().item(), t.abs().max().item())


def render_inputs(rc=RENDER_CASE):
    """Synthetic density/semantic/colour volumes [1,200,200,16(,C)] (the
    NerfHead hard-codes a 200x200x16 world, nerf_head.py:150) and rays
    [1,R,16].  Densities span ~[0, 25] so that both the alpha>1e-7 filter and
    the T<1e-3 early stop of alpha2weight trigger."""
    g = torch.Generator().manual_seed(rc['seed'])
    img_inputs = S.make_img_inputs(1, (256, 704), seed=rc['seed'])
    rays = S.make_rays(img_inputs, rc['num_rays'], seed=rc['seed'] + 1)
    rays[0, ::7, 2] = 60.0                 # some rays beyond the 52 m cut
    X, Y, Z = 200, 200, 16
    low = torch.randn(1, 1, 25, 25, 4, generator=g)
    field = torch.nn.functional.interpolate(
        low, size=(X, Y, Z), mode='trilinear', align_corners=True)[0, 0]
    density = torch.nn.functional.softplus(
        field * 9.0 + 3.0 + torch.randn(X, Y, Z, generator=g))[None]
    semantic = torch.randn(1, X, Y, Z, 17, generator=g)
    color = torch.randn(1, X, Y, Z, 3, generator=g)
    bda = img_inputs[6]
    return rays, bda, density, semantic, color
