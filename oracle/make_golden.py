"""TEST INFRASTRUCTURE ONLY -- golden-vector generator (build container only).

Runs the reference's own ``.py`` files *verbatim from /root/reference* on CPU
(through oracle/ref_shim.py) on seeded synthetic inputs, checks the travelling
restatement ``oracle/torch_ref.py`` against them stage by stage, and writes
small fixtures to ``tests/golden/``.  /root/reference does not exist on the GPU
box, so the fixtures (plus this script) are what is committed.

    python -m oracle.make_golden            # all cases
    python -m oracle.make_golden tiny_finetune full_finetune

Fixture contents (npz): the reference's final uint8 occupancy grids in full,
and for every float stage a strided sub-sample + (mean, mean|.|, max|.|) so a
test can localise a divergence without shipping 100 MB tensors.
"""
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

from oracle import ref_shim, torch_ref
from oracle.cases import CASES, build_case_inputs, model_cfg_for, \
    stage_sample, stage_stats, render_inputs, RENDER_CASE

warnings.filterwarnings('ignore')
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(
    os.path.abspath(__file__))), 'tests', 'golden')


def _ref_model(case):
    from preworld_b200 import synthetic as S
    if case['backbone'] == 'swin':
        # the real swin.py must be in place before detectors/bevdet.py binds its
        # isinstance() target (bevdet.py:11,589)
        from oracle import swin_shim
        swin_shim.load()
    builder = ref_shim.load_all()
    from preworld_b200.config import ConfigDict
    model = builder.build_model(ConfigDict(model_cfg_for(case)))
    model.eval()
    S.lively_init_(model, case['seed'])
    return model


def _hook_stages(model, stages):
    """Capture the reference's intermediate tensors with forward hooks."""
    calls = {'vt': 0}

    def vt_hook(mod, inp, out):
        fid = 1 - calls['vt']          # extract_img_feat visits fid=1 then 0
        stages[f'lifted_{fid}'] = out[0].detach()
        stages[f'depth_{fid}'] = out[1].detach()
        calls['vt'] += 1

    def occ_hook(mod, inp, out):        # first call == current frame
        stages.setdefault('logits', out['output_voxels'][0].detach())

    hs = [model.img_view_transformer.register_forward_hook(vt_hook),
          model.img_bev_encoder_neck.register_forward_hook(
              lambda m, i, o: stages.__setitem__('encoded', o.detach())),
          model.final_conv.register_forward_hook(
              lambda m, i, o: stages.__setitem__(
                  'voxel_feats', o.detach().permute(0, 4, 3, 2, 1))),
          model.occupancy_head.register_forward_hook(occ_hook)]
    return hs


def run_case(name):
    case = CASES[name]
    t0 = time.time()
    model = _ref_model(case)
    inputs, extra = build_case_inputs(case)
    ref_stages = {}
    hooks = _hook_stages(model, ref_stages)
    with torch.no_grad():
        ref_out = model.simple_test(None, None, img=inputs, **extra)
    for h in hooks:
        h.remove()
    t_ref = time.time() - t0

    # the travelling restatement on the same weights / inputs
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    pc = torch_ref.PathConfig(model_cfg_for(case))
    ora_stages = {}
    t0 = time.time()
    with torch.no_grad():
        if case['detector'] == 'PreWorld4DTraj':
            ora_out = torch_ref.preworld4d_simple_test(
                sd, pc, inputs, extra['temporal_ego_states'][0], ora_stages)
        else:
            ora_out = torch_ref.preworld_simple_test(sd, pc, inputs,
                                                     ora_stages)
    t_ora = time.time() - t0

    report = {'case': name, 't_reference_s': round(t_ref, 2),
              't_oracle_s': round(t_ora, 2), 'stages': {}, 'outputs': {}}
    for k, r in ref_stages.items():
        o = ora_stages[k]
        d = (o - r).abs().max().item()
        report['stages'][k] = {'max_abs_diff': d,
                               'ref_max_abs': r.abs().max().item()}
    for k, r in ref_out.items():
        mism = int((ora_out[k][0] != r[0]).sum())
        report['outputs'][k] = {'mismatch': mism, 'size': int(r[0].size)}

    fx = {}
    for k, r in ref_out.items():
        fx['out/' + k] = r[0]
    for k, r in ref_stages.items():
        fx['sample/' + k] = stage_sample(r).numpy()
        fx['stats/' + k] = np.asarray(stage_stats(r), np.float64)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + '.npz'), **fx)
    return report


def run_render():
    """NerfHead.render_one_scene + render_{depth,semantic,color} verbatim
    (nerf/nerf_head.py:165-269,332-353) on synthetic volumes."""
    ref_shim.load_all()
    nh = sys.modules['mmdet3d.models.nerf.nerf_head']
    head = nh.NerfHead(point_cloud_range=[-40., -40., -1., 40., 40., 5.4],
                       voxel_size=0.4, scene_center=[0, 0, 2.2], radius=39)
    rays, bda, density, semantic, color = render_inputs(RENDER_CASE)
    gt_depth = rays[0, :, 2].clone()
    gt_depth[gt_depth > 52] = 0
    mask = gt_depth > 0
    with torch.no_grad():
        res = head.render_one_scene(rays[0, :, 4:7], rays[0, :, 7:10], bda[0],
                                    density[0], semantic[0], color[0],
                                    mask=mask)
        depth = head.render_depth(res)
        sem = head.render_semantic(res)
        col = head.render_color(res)
    ng = torch_ref.NerfGeometry([-40., -40., -1., 40., 40., 5.4])
    with torch.no_grad():
        o = torch_ref.render_rays(ng, rays[0], bda[0], density[0], semantic[0],
                                  color[0])
    report = {'case': 'render', 'n_rays': int(mask.sum()),
              'n_samples_ref': int(res['weights'].numel()),
              'n_samples_oracle': o['n_samples'],
              'depth_diff': (o['render_depth'] - depth).abs().max().item(),
              'sem_diff': (o['render_semantic'] - sem).abs().max().item(),
              'col_diff': (o['render_color'] - col).abs().max().item(),
              'last_diff': (o['alphainv_last'] - res['alphainv_last'])
              .abs().max().item()}
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, 'render.npz'),
        render_depth=depth.numpy(), render_semantic=sem.numpy(),
        render_color=col.numpy(), alphainv_last=res['alphainv_last'].numpy(),
        ray_mask=mask.numpy(), n_samples=np.int64(res['weights'].numel()))
    return report


def main(argv):
    names = argv or (list(CASES) + ['render'])
    reports = []
    for n in names:
        r = run_render() if n == 'render' else run_case(n)
        print(json.dumps(r, indent=1))
        reports.append(r)
    path = os.path.join(GOLDEN_DIR, 'REPORT.json')
    old = {}
    if os.path.exists(path):
        old = {r['case']: r for r in json.load(open(path))}
    for r in reports:
        old[r['case']] = r
    json.dump(list(old.values()), open(path, 'w'), indent=1)


if __name__ == '__main__':
    main(sys.argv[1:])
