"""Generates tests/golden/voxel_losses.json by running the REFERENCE's own
mmdet3d/models/detectors/loss.py (imported by path; it needs only torch) on the
seeded cases of oracle/loss_ref.py.  Run in the build container:

    python oracle/make_loss_golden.py [/root/reference]
"""
import importlib.util
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import loss_ref  # noqa: E402


def load_reference_file(root, name):
    path = os.path.join(root, 'mmdet3d', 'models', 'detectors', name)
    spec = importlib.util.spec_from_file_location('ref_' + name[:-3], path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference(root):
    return load_reference_file(root, 'loss.py')


def reference_values(ref, seed, use_mask):
    pred, target, cam, cw = loss_ref.seeded_case(seed)
    cwz = torch.cat([cw, torch.zeros(1)])
    m = cam if use_mask else None
    return dict(
        ce=float(ref.CE_ssc_loss(pred, target, cwz, 255)),
        sem=float(ref.sem_scal_loss(pred, target, 255, camera_mask=m)),
        geo=float(ref.geo_scal_loss(pred, target, 255, non_empty_idx=17, camera_mask=m)))


DEPTH_CFG = [1.0, 45.0, 0.5]                 # bevstereo-occ.py grid_config['depth']


def reference_depth_loss(seed):
    """LSSViewTransformerBEVDepth.get_depth_loss of the reference file (loaded
    under oracle/ref_shim.py) on oracle.loss_ref.seeded_depth_case(seed)."""
    import types
    from oracle import ref_shim
    ref_shim.install()
    cls = ref_shim.load('mmdet3d.models.necks.view_transformer').LSSViewTransformerBEVDepth
    gt, preds = loss_ref.seeded_depth_case(seed)
    me = types.SimpleNamespace(downsample=16, sid=False, D=88, loss_depth_weight=3.0,
                               grid_config={'depth': DEPTH_CFG})
    me.get_downsampled_gt_depth = lambda g: cls.get_downsampled_gt_depth(me, g)
    labels = cls.get_downsampled_gt_depth(me, gt)
    return dict(loss=float(cls.get_depth_loss(me, gt, preds)),
                n_fg=int((labels.sum(1) > 0).sum()),
                label_checksum=int((labels.argmax(1) * (labels.sum(1) > 0)).sum()))


def reference_focal_loss(root, seed, use_mask):
    """CustomFocalLoss.forward of the reference (loss_utils/focal_loss.py:221-270) with
    its CPU branch py_sigmoid_focal_loss; the module is imported by path with mmcv.ops /
    mmdet.models.losses.utils stubbed (neither is called on this branch).  `self.c` is
    rebuilt exactly as in __init__ (:197-203, minus the .cuda())."""
    import types
    from oracle import ref_shim
    ref_shim.install()
    ref_shim._install('mmcv.ops', sigmoid_focal_loss=None)
    ref_shim._install('mmdet.models.losses')
    ref_shim._install('mmdet.models.losses.utils', weight_reduce_loss=None)
    import sys as _sys
    if not hasattr(_sys.modules['mmdet.models.builder'], 'LOSSES'):
        _sys.modules['mmdet.models.builder'].LOSSES = _sys.modules['mmdet.models.builder'].HEADS
    path = os.path.join(root, 'mmdet3d', 'models', 'loss_utils', 'focal_loss.py')
    spec = importlib.util.spec_from_file_location('ref_focal', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    pred, target, cam, cw = loss_ref.seeded_case(seed)
    cwz = torch.cat([cw, torch.zeros(1)])
    B, H, W, D = target.shape
    xy, yx = torch.meshgrid([torch.arange(H) - H / 2, torch.arange(W) - W / 2])
    c = torch.norm(torch.stack([xy, yx], 2), 2, -1)
    me = types.SimpleNamespace(c=c / c.max() + 1, use_sigmoid=True, activated=False, gamma=2.0,
                               alpha=0.25, reduction='mean', loss_weight=100.0)
    return float(mod.CustomFocalLoss.forward(me, pred, target, cwz, None, 255, None,
                                             cam if use_mask else None))


def main():
    root = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    ref = load_reference(root)
    out = {}
    for seed in range(4):
        for use_mask in (False, True):
            out[f'seed{seed}_mask{int(use_mask)}'] = reference_values(ref, seed, use_mask)
    for seed in range(3):
        out[f'depth_seed{seed}'] = reference_depth_loss(seed)
    lov = load_reference_file(root, 'lovasz_softmax.py')
    for seed in range(4):
        for use_mask in (False, True):
            pred, target, cam, _ = loss_ref.seeded_case(seed)
            out[f'lovasz_seed{seed}_mask{int(use_mask)}'] = float(lov.lovasz_softmax(
                torch.softmax(pred, dim=1), target, ignore=17,
                camera_mask=cam if use_mask else None))
    for seed in range(3):
        for use_mask in (False, True):
            out[f'focal_seed{seed}_mask{int(use_mask)}'] = reference_focal_loss(root, seed, use_mask)
    path = os.path.join(ROOT, 'tests', 'golden', 'voxel_losses.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print('wrote', path)
    for k, v in out.items():
        print(k, v)


if __name__ == '__main__':
    main()
