"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/occ_metrics.json by running the
REFERENCE's own mmdet3d/datasets/occ_metrics.py (imported by path; `tqdm`, `sklearn`
and `termcolor` -- absent here, unused by the two classes -- are stubbed) on seeded
grids: Metric_mIoU (occ_metrics.py:52-185) and Metric_mIoU_Temporal (:413-596).
Run in the build container:

    python oracle/make_metrics_golden.py [/root/reference]
"""
import contextlib
import importlib.util
import io
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load_reference(root):
    for name, attrs in (('tqdm', dict(tqdm=lambda x, *a, **k: x)),
                        ('sklearn', {}), ('sklearn.neighbors', dict(KDTree=object)),
                        ('termcolor', dict(colored=lambda s, *a, **k: s))):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                m = types.ModuleType(name)
                m.__dict__.update(attrs)
                sys.modules[name] = m
    path = os.path.join(root, 'mmdet3d', 'datasets', 'occ_metrics.py')
    spec = importlib.util.spec_from_file_location('ref_occ_metrics', path)
    mod = importlib.util.module_from_spec(spec)
    if not hasattr(np, 'int'):
        np.int = int                      # occ_metrics.py:44 (numpy < 1.24 alias; unused here)
    spec.loader.exec_module(mod)
    return mod


def seeded_case(seed, shape=(20, 24, 8)):
    """(pred, gt, mask_lidar, mask_camera) uint8/bool grids; gt has ignored (255) voxels."""
    rng = np.random.default_rng(seed)
    pred = rng.integers(0, 18, shape).astype(np.uint8)
    gt = rng.integers(0, 18, shape).astype(np.uint8)
    gt[rng.random(shape) < 0.1] = 255
    gt[rng.random(shape) < 0.4] = 17
    pred[rng.random(shape) < 0.3] = 17
    return pred, gt, rng.random(shape) < 0.6, rng.random(shape) < 0.7


def seeded_temporal_case(seed, shape=(20, 24, 8)):
    preds, gts, ml, mc = [], {}, {}, {}
    for k in range(4):                                    # predictions 0s..3s
        p, _, _, _ = seeded_case(100 * seed + k, shape)
        preds.append(p)
    for idx in (0, 2, 4, 6):
        _, g, a, b = seeded_case(100 * seed + 50 + idx, shape)
        gts[idx], ml[idx], mc[idx] = g, a, b
    return preds, gts, ml, mc


MODES = (('none', {}), ('lidar', dict(use_lidar_mask=True)), ('image', dict(use_image_mask=True)))


def reference_values(ref):
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        for mode, kw in MODES:
            m = ref.Metric_mIoU(num_classes=18, **kw)
            for seed in range(3):
                m.add_batch(*seeded_case(seed))
            _, miou, cnt, miou_res = m.count_miou()
            _, iou, _, iou_res = m.count_iou()
            out['single_' + mode] = dict(hist=m.hist.tolist(), occ_hist=m.occ_hist.tolist(),
                                         cnt=cnt, miou=miou_res, iou=iou_res,
                                         per_class=np.nan_to_num(miou, nan=-1.0).tolist())
            t = ref.Metric_mIoU_Temporal(num_classes=18, **kw)
            for seed in range(2):
                t.add_batch(*seeded_temporal_case(seed))
            _, miou_list = t.count_miou()
            out['temporal_' + mode] = dict(
                hists=[getattr(t, f'hist_{k}s').tolist() for k in range(4)],
                occ_hists=[getattr(t, f'occ_hist_{k}s').tolist() for k in range(4)],
                miou=miou_list, iou=t.count_iou())
    return out


if __name__ == '__main__':
    root = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    vals = reference_values(load_reference(root))
    path = os.path.join(ROOT, 'tests', 'golden', 'occ_metrics.json')
    json.dump(vals, open(path, 'w'))
    print('wrote', path, {k: (v['miou'], v['iou']) for k, v in vals.items()})
