"""Writes tests/golden/rays.npz from the REFERENCE's mmdet3d/datasets/ray.py
(pure torch; imported by path) on oracle.ray_ref.seeded_case.

    python oracle/make_ray_golden.py [/root/reference]
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ray_ref  # noqa: E402


def load_reference(root):
    path = os.path.join(root, 'mmdet3d', 'datasets', 'ray.py')
    spec = importlib.util.spec_from_file_location('ref_ray', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference(sys.argv[1] if len(sys.argv) > 1 else '/root/reference')
    coors, depths, segs, imgs, c2ws, Ks, time_ids, dyn = ray_ref.seeded_case(0)
    rays = ref.generate_rays(coors, depths, segs, imgs, c2ws, Ks, max_ray_nums=0,
                             time_ids=time_ids, dynamic_class=dyn, use_wrs=False)
    path = os.path.join(ROOT, 'tests', 'golden', 'rays.npz')
    np.savez_compressed(path, rays=rays.numpy())
    print('wrote', path, tuple(rays.shape))


if __name__ == '__main__':
    main()
