"""TEST INFRASTRUCTURE ONLY -- never imported by preworld_b200 or by bench.py's timed path.

CPU oracle of the pixel side of ``PrepareImageInputs`` (reference
mmdet3d/datasets/pipelines/loading.py).  The reference's arithmetic here IS third-party code:
four PIL calls and mmcv's ``imnormalize`` (cv2).  PIL 12 and cv2 4.13 are present in the image
(build container and GPU box), so the oracle calls them exactly as the reference does:

  img_transform_core   loading.py:954-961   resize -> crop -> transpose -> rotate (verbatim calls)
  mmlab_normalize      loading.py:847-854 + mmcv 1.6.0 image/photometric.py:imnormalize_
                       (mmcv is absent: its five cv2 lines are restated -- copy to fp32,
                       BGR2RGB in place, cv2.subtract(mean), cv2.multiply(1/std))
  reference_class      the reference's own PrepareImageInputs ClassDef executed from its file
                       (build container only), used by tests to pin the two functions above and
                       preworld_b200/pixels.py:View

Pinned: tests/test_oracle.py::test_pixel_oracle_matches_reference_class.
"""
import ast
import os

import numpy as np
import torch


def img_transform_core(img, resize_dims, crop, flip, rotate):
    from PIL import Image
    img = img.resize(resize_dims)
    img = img.crop(crop)
    if flip:
        img = img.transpose(method=Image.FLIP_LEFT_RIGHT)
    return img.rotate(rotate)


def imnormalize(img, mean, std, to_rgb=True):
    import cv2
    img = img.copy().astype(np.float32)
    mean = np.float64(mean.reshape(1, -1))
    stdinv = 1 / np.float64(std.reshape(1, -1))
    if to_rgb:
        cv2.cvtColor(img, cv2.COLOR_BGR2RGB, img)
    cv2.subtract(img, mean, img)
    cv2.multiply(img, stdinv, img)
    return img


def mmlab_normalize(img):
    mean = np.array([123.675, 116.28, 103.53], dtype=np.float32)
    std = np.array([58.395, 57.12, 57.375], dtype=np.float32)
    out = imnormalize(np.array(img), mean, std, True)
    return torch.tensor(out).float().permute(2, 0, 1).contiguous()


def network_input(img_u8_hwc, view):
    """uint8 HWC array + preworld_b200.pixels.View -> [3,fH,fW] fp32, through PIL / cv2."""
    from PIL import Image
    img = Image.fromarray(img_u8_hwc, 'RGB')
    out = img_transform_core(img, view.resized_wh, view.crop, view.flip, view.rotate)
    return mmlab_normalize(out)


def synthetic_photo(h, w, seed):
    """A uint8 RGB test image with smooth gradients, sharp edges and noise (resampling is
    sensitive to all three)."""
    g = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([127 + 120 * np.sin(xx / 37.0 + c) * np.cos(yy / 23.0 - c) for c in range(3)], -1)
    img += 60 * (((xx // 16 + yy // 16) % 2) - 0.5)[..., None]
    img += g.randn(h, w, 3) * 25
    return np.clip(img, 0, 255).astype(np.uint8)


def reference_class(reference_root):
    """PrepareImageInputs of the reference, compiled from its own ClassDef (the module imports
    mmcv / pyquaternion / nuscenes at the top), with PIL.Image and the normaliser above."""
    from PIL import Image
    path = os.path.join(reference_root, 'mmdet3d', 'datasets', 'pipelines', 'loading.py')
    tree = ast.parse(open(path).read())
    node = next(n for n in tree.body
                if isinstance(n, ast.ClassDef) and n.name == 'PrepareImageInputs')
    node.decorator_list = []
    ns = dict(torch=torch, np=np, mmlabNormalize=mmlab_normalize, Image=Image, os=os)
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, 'exec'), ns)
    return ns['PrepareImageInputs']
