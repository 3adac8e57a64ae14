"""Ray records for the rendering head on the device:
``pts2ray`` and ``generate_rays`` of mmdet3d/datasets/ray.py:34-119 with the same
signatures.  ``pts2ray`` is one kernel (``pw_pts2ray``); ``generate_rays`` keeps the
reference's weighted ray sampling (frame weight x class-balance weight,
ray.py:88-113) and draws with ``torch.multinomial`` on the rays' device --
``WeightedRandomSampler`` is the same draw on the CPU, so the sampled SET differs
from a CPU run only through the random stream.  CUDA only."""
import ctypes

import torch

from . import _lib
from .ops import _ptr, _require_cuda, _stream, check


def pts2ray(coor, label_depth, label_seg, label_img, c2w, cam_intrinsic):
    """ray.py:49-56.  coor [n,2], label_depth/label_seg [n], label_img [n,3],
    c2w [4,4], cam_intrinsic [3,3] -> [n,16] fp32."""
    f = lambda t: t.float().contiguous()
    coor, label_depth, label_seg, label_img, c2w, cam_intrinsic = map(
        f, (coor, label_depth, label_seg, label_img, c2w, cam_intrinsic))
    _require_cuda(coor, label_depth, label_seg, label_img, c2w, cam_intrinsic)
    n = coor.shape[0]
    assert coor.shape == (n, 2) and label_img.shape == (n, 3)
    assert c2w.shape == (4, 4) and cam_intrinsic.shape == (3, 3)
    rays = torch.empty((n, 16), device=coor.device, dtype=torch.float32)
    check(_lib.lib().pw_pts2ray(_ptr(coor), _ptr(label_depth), _ptr(label_seg),
                                _ptr(label_img), _ptr(c2w), _ptr(cam_intrinsic),
                                ctypes.c_longlong(n), _ptr(rays), _stream()),
          'pw_pts2ray')
    return rays


def ray_weights(rays, ids, dynamic_class, balance_weight=None, weight_adj=0.3,
                weight_dyn=0.0):
    """The sampling weight of every ray (ray.py:88-108): class balance x frame
    weight (1 for the key frame, ``weight_adj`` for adjacent frames, ``weight_dyn``
    for dynamic classes in adjacent frames).  -> (weights list, balance_weight)."""
    dev = rays[0].device
    if balance_weight is None:       # from the batch itself (ray.py:90-95)
        classes = torch.cat([r[:, 3] for r in rays])
        class_nums = torch.stack([(classes == c).sum() for c in range(17)]).float()
        balance_weight = torch.exp(0.005 * (class_nums.max() / class_nums - 1))
    balance_weight = balance_weight.to(dev)
    dynamic_class = torch.as_tensor(dynamic_class, device=dev, dtype=torch.float32)
    weights = []
    for r, tid in zip(rays, ids):
        wt = torch.full((r.shape[0],), 1.0 if tid == 0 else weight_adj, device=dev)
        if tid != 0:
            dyn = (dynamic_class == r[:, 3, None]).any(dim=-1)
            wt[dyn] = weight_dyn
        weights.append(balance_weight[r[:, 3].long()] * wt)
    return weights, balance_weight


def generate_rays(coors, label_depths, label_segs, label_imgs, c2w, intrins,
                  max_ray_nums=0, time_ids=None, dynamic_class=None,
                  balance_weight=None, weight_adj=0.3, weight_dyn=0.0, use_wrs=True,
                  generator=None):
    """ray.py:59-119."""
    rays, ids = [], []
    for time_id in time_ids:                 # frames
        for i in time_ids[time_id]:          # cameras of that frame
            rays.append(pts2ray(coors[i], label_depths[i], label_segs[i],
                                label_imgs[i], c2w[i], intrins[i]))
            ids.append(time_id)
    if not use_wrs:
        return torch.cat(rays, dim=0)
    weights, _ = ray_weights(rays, ids, dynamic_class, balance_weight, weight_adj,
                             weight_dyn)
    rays = torch.cat(rays, dim=0)
    weights = torch.cat(weights, dim=0)
    if max_ray_nums != 0 and rays.shape[0] > max_ray_nums:
        pick = torch.multinomial(weights.double(), max_ray_nums, replacement=False,
                                 generator=generator)
        rays = rays[pick]
    return rays
