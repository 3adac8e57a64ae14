"""TEST INFRASTRUCTURE -- CPU restatement of the reference's ray records
(mmdet3d/datasets/ray.py:34-119: get_rays, pts2ray, the weights of generate_rays).
Only tests/ import it.  Pinned against the reference file itself (imported by path
in the build container, tests/test_oracle.py) and tests/golden/rays.npz, which
oracle/make_ray_golden.py wrote from the reference's own functions."""
import numpy as np
import torch


def get_rays(i, j, K, c2w):
    """ray.py:34-46 (inverse_y=True)."""
    dirs = torch.stack([(i - K[0][2]) / K[0][0], (j - K[1][2]) / K[1][1], torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, 3].expand(rays_d.shape)
    return rays_o, rays_d, rays_d / rays_d.norm(dim=-1, keepdim=True)


def pts2ray(coor, label_depth, label_seg, label_img, c2w, K):
    """ray.py:49-56."""
    o, d, v = get_rays(coor[:, 0] + 0.5, coor[:, 1] + 0.5, K, c2w)
    return torch.cat([coor, label_depth[:, None], label_seg[:, None], o, d, v, label_img], 1)


def ray_weights(rays, ids, dynamic_class, weight_adj=0.3, weight_dyn=0.0):
    """ray.py:88-108 with balance_weight computed from the batch."""
    classes = torch.cat([r[:, 3] for r in rays])
    class_nums = torch.Tensor([0] * 17)
    for c in range(17):
        class_nums[c] += (classes == c).sum().item()
    balance = torch.exp(0.005 * (class_nums.max() / class_nums - 1))
    out = []
    for r, tid in zip(rays, ids):
        wt = torch.full((r.shape[0],), 1.0 if tid == 0 else weight_adj)
        if tid != 0:
            wt[(dynamic_class == r[:, 3, None]).any(dim=-1)] = weight_dyn
        out.append(balance[r[..., 3].long()] * wt)
    return out


def seeded_case(seed=0, cams=6, n=257, H=256, W=704):
    """Pixel lists of `cams` cameras over 2 frames (time ids 0 and 1), every class
    present, a nuScenes-like rig."""
    g = torch.Generator().manual_seed(seed)
    coors, depths, segs, imgs, c2ws, Ks = [], [], [], [], [], []
    for c in range(2 * cams):
        coors.append(torch.stack([torch.randint(0, W, (n,), generator=g),
                                  torch.randint(0, H, (n,), generator=g)], 1).float())
        depths.append(torch.rand(n, generator=g) * 50 + 1)
        seg = torch.randint(0, 17, (n,), generator=g).float()
        seg[:17] = torch.arange(17).float()
        segs.append(seg)
        imgs.append(torch.randn(n, 3, generator=g))
        yaw = torch.tensor(2 * np.pi * (c % cams) / cams)
        R = torch.tensor([[torch.cos(yaw), -torch.sin(yaw), 0.], [torch.sin(yaw), torch.cos(yaw), 0.],
                          [0., 0., 1.]]) @ torch.tensor([[0., 0., 1.], [-1., 0., 0.], [0., -1., 0.]])
        m = torch.eye(4)
        m[:3, :3] = R + 0.01 * torch.randn(3, 3, generator=g)
        m[:3, 3] = torch.tensor([1.5 * torch.cos(yaw), 1.5 * torch.sin(yaw), 1.5])
        c2ws.append(m)
        Ks.append(torch.tensor([[1266., 0., 816.], [0., 1266., 491.], [0., 0., 1.]]) * 0.44
                  + torch.tensor([[0., 0., 0.], [0., 0., 0.], [0., 0., 0.56]]))
    time_ids = {0: list(range(cams)), 1: list(range(cams, 2 * cams))}
    dynamic_class = torch.tensor([0., 2., 3., 4., 5., 6., 7., 9., 10.])
    return coors, depths, segs, imgs, c2ws, Ks, time_ids, dynamic_class
