"""Within-sample camera sharding (BASELINE.json north_star).

The image side of one sample -- backbone, FPN, DepthNet and the stereo cost
volume -- is independent per camera (the cost volume needs only the SAME
camera's previous-frame stereo feature, reference view_transformer.py:577-598).
Each rank therefore runs it for a contiguous block of cameras, then ONE
all-gather exchanges every frame's per-camera depth distribution
``[D,h,w]`` and context feature ``[h,w,C]`` (0.33 MB per camera and frame at
256x704), after which every rank runs the same deterministic lift and 3-D
stages and holds bit-identical voxel features.  (A sum over cameras, which a
sharded lift would need, would change the fp32 summation order.)

``torch.distributed`` is used for the plumbing only (NCCL over NVLink on the
GPUs, gloo in the CPU tests); without an initialised process group the shard
degenerates to the single-rank identity.
"""
import torch
import torch.distributed as dist


def camera_split(n_cams, world):
    """Contiguous block (start, count) of cameras for every rank; the first
    ``n_cams % world`` ranks take one more, ranks beyond ``n_cams`` idle."""
    base, extra = divmod(n_cams, world)
    out, start = [], 0
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        out.append((start, cnt))
        start += cnt
    return out


class CameraShard:
    def __init__(self, rank=None, world=None, group=None):
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        assert 0 <= rank < world
        self.rank, self.world, self.group = rank, world, group

    def local_range(self, n_cams):
        return camera_split(n_cams, self.world)[self.rank]

    def all_gather_cams(self, local, n_cams):
        """local [B, n_local, F] -> [B, n_cams, F] in camera order (one
        all-gather of equally padded blocks)."""
        split = camera_split(n_cams, self.world)
        start, cnt = split[self.rank]
        assert local.shape[1] == cnt, (local.shape, cnt)
        if self.world == 1:
            return local
        mx = max(c for _, c in split)
        B, _, F = local.shape
        pad = local.new_zeros((B, mx, F))
        pad[:, :cnt] = local
        outs = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(outs, pad, group=self.group)
        return torch.cat([o[:, :c] for o, (_, c) in zip(outs, split)], dim=1)
