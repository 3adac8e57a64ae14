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
        self._bufs = {}

    def local_range(self, n_cams):
        return camera_split(n_cams, self.world)[self.rank]

    def _plan(self, n_cams, B, F, like):
        """Preallocated exchange buffers per (n_cams, B, F, device): the padded
        send block, the [world, B, mx, F] receive buffer and -- for an uneven
        split -- the row index that drops the padding."""
        key = (n_cams, B, F, like.device, like.dtype)
        plan = self._bufs.get(key)
        if plan is None:
            split = camera_split(n_cams, self.world)
            mx = max(c for _, c in split)
            send = like.new_zeros((B, mx, F))
            recv = like.new_empty((self.world, B, mx, F))
            even = all(c == mx for _, c in split)
            idx = None
            if not even:
                idx = torch.tensor([r * mx + k for r, (_, c) in enumerate(split)
                                    for k in range(c)], device=like.device)
            plan = self._bufs[key] = (split, mx, send, recv, even, idx)
        return plan

    def all_gather_cams(self, local, n_cams):
        """local [B, n_local, F] -> [B, n_cams, F] in camera order: ONE
        ``all_gather_into_tensor`` of equally sized blocks into a preallocated
        buffer, then one re-ordering copy (rank-major -> sample-major; for an
        uneven split the same copy drops the padding rows)."""
        B, cnt_have, F = local.shape
        if self.world == 1:
            return local
        split, mx, send, recv, even, idx = self._plan(n_cams, B, F, local)
        assert cnt_have == split[self.rank][1], (local.shape, split[self.rank])
        if cnt_have == mx and local.is_contiguous():
            src = local
        else:
            send[:, :cnt_have] = local            # padding rows stay zero
            src = send
        dist.all_gather_into_tensor(recv.view(-1), src.reshape(-1), group=self.group)
        out = recv.permute(1, 0, 2, 3).reshape(B, self.world * mx, F)
        return out if even else out.index_select(1, idx)


# ---- end of evaluation: result gather (mmdet3d/apis/test.py:165-195) -----------------
def interleave_parts(part_list, size):
    """The reference's ordering: the sampler deals samples round-robin, so rank r holds
    samples r, r + world, ...; ``zip`` the parts, flatten, drop the dataloader's padding."""
    ordered = []
    for res in zip(*part_list):
        ordered.extend(list(res))
    return ordered[:size]


def collect_occupancy(grids, size, group=None):
    """``collect_results_gpu`` for what this path produces -- a list of equally shaped
    uint8 occupancy grids per rank (device tensors or arrays) -- WITHOUT the pickle round
    trip: the grids are stacked, exchanged by ONE ``all_gather_into_tensor`` and re-ordered
    on rank 0.  Returns the ordered list of ``size`` grids on rank 0, ``None`` elsewhere
    (the reference's contract).  Every rank must hold the same number of grids (the
    reference's DistributedSampler pads to that)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    part = torch.stack([torch.as_tensor(g) for g in grids])
    if world == 1:
        return list(part[:size])
    if dist.get_backend(group) == 'nccl' and not part.is_cuda:
        part = part.cuda()                   # NCCL moves device memory only
    recv = part.new_empty((world,) + tuple(part.shape))
    dist.all_gather_into_tensor(recv.view(-1), part.contiguous().view(-1), group=group)
    if rank != 0:
        return None
    return interleave_parts([list(recv[r]) for r in range(world)], size)


def collect_results(result_part, size, group=None):
    """Generic form (arbitrary picklable per-sample results, e.g. the detectors' dicts):
    same contract and ordering as ``collect_results_gpu`` (apis/test.py:165-195)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return list(result_part)[:size]
    parts = [None] * world
    dist.all_gather_object(parts, list(result_part), group=group)
    if rank != 0:
        return None
    return interleave_parts(parts, size)


def reduce_confusion(hists, group=None):
    """Sum the ranks' confusion matrices in place (ONE all-reduce of the concatenated
    counters): with `Metric_mIoU` accumulated on each rank's device this replaces the
    gather of every prediction for the evaluation itself."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return hists
    flat = torch.cat([h.reshape(-1) for h in hists])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    o = 0
    for h in hists:
        h.copy_(flat[o:o + h.numel()].view_as(h))
        o += h.numel()
    return hists
