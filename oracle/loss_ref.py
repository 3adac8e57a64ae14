"""TEST INFRASTRUCTURE -- CPU restatement of the reference's voxel SSC losses
(mmdet3d/models/detectors/loss.py:20-113, called from preworld.py:133-155 and
preworld_temporal_traj.py:176-203).  Only tests/ may import this module; the
product path (preworld_b200/losses.py -> csrc/losses.cu) never does.

Pinned: tests/test_oracle.py runs these functions against the reference's own
loss.py (imported by path when /root/reference is present) and against the
committed values in tests/golden/voxel_losses.json (oracle/make_loss_golden.py).

The restatement is written from the sums the reference takes, not from its
control flow: every term is a masked reduction over voxels, which is also how
the CUDA kernel computes it (one pass over the logits instead of ~60)."""
import torch
import torch.nn.functional as F


def _bce_to_one(x):
    """F.binary_cross_entropy(x, ones) == -max(log x, -100)  (loss.py:61-76)."""
    return -torch.clamp(torch.log(x), min=-100.0)


def ce_ssc_loss(pred, target, class_weights, ignore_index=255):
    """loss.py:20-30.  pred [B,C,H,W,D] logits, target [B,H,W,D] integer."""
    logp = F.log_softmax(pred.float(), dim=1)
    t = target.long()
    valid = t != ignore_index
    tt = torch.where(valid, t, torch.zeros_like(t))
    w = class_weights.float()[tt] * valid
    picked = logp.gather(1, tt.unsqueeze(1)).squeeze(1)
    return -(w * picked).sum() / w.sum()


def sem_scal_loss(pred, ssc_target, ignore_index=255, camera_mask=None):
    """loss.py:33-80: per class present among the kept voxels, BCE-to-one of
    precision, recall and specificity; mean over those classes."""
    p = F.softmax(pred.float(), dim=1)
    mask = ssc_target != ignore_index
    if camera_mask is not None:
        mask = torch.logical_and(mask, camera_mask.bool())
    n_classes = p.shape[1]
    t = ssc_target[mask]
    loss, count = 0.0, 0.0
    for i in range(n_classes):
        pi = p[:, i][mask]
        hit = (t == i).float()
        cnt = hit.sum()
        if cnt > 0:
            count += 1.0
            nom = (pi * hit).sum()
            lc = 0.0
            if pi.sum() > 0:
                lc = lc + _bce_to_one(nom / pi.sum())
            lc = lc + _bce_to_one(nom / cnt)
            miss = 1.0 - hit
            if miss.sum() > 0:
                lc = lc + _bce_to_one(((1.0 - pi) * miss).sum() / miss.sum())
            loss = loss + lc
    return loss / count


def geo_scal_loss(pred, ssc_target, ignore_index=255, non_empty_idx=0,
                  camera_mask=None):
    """loss.py:83-113.  NB: the "non-empty target" is target != non_empty_idx over
    ALL voxels (ignore_index voxels count as non-empty), as in the reference."""
    p = F.softmax(pred.float(), dim=1)
    empty = p[:, non_empty_idx].reshape(-1)
    nonempty = 1.0 - empty
    mask = ssc_target != non_empty_idx
    if camera_mask is not None:
        mask = torch.logical_and(mask, camera_mask.bool())
    tgt = mask.reshape(-1).float()
    inter = (tgt * nonempty).sum()
    precision = inter / nonempty.sum()
    recall = inter / tgt.sum()
    spec = ((1.0 - tgt) * empty).sum() / (1.0 - tgt).sum()
    return _bce_to_one(precision) + _bce_to_one(recall) + _bce_to_one(spec)


def loss_voxel(pred, target, class_weights, empty_idx, camera_mask=None,
               w_ce=1.0, w_sem=1.0, w_geo=1.0):
    """The first three terms of preworld.py:151-154 (Lovasz, :155: lovasz_softmax below)."""
    cw = torch.cat([class_weights.float(), torch.zeros(1)])
    return dict(
        loss_voxel_ce=w_ce * ce_ssc_loss(pred, target, cw, 255),
        loss_voxel_sem=w_sem * sem_scal_loss(pred, target, 255, camera_mask),
        loss_voxel_geo=w_geo * geo_scal_loss(pred, target, 255, empty_idx, camera_mask))


def seeded_case(seed=0, shape=(1, 18, 20, 20, 8), ignore_frac=0.1, empty_idx=17):
    """Logits, a target that is mostly `empty_idx` with some 255s, a camera mask
    and class weights -- the same tensors on every machine."""
    g = torch.Generator().manual_seed(seed)
    b, c, h, w, d = shape
    pred = torch.randn(shape, generator=g) * 2.0
    target = torch.randint(0, c, (b, h, w, d), generator=g)
    target[torch.rand((b, h, w, d), generator=g) < 0.6] = empty_idx
    target[torch.rand((b, h, w, d), generator=g) < ignore_frac] = 255
    if seed % 2:                                   # a class that never occurs
        target[target == 3] = 4
    camera_mask = torch.rand((b, h, w, d), generator=g) < 0.7
    class_weights = torch.rand(c - 1, generator=g) + 0.5
    return pred, target, camera_mask, class_weights


# ---- depth loss (view_transformer.py:736-789, sid=False) ----------------------
def downsampled_gt_depth(gt_depths, downsample, depth_cfg, D):
    """get_downsampled_gt_depth: [B,N,H,W] lidar depths (0 = none) -> one-hot
    [B*N*h*w, D] float (all-zero rows = background)."""
    B, N, H, W = gt_depths.shape
    h, w = H // downsample, W // downsample
    g = gt_depths.view(B * N, h, downsample, w, downsample).permute(0, 1, 3, 2, 4)
    g = g.reshape(-1, downsample * downsample)
    g = torch.where(g == 0.0, torch.full_like(g, 1e5), g).min(dim=-1).values
    g = (g - (depth_cfg[0] - depth_cfg[2])) / depth_cfg[2]
    g = torch.where((g < D + 1) & (g >= 0.0), g, torch.zeros_like(g))
    return F.one_hot(g.long(), num_classes=D + 1)[:, 1:].float()


def depth_loss(depth_labels, depth_preds, downsample, depth_cfg, D, weight):
    """get_depth_loss: BCE of the D depth probabilities [B*N,D,h,w] against the
    one-hot labels, over foreground cells, / max(1, #foreground)."""
    labels = downsampled_gt_depth(depth_labels, downsample, depth_cfg, D)
    preds = depth_preds.permute(0, 2, 3, 1).reshape(-1, D)
    fg = labels.max(dim=1).values > 0.0
    loss = F.binary_cross_entropy(preds[fg], labels[fg], reduction='none')
    return weight * loss.sum() / max(1.0, float(fg.sum()))


def seeded_depth_case(seed=0, bn=(1, 6), H=64, W=176, downsample=16, D=88):
    """A sparse lidar depth map (most pixels 0, some cells entirely empty, some
    depths beyond the last bin) and softmax depth predictions."""
    g = torch.Generator().manual_seed(seed)
    B, N = bn
    gt = torch.rand((B, N, H, W), generator=g) * 60.0
    gt[torch.rand((B, N, H, W), generator=g) < 0.97] = 0.0
    gt[:, :, :downsample, :downsample * 2] = 0.0          # empty cells
    gt[:, 0, -1, -1] = 0.3                                 # below the first bin
    h, w = H // downsample, W // downsample
    preds = torch.softmax(torch.randn((B * N, D, h, w), generator=g) * 2.0, dim=1)
    return gt, preds


# ---- Lovasz-Softmax (lovasz_softmax.py:20-33,157-239) ---------------------------
def lovasz_grad(gt_sorted):
    """lovasz_softmax.py:20-33: first difference of the Jaccard index along the
    sorted order."""
    gts = gt_sorted.sum()
    inter = gts - gt_sorted.float().cumsum(0)
    union = gts + (1 - gt_sorted).float().cumsum(0)
    jac = 1.0 - inter / union
    if len(gt_sorted) > 1:
        jac = torch.cat([jac[:1], jac[1:] - jac[:-1]])
    return jac


def lovasz_softmax(probas, labels, ignore=None, camera_mask=None):
    """lovasz_softmax(..., classes='present', per_image=False): probas
    [B,C,H,W,D], labels [B,H,W,D]."""
    C = probas.shape[1]
    p = probas.permute(0, 2, 3, 4, 1).reshape(-1, C)
    t = labels.reshape(-1)
    valid = torch.ones_like(t, dtype=torch.bool) if ignore is None else t != ignore
    if camera_mask is not None:
        valid = valid & camera_mask.reshape(-1).bool()
    p, t = p[valid], t[valid]
    losses = []
    for c in range(C):
        fg = (t == c).float()
        if fg.sum() == 0:
            continue
        errors = (fg - p[:, c]).abs()
        errors_sorted, perm = torch.sort(errors, 0, descending=True)
        losses.append(torch.dot(errors_sorted, lovasz_grad(fg[perm]).detach()))
    return sum(losses) / len(losses)


# ---- CustomFocalLoss (loss_utils/focal_loss.py:11-58,162-273) --------------------
def radial_weight(H, W):
    """focal_loss.py:197-203: 1 + distance from the grid centre / its maximum."""
    xy, yx = torch.meshgrid([torch.arange(H) - H / 2, torch.arange(W) - W / 2], indexing='ij')
    c = torch.norm(torch.stack([xy, yx], 2), 2, -1)
    return c / c.max() + 1


def custom_focal_loss(pred, target, class_weights, ignore_index=255, camera_mask=None,
                      gamma=2.0, alpha=0.25, loss_weight=100.0):
    """CustomFocalLoss.forward with py_sigmoid_focal_loss (the file's own torch
    version of mmcv's CUDA op): pred [B,C,H,W,D] logits, target [B,H,W,D]."""
    B, H, W, D = target.shape
    C = pred.shape[1]
    c = radial_weight(H, W)[None, :, :, None].repeat(B, 1, 1, D).reshape(-1)
    valid = target != ignore_index
    if camera_mask is not None:
        valid = valid & camera_mask.bool()
    vis = valid.reshape(-1).nonzero().squeeze(-1)
    wm = class_weights.float()[None, :] * c[vis, None]
    x = pred.permute(0, 2, 3, 4, 1).reshape(-1, C)[vis].float()
    t = F.one_hot(target.reshape(-1)[vis].long(), num_classes=C + 1)[:, :C].float()
    p = x.sigmoid()
    pt = (1 - p) * t + p * (1 - t)
    fw = (alpha * t + (1 - alpha) * (1 - t)) * pt.pow(gamma)
    loss = F.binary_cross_entropy_with_logits(x, t, reduction='none') * fw * wm
    return loss_weight * loss.sum(-1).mean()
