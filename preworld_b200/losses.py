"""Voxel SSC training losses on the device: the CE,
semantic-scal and geometric-scal terms of the reference's ``loss_voxel``
(mmdet3d/models/detectors/preworld.py:129-157, functions in
mmdet3d/models/detectors/loss.py:20-113), computed by ONE pass of
``csrc/losses.cu`` over the logits, with closed-form gradients from the same
statistics (one more pass per term that is back-propagated).

Same names, argument meaning and results as the reference functions; inputs are
the reference's logical tensors (``pred`` [B,C,H,W,D] logits, ``target``
[B,H,W,D] integer labels with 255 = ignore, ``camera_mask`` [B,H,W,D] bool).
The Lovasz term (preworld.py:155) is ``lovasz_softmax`` below.  CUDA only: there is no CPU
fallback (the CPU restatement lives in ``oracle/loss_ref.py`` for the tests)."""
import torch

from . import ops

_TERMS = ('ce', 'sem', 'geo')


def _rows(pred):
    """[B,C,*spatial] logits -> [V, C] rows (a view when pred is channels-last,
    as the heads of this package produce it; one transposing copy otherwise)."""
    d = pred.dim()
    v = pred.permute(0, *range(2, d), 1)
    if not v.is_contiguous():
        v = v.contiguous()
    return v.reshape(-1, pred.shape[1])


def _labels(target):
    t = target.reshape(-1)
    return (t if t.dtype == torch.uint8 else t.to(torch.uint8)).contiguous()


class _Term(torch.autograd.Function):
    """One of the three losses as a differentiable scalar; the statistics were
    computed once for all three."""

    @staticmethod
    def forward(ctx, rows, k, target, camera_mask, class_weights, empty_idx,
                ignore_index, stats, losses):
        ctx.k, ctx.empty_idx, ctx.ignore_index = k, empty_idx, ignore_index
        ctx.save_for_backward(rows, target, class_weights, stats)
        ctx.camera_mask = camera_mask
        return losses[k].clone()

    @staticmethod
    def backward(ctx, g):
        rows, target, class_weights, stats = ctx.saved_tensors
        w = [0.0, 0.0, 0.0]
        w[ctx.k] = 1.0
        grad = ops.voxel_loss_grad(rows, target, ctx.camera_mask, class_weights,
                                   ctx.empty_idx, stats, *w,
                                   ignore_index=ctx.ignore_index)
        return (grad * g,) + (None,) * 8


def voxel_loss_terms(pred, target, class_weights, ignore_index=255,
                     non_empty_idx=0, camera_mask=None):
    """-> dict(ce=, sem=, geo=) of differentiable scalars: CE_ssc_loss(pred,
    target, class_weights, ignore_index), sem_scal_loss(pred, target,
    ignore_index, camera_mask) and geo_scal_loss(pred, target, ignore_index,
    non_empty_idx, camera_mask) of loss.py:20-113 from one pass."""
    if not pred.is_cuda:
        raise RuntimeError('preworld_b200.losses needs CUDA tensors '
                           '(there is no CPU fallback)')
    rows = _rows(pred.float())
    t = _labels(target)
    cam = None if camera_mask is None else \
        camera_mask.reshape(-1).to(torch.uint8).contiguous()
    cw = class_weights.to(pred.device).float().contiguous()
    stats, losses = ops.voxel_loss_stats(rows.detach(), t, cam, cw, non_empty_idx,
                                         ignore_index)
    return {name: _Term.apply(rows, k, t, cam, cw, non_empty_idx, ignore_index,
                              stats, losses)
            for k, name in enumerate(_TERMS)}


def CE_ssc_loss(pred, target, class_weights, ignore_index):
    """loss.py:20-30."""
    return voxel_loss_terms(pred, target, class_weights, ignore_index)['ce']


def sem_scal_loss(pred, ssc_target, ignore_index, camera_mask=None):
    """loss.py:33-80."""
    cw = torch.ones(pred.shape[1], device=pred.device)
    return voxel_loss_terms(pred, ssc_target, cw, ignore_index,
                            camera_mask=camera_mask)['sem']


def geo_scal_loss(pred, ssc_target, ignore_index, non_empty_idx=0,
                  camera_mask=None):
    """loss.py:83-113."""
    cw = torch.ones(pred.shape[1], device=pred.device)
    return voxel_loss_terms(pred, ssc_target, cw, ignore_index, non_empty_idx,
                            camera_mask)['geo']


def loss_voxel(output_voxels, target_voxels, class_weights, empty_idx,
               camera_mask=None, weight_voxel_ce=1.0, weight_voxel_sem_scal=1.0,
               weight_voxel_geo_scal=1.0, weight_voxel_lovasz=1.0,
               use_focal_loss=True, focal_loss=None):
    """PreWorld.loss_voxel (preworld.py:129-157), all four terms
    (``weight_voxel_lovasz=None`` leaves the Lovasz term out; the CE term is the
    focal loss unless ``use_focal_loss=False``, as in the reference):
    ``class_weights`` are the per-class weights WITHOUT the empty class (a zero
    is appended, preworld.py:150)."""
    # preworld.py:130-131: non-finite logits are zeroed before any term is formed (one
    # NaN would otherwise poison all four losses and their gradients); the gradient
    # of a sanitised logit is zero, as with the reference's in-place assignment
    output_voxels = torch.nan_to_num(output_voxels, nan=0.0, posinf=0.0, neginf=0.0)
    cw = torch.cat([class_weights.to(output_voxels.device).float(),
                    torch.zeros(1, device=output_voxels.device)])
    t = voxel_loss_terms(output_voxels, target_voxels, cw, 255, empty_idx,
                         camera_mask)
    ce = t['ce']
    if use_focal_loss:                                  # preworld.py:146-148 (the default)
        focal_loss = focal_loss or CustomFocalLoss()
        ce = focal_loss(output_voxels, target_voxels, cw, 255, camera_mask=camera_mask)
    out = dict(loss_voxel_ce=weight_voxel_ce * ce,
               loss_voxel_sem=weight_voxel_sem_scal * t['sem'],
               loss_voxel_geo=weight_voxel_geo_scal * t['geo'])
    if weight_voxel_lovasz is not None:
        # preworld.py:155: lovasz_softmax(softmax(out), target, ignore=empty_idx, mask)
        out['loss_voxel_lovasz'] = weight_voxel_lovasz * lovasz_softmax(
            output_voxels, target_voxels, ignore=empty_idx, camera_mask=camera_mask,
            from_logits=True)
    return out


class _DepthLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth_preds, gt_depth, downsample, depth_min, depth_step, weight):
        loss, labels, sums = ops.depth_loss(gt_depth, depth_preds, downsample,
                                            depth_min, depth_step, weight)
        ctx.save_for_backward(depth_preds, labels, sums)
        ctx.weight = weight
        return loss[0].clone()

    @staticmethod
    def backward(ctx, g):
        depth_preds, labels, sums = ctx.saved_tensors
        return (ops.depth_loss_grad(labels, depth_preds, sums, ctx.weight) * g,
                None, None, None, None, None)


def get_depth_loss(depth_labels, depth_preds, downsample, depth_cfg, loss_depth_weight):
    """LSSViewTransformerBEVDepth.get_depth_loss (view_transformer.py:736-789,
    sid=False): depth_labels [B,N,H,W] lidar depth maps, depth_preds [B*N,D,h,w]
    depth probabilities, depth_cfg = grid_config['depth'] = (min, max, step)."""
    if not depth_preds.is_cuda:
        raise RuntimeError('preworld_b200.losses needs CUDA tensors '
                           '(there is no CPU fallback)')
    B, N, H, W = depth_labels.shape
    return _DepthLoss.apply(depth_preds.float(), depth_labels.reshape(B * N, H, W),
                            int(downsample), float(depth_cfg[0]), float(depth_cfg[2]),
                            float(loss_depth_weight))


class _Lovasz(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rows, is_logits, target, camera_mask, ignore_label):
        loss, gp = ops.lovasz_softmax_rows(rows, is_logits, target, camera_mask,
                                           ignore_label, want_grad=True)
        ctx.is_logits = is_logits
        ctx.save_for_backward(rows, gp)
        return loss[0].clone()

    @staticmethod
    def backward(ctx, g):
        rows, gp = ctx.saved_tensors
        grad = ops.softmax_backward(rows, gp) if ctx.is_logits else gp
        return grad * g, None, None, None, None


def lovasz_softmax(probas, labels, classes='present', per_image=False,
                   ignore=None, camera_mask=None, from_logits=False):
    """lovasz_softmax.py:157-175 for the configuration the path uses
    (classes='present', per_image=False): ``probas`` [B,C,H,W,D] class
    probabilities, ``labels`` [B,H,W,D]; voxels with label == ``ignore`` (and
    outside ``camera_mask``) are dropped.  ``from_logits=True`` takes logits and
    fuses the softmax (what ``loss_voxel`` uses)."""
    if classes != 'present' or per_image:
        raise NotImplementedError('lovasz_softmax: only classes="present", '
                                  'per_image=False (preworld.py:155)')
    if not probas.is_cuda:
        raise RuntimeError('preworld_b200.losses needs CUDA tensors '
                           '(there is no CPU fallback)')
    rows = _rows(probas.float())
    t = _labels(labels)
    cam = None if camera_mask is None else \
        camera_mask.reshape(-1).to(torch.uint8).contiguous()
    ign = -1 if ignore is None else int(ignore)
    return _Lovasz.apply(rows, bool(from_logits), t, cam, ign)


class _Focal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rows, target, camera_mask, class_weights, radial, depth,
                gamma, alpha, loss_weight, ignore_index):
        loss, sums = ops.focal_loss(rows, target, camera_mask, class_weights, radial,
                                    depth, gamma, alpha, loss_weight, ignore_index)
        ctx.args = (camera_mask, radial, depth, gamma, alpha, loss_weight, ignore_index)
        ctx.save_for_backward(rows, target, class_weights, sums)
        return loss[0].clone()

    @staticmethod
    def backward(ctx, g):
        rows, target, class_weights, sums = ctx.saved_tensors
        camera_mask, radial, depth, gamma, alpha, loss_weight, ignore_index = ctx.args
        grad = ops.focal_loss_grad(rows, target, camera_mask, class_weights, radial,
                                   depth, gamma, alpha, loss_weight, sums, ignore_index)
        return (grad * g,) + (None,) * 9


class CustomFocalLoss(torch.nn.Module):
    """mmdet3d/models/loss_utils/focal_loss.py:162-273 (registered as
    ``CustomFocalLoss``; ``PreWorld`` builds it when ``use_focal_loss=True``, the
    default, preworld.py:43,116-117): sigmoid focal loss weighted per class and by
    the distance of the voxel column from the grid centre.  The centre-distance
    map follows the input's H x W (the reference hard-codes 200 x 200)."""

    def __init__(self, use_sigmoid=True, gamma=2.0, alpha=0.25, reduction='mean',
                 loss_weight=100.0, activated=False):
        super().__init__()
        if not use_sigmoid or activated:
            raise NotImplementedError('CustomFocalLoss: only sigmoid logits '
                                      '(the configuration preworld.py:117 builds)')
        self.gamma, self.alpha = gamma, alpha
        self.reduction, self.loss_weight = reduction, loss_weight
        self._radial = {}

    def radial(self, H, W, device):
        key = (H, W, str(device))
        if key not in self._radial:
            xy, yx = torch.meshgrid([torch.arange(H) - H / 2, torch.arange(W) - W / 2],
                                    indexing='ij')
            c = torch.norm(torch.stack([xy, yx], 2), 2, -1)
            self._radial[key] = (c / c.max() + 1).reshape(-1).to(device).contiguous()
        return self._radial[key]

    def forward(self, pred, target, weight=None, avg_factor=None, ignore_index=255,
                reduction_override=None, camera_mask=None):
        if not pred.is_cuda:
            raise RuntimeError('preworld_b200.losses needs CUDA tensors '
                               '(there is no CPU fallback)')
        B, H, W, D = target.shape
        rows = _rows(pred.float())
        cam = None if camera_mask is None else \
            camera_mask.reshape(-1).to(torch.uint8).contiguous()
        cw = (weight if weight is not None else torch.ones(pred.shape[1])) \
            .to(pred.device).float().contiguous()
        return _Focal.apply(rows, _labels(target), cam, cw, self.radial(H, W, pred.device),
                            D, float(self.gamma), float(self.alpha),
                            float(self.loss_weight), int(ignore_index))
