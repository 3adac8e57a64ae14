"""Occupancy metrics accumulated on the device.

``Metric_mIoU`` mirrors mmdet3d/datasets/occ_metrics.py:52-185 (same
constructor flags, ``add_batch`` / ``count_miou`` / ``count_iou``), but takes
the uint8 occupancy grids as CUDA tensors and keeps the 18x18 and 2x2 confusion
matrices on the GPU (``pw_occ_confusion``): the eval loop of
mmdet3d/apis/test.py:62-103 no longer needs a ``.cpu().numpy()`` per sample;
only ``count_*`` reads 328 integers back."""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check

CLASS_NAMES = ['others', 'barrier', 'bicycle', 'bus', 'car',
               'construction_vehicle', 'motorcycle', 'pedestrian',
               'traffic_cone', 'trailer', 'truck', 'driveable_surface',
               'other_flat', 'sidewalk', 'terrain', 'manmade', 'vegetation',
               'free']


def _u8(t, dev):
    t = torch.as_tensor(t)
    if t.dtype == torch.bool:
        t = t.to(torch.uint8)
    if t.dtype != torch.uint8:
        raise TypeError('occupancy grids / masks are uint8 (or bool) tensors')
    t = t.to(dev).contiguous()
    if not t.is_cuda:
        raise RuntimeError('preworld_b200 metrics run on CUDA tensors only')
    return t


class Metric_mIoU:
    def __init__(self, save_dir='.', num_classes=18, use_lidar_mask=False,
                 use_image_mask=False, device='cuda'):
        self.class_names = CLASS_NAMES
        self.occ_names = ['free', 'occupied']
        self.num_classes = num_classes
        self.use_lidar_mask = use_lidar_mask
        self.use_image_mask = use_image_mask
        self.device = torch.device(device)
        self.hist_dev = torch.zeros(num_classes * num_classes, dtype=torch.int64,
                                    device=self.device)
        self.occ_hist_dev = torch.zeros(4, dtype=torch.int64, device=self.device)
        self.cnt = 0

    # occ_metrics.py:133-157
    def add_batch(self, semantics_pred, semantics_gt, mask_lidar=None,
                  mask_camera=None):
        self.cnt += 1
        pred = _u8(semantics_pred, self.device)
        gt = _u8(semantics_gt, self.device)
        assert pred.shape == gt.shape
        mask = None
        if self.use_image_mask:
            mask = _u8(mask_camera, self.device)
        elif self.use_lidar_mask:
            mask = _u8(mask_lidar, self.device)
        p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        check(_lib.lib().pw_occ_confusion(
            p(pred), p(gt), p(mask), pred.numel(), self.num_classes,
            self.num_classes - 1, p(self.hist_dev), p(self.occ_hist_dev),
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
            'pw_occ_confusion')

    def all_reduce(self, group=None):
        """Sum the confusion matrices over the ranks of a distributed evaluation (one
        all-reduce of 328 counters instead of gathering every prediction to rank 0,
        mmdet3d/apis/test.py:113-117,165-195).  Call once, after the last add_batch."""
        from .parallel import reduce_confusion
        reduce_confusion([self.hist_dev, self.occ_hist_dev], group)
        return self

    @property
    def hist(self):
        n = self.num_classes
        return self.hist_dev.cpu().numpy().reshape(n, n).astype(np.float64)

    @property
    def occ_hist(self):
        return self.occ_hist_dev.cpu().numpy().reshape(2, 2).astype(np.float64)

    @staticmethod
    def per_class_iu(hist):                      # occ_metrics.py:115-117
        with np.errstate(divide='ignore', invalid='ignore'):
            return np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist))

    def count_miou(self):                        # occ_metrics.py:159-175
        mIoU = self.per_class_iu(self.hist)
        res = round(np.nanmean(mIoU[:self.num_classes - 1]) * 100, 2)
        return self.class_names, mIoU, self.cnt, res

    def count_iou(self):                         # occ_metrics.py:177-185
        IoU = self.per_class_iu(self.occ_hist)
        return self.occ_names, IoU, self.cnt, round(IoU[-1] * 100, 2)


class Metric_mIoU_Temporal:
    """mmdet3d/datasets/occ_metrics.py:413-596 on the device: one
    (18x18, 2x2) confusion-matrix pair per horizon 0s/1s/2s/3s; ``add_batch``
    takes the forecast list ``semantics_pred[k]`` (k = idx // 2) and the ground
    truth / mask dicts keyed idx = 0, 2, 4, 6 exactly as the reference."""

    def __init__(self, save_dir='.', num_classes=18, use_lidar_mask=False,
                 use_image_mask=False, device='cuda'):
        self.class_names = CLASS_NAMES
        self.occ_names = ['free', 'occupied']
        self.num_classes = num_classes
        self.cnt = 0
        self._m = {idx: Metric_mIoU(save_dir, num_classes, use_lidar_mask,
                                    use_image_mask, device)
                   for idx in (0, 2, 4, 6)}

    def add_batch(self, semantics_pred, semantics_gt_temp, mask_lidar_temp,
                  mask_camera_temp):
        self.cnt += 1
        for idx in semantics_gt_temp.keys():
            if idx not in self._m:
                continue                         # the reference ignores other keys
            self._m[idx].add_batch(
                semantics_pred[idx // 2], semantics_gt_temp[idx],
                mask_lidar_temp[idx] if mask_lidar_temp is not None else None,
                mask_camera_temp[idx] if mask_camera_temp is not None else None)

    def all_reduce(self, group=None):
        from .parallel import reduce_confusion
        reduce_confusion([t for m in self._m.values() for t in (m.hist_dev, m.occ_hist_dev)],
                         group)
        return self

    def _hist(self, k):
        return self._m[2 * k].hist

    hist_0s = property(lambda self: self._hist(0))
    hist_1s = property(lambda self: self._hist(1))
    hist_2s = property(lambda self: self._hist(2))
    hist_3s = property(lambda self: self._hist(3))
    occ_hist_0s = property(lambda self: self._m[0].occ_hist)
    occ_hist_1s = property(lambda self: self._m[2].occ_hist)
    occ_hist_2s = property(lambda self: self._m[4].occ_hist)
    occ_hist_3s = property(lambda self: self._m[6].occ_hist)

    def count_miou(self):                        # occ_metrics.py:541-571
        n = self.num_classes
        per = [Metric_mIoU.per_class_iu(self._hist(k)) for k in (1, 2, 3)]
        return per[0], [round(np.nanmean(m[:n - 1]) * 100, 2) for m in per]

    def count_iou(self):                         # occ_metrics.py:573-596
        return [round(Metric_mIoU.per_class_iu(self._m[idx].occ_hist)[-1] * 100, 2)
                for idx in (2, 4, 6)]
