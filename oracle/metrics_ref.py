"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of
Metric_mIoU (mmdet3d/datasets/occ_metrics.py:93-185) used to check
preworld_b200.metrics / pw_occ_confusion.  Parity pin: the reference ships no
test for it; the restatement is pinned against the reference's own file --
occ_metrics.py imported by path in the build container (oracle/make_metrics_golden.py)
and the committed tests/golden/occ_metrics.json it produced (confusion matrices
bit-equal, mIoU / IoU to the reference's rounding; tests/test_oracle.py) -- and
against a brute-force loop."""
import numpy as np


def hist_info(n_cl, pred, gt):                   # occ_metrics.py:93-113
    k = (gt >= 0) & (gt < n_cl)
    return np.bincount(n_cl * gt[k].astype(int) + pred[k].astype(int),
                       minlength=n_cl ** 2).reshape(n_cl, n_cl)


def per_class_iu(hist):                          # occ_metrics.py:115-117
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist))


class MetricRef:
    def __init__(self, num_classes=18, use_lidar_mask=False,
                 use_image_mask=False):
        self.num_classes = num_classes
        self.use_lidar_mask, self.use_image_mask = use_lidar_mask, use_image_mask
        self.hist = np.zeros((num_classes, num_classes))
        self.occ_hist = np.zeros((2, 2))
        self.cnt = 0

    def add_batch(self, pred, gt, mask_lidar=None, mask_camera=None):
        self.cnt += 1                            # occ_metrics.py:133-157
        if self.use_image_mask:
            gt, pred = gt[mask_camera], pred[mask_camera]
        elif self.use_lidar_mask:
            gt, pred = gt[mask_lidar], pred[mask_lidar]
        self.hist += hist_info(self.num_classes, pred.flatten(), gt.flatten())
        free = 17
        occ_pred = np.zeros_like(pred); occ_pred[pred != free] = 1
        occ_gt = np.zeros_like(gt); occ_gt[gt != free] = 1
        self.occ_hist += hist_info(2, occ_pred.flatten(), occ_gt.flatten())

    def count_miou(self):
        m = per_class_iu(self.hist)
        return m, round(np.nanmean(m[:self.num_classes - 1]) * 100, 2)

    def count_iou(self):
        i = per_class_iu(self.occ_hist)
        return i, round(i[-1] * 100, 2)


class MetricTemporalRef:
    """occ_metrics.py:413-596: per-horizon accumulation (idx 0, 2, 4, 6 ->
    0s..3s, prediction index idx // 2)."""

    def __init__(self, num_classes=18, use_lidar_mask=False,
                 use_image_mask=False):
        self.num_classes = num_classes
        self.m = {idx: MetricRef(num_classes, use_lidar_mask, use_image_mask)
                  for idx in (0, 2, 4, 6)}

    def add_batch(self, preds, gts, mask_lidar, mask_camera):
        for idx in gts.keys():
            if idx in self.m:
                self.m[idx].add_batch(preds[idx // 2], gts[idx],
                                      mask_lidar[idx], mask_camera[idx])

    def count_miou(self):
        return [self.m[idx].count_miou()[1] for idx in (2, 4, 6)]

    def count_iou(self):
        return [self.m[idx].count_iou()[1] for idx in (2, 4, 6)]
