"""Host side of the input format: the image-view
augmentation parameters and the ``post_rot`` / ``post_tran`` matrices the view
transformer consumes, as ``PrepareImageInputs`` builds them
(mmdet3d/datasets/pipelines/loading.py:925-1000,1057-1065).  Pure host logic (a
few scalars per camera): numpy / torch on the CPU, same random draws in the same
order as the reference, so a seeded loader produces identical matrices.  The
pixel resampling itself (PIL resize / crop / rotate, loading.py:954-961) is not
part of this module."""
import numpy as np
import torch


def sample_augmentation(H, W, data_config, is_train=False, flip=None, scale=None):
    """loading.py:974-1000 -> (resize, resize_dims, crop, flip, rotate).  The
    training branch draws from ``np.random`` exactly like the reference
    (resize, crop_h, crop_w, flip, rotate -- in that order)."""
    fH, fW = data_config['input_size']
    if is_train:
        resize = float(fW) / float(W)
        resize += np.random.uniform(*data_config['resize'])
        resize_dims = (int(W * resize), int(H * resize))
        newW, newH = resize_dims
        crop_h = int((1 - np.random.uniform(*data_config['crop_h'])) * newH) - fH
        crop_w = int(np.random.uniform(0, max(0, newW - fW)))
        crop = (crop_w, crop_h, crop_w + fW, crop_h + fH)
        flip = data_config['flip'] and np.random.choice([0, 1])
        rotate = np.random.uniform(*data_config['rot'])
    else:
        resize = float(fW) / float(W)
        resize += scale if scale is not None else data_config.get('resize_test', 0.0)
        resize_dims = (int(W * resize), int(H * resize))
        newW, newH = resize_dims
        crop_h = int((1 - np.mean(data_config['crop_h'])) * newH) - fH
        crop_w = int(max(0, newW - fW) / 2)
        crop = (crop_w, crop_h, crop_w + fW, crop_h + fH)
        flip = False if flip is None else flip
        rotate = 0
    return resize, resize_dims, crop, flip, rotate


def _rot(h):
    return torch.Tensor([[np.cos(h), np.sin(h)], [-np.sin(h), np.cos(h)]])


def aug_matrices(resize, crop, flip, rotate):
    """The post-homography of loading.py:940-952 applied to identity, widened to
    3x3 / 3 as in :1057-1061 -> (post_rot [3,3], post_tran [3])."""
    post_rot = torch.eye(2) * resize
    post_tran = torch.zeros(2) - torch.Tensor(crop[:2])
    if flip:
        A = torch.Tensor([[-1, 0], [0, 1]])
        b = torch.Tensor([crop[2] - crop[0], 0])
        post_rot = A.matmul(post_rot)
        post_tran = A.matmul(post_tran) + b
    A = _rot(rotate / 180 * np.pi)
    b = torch.Tensor([crop[2] - crop[0], crop[3] - crop[1]]) / 2
    b = A.matmul(-b) + b
    post_rot = A.matmul(post_rot)
    post_tran = A.matmul(post_tran) + b
    rot3, tran3 = torch.eye(3), torch.zeros(3)
    rot3[:2, :2] = post_rot
    tran3[:2] = post_tran
    return rot3, tran3


def camera_augmentations(image_sizes, data_config, is_train=False, flip=None,
                         scale=None, num_frames=1):
    """One augmentation per camera (shared by that camera's adjacent frames,
    loading.py:1117-1121) -> post_rots [num_frames*N,3,3], post_trans
    [num_frames*N,3] in the frame-major order ``get_inputs`` returns, plus the
    per-camera (resize_dims, crop, flip, rotate) the pixel pipeline needs."""
    rots, trans, params = [], [], []
    for (H, W) in image_sizes:
        resize, resize_dims, crop, flip_c, rotate = sample_augmentation(
            H, W, data_config, is_train, flip, scale)
        r, t = aug_matrices(resize, crop, flip_c, rotate)
        rots.append(r)
        trans.append(t)
        params.append(dict(resize=resize, resize_dims=resize_dims, crop=crop,
                           flip=flip_c, rotate=rotate))
    return (torch.stack(rots * num_frames), torch.stack(trans * num_frames), params)
