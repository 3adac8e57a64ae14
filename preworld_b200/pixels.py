"""Device-side pixel pipeline of ``PrepareImageInputs`` (reference
mmdet3d/datasets/pipelines/loading.py:847-854 ``mmlabNormalize``, :954-961
``img_transform_core``, :925-952 ``img_transform``, :974-1000 ``sample_augmentation``).

The reference turns every decoded camera JPEG into a network input on the CPU with four PIL
calls -- ``resize`` (antialiased bicubic, 8-bit fixed point), ``crop``, ``transpose``,
``rotate`` (nearest) -- followed by mmcv's ``imnormalize`` and uploads fp32 CHW tensors.
Here the decoded uint8 HWC image is uploaded as it is and two kernels (csrc/pixels.cu) produce
the same fp32 CHW tensor on the device, bit for bit:

    pw_resample_rows_u8     horizontal pass of PIL's separable resampling (uint8 -> uint8)
    pw_resample_view_norm   vertical pass evaluated only at the pixels the view keeps
                            (crop, flip, nearest rotation resolved as index arithmetic),
                            channel swap + normalisation table

This module is the host side: the view (resize / crop / flip / rotate) of the test and the
training branch, the post-homography the view transformer consumes (``post_rot``,
``post_tran``), and the integer coefficient tables of the two resampling passes.
"""
import math

import numpy as np
import torch

from . import _lib
from .ops import _ptr, _require_cuda, _stream, check

MEAN = (123.675, 116.28, 103.53)          # loading.py:849-850
STD = (58.395, 57.12, 57.375)
_PRECISION_BITS = 32 - 8 - 2              # PIL's 8 bit-per-channel resampling


class View:
    """One camera's image view: the resized size, the crop box in the resized image, the
    horizontal flip and the rotation in degrees (what loading.py:974-1000 samples)."""

    def __init__(self, src_hw, scale, crop, flip=False, rotate=0.0):
        h, w = src_hw
        self.src_hw = (int(h), int(w))
        self.scale = float(scale)
        self.resized_wh = (int(w * self.scale), int(h * self.scale))
        self.crop = tuple(int(v) for v in crop)                  # x0, y0, x1, y1
        self.flip = bool(flip)
        self.rotate = float(rotate)

    @property
    def out_hw(self):
        x0, y0, x1, y1 = self.crop
        return (y1 - y0, x1 - x0)

    @classmethod
    def for_test(cls, src_hw, data_config, flip=None, scale=None):
        """The deterministic branch (is_train=False): fit the width, keep the bottom of the
        image (crop_h), centre horizontally."""
        h, w = src_hw
        fh, fw = data_config['input_size']
        s = float(fw) / float(w)
        s += scale if scale is not None else data_config.get('resize_test', 0.0)
        nw, nh = int(w * s), int(h * s)
        top = int((1 - np.mean(data_config['crop_h'])) * nh) - fh
        left = int(max(0, nw - fw) / 2)
        return cls(src_hw, s, (left, top, left + fw, top + fh), bool(flip), 0.0)

    @classmethod
    def for_train(cls, src_hw, data_config):
        """The random branch; draws from numpy's global generator in the reference's order
        (resize, crop_h, crop_w, flip, rot) so a seeded run samples the same views."""
        h, w = src_hw
        fh, fw = data_config['input_size']
        s = float(fw) / float(w) + np.random.uniform(*data_config['resize'])
        nw, nh = int(w * s), int(h * s)
        top = int((1 - np.random.uniform(*data_config['crop_h'])) * nh) - fh
        left = int(np.random.uniform(0, max(0, nw - fw)))
        flip = data_config['flip'] and np.random.choice([0, 1])
        rot = np.random.uniform(*data_config['rot'])
        return cls(src_hw, s, (left, top, left + fw, top + fh), bool(flip), rot)

    def post_homography(self):
        """-> (post_rot [3,3], post_tran [3]) fp32: pixel coordinates of the source image ->
        pixel coordinates of the network input (loading.py:925-952, 1059-1063), evaluated
        with the reference's float32 operation order."""
        x0, y0, x1, y1 = self.crop
        rot = torch.eye(2) * self.scale
        tran = torch.zeros(2) - torch.tensor([x0, y0], dtype=torch.float32)
        if self.flip:
            mirror = torch.tensor([[-1., 0.], [0., 1.]])
            rot = mirror @ rot
            tran = mirror @ tran + torch.tensor([float(x1 - x0), 0.])
        a = self.rotate / 180 * np.pi
        spin = torch.tensor([[np.cos(a), np.sin(a)], [-np.sin(a), np.cos(a)]],
                            dtype=torch.float32)
        half = torch.tensor([float(x1 - x0), float(y1 - y0)]) / 2
        rot3, tran3 = torch.eye(3), torch.zeros(3)
        rot3[:2, :2] = spin @ rot
        tran3[:2] = spin @ tran + (spin @ (-half) + half)
        return rot3, tran3

    def rotation_fixed_point(self):
        """The 16.16 fixed-point inverse map PIL's nearest-neighbour ``rotate`` walks
        (Image.rotate builds the matrix about the image centre with cos / sin rounded to 15
        digits; the affine transform adds the half-pixel offset and rounds each entry with
        floor(v * 65536 + 0.5)).  None for a rotation by 0 (PIL returns a copy)."""
        if self.rotate == 0:
            return None
        if self.rotate % 90 == 0:
            raise NotImplementedError('rotations by multiples of 90 degrees are transposes in PIL')
        h, w = self.out_hw
        cx, cy = w / 2, h / 2
        ang = -math.radians(self.rotate)
        m = [round(math.cos(ang), 15), round(math.sin(ang), 15), 0.0,
             round(-math.sin(ang), 15), round(math.cos(ang), 15), 0.0]
        m[2] = m[0] * -cx + m[1] * -cy + m[2] + cx
        m[5] = m[3] * -cx + m[4] * -cy + m[5] + cy
        m[2] += m[0] * 0.5 + m[1] * 0.5
        m[5] += m[3] * 0.5 + m[4] * 0.5
        return [int(math.floor(v * 65536.0 + 0.5)) for v in m]


def _bicubic(x):
    x = np.abs(x)
    near = ((-0.5 + 2.0) * x - (-0.5 + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * -0.5
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


def resample_tables(in_size, out_size):
    """Integer taps of one axis of PIL's antialiased bicubic resampling (the default filter
    of ``Image.resize``) from in_size to out_size pixels: for output pixel i the window
    starts at first[i], has count[i] taps, taps[i, :count[i]] scaled by 2^22.  Same double
    precision operations in the same order as PIL's coefficient set-up, so the integers are
    the ones PIL uses."""
    scale = float(in_size) / out_size
    fscale = max(scale, 1.0)
    support = 2.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    inv = 1.0 / fscale
    centre = 0.0 + (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    first = np.maximum((centre - support + 0.5).astype(np.int64), 0)
    last = np.minimum((centre + support + 0.5).astype(np.int64), in_size)
    count = last - first
    k = np.arange(ksize, dtype=np.float64)[None, :]
    w = _bicubic((k + first[:, None] - centre[:, None] + 0.5) * inv)
    w = np.where(k < count[:, None], w, 0.0)
    total = np.zeros(out_size)
    for j in range(ksize):                        # PIL sums the taps left to right
        total = total + w[:, j]
    w = np.where(total[:, None] != 0.0, w / np.where(total == 0.0, 1.0, total)[:, None], w)
    q = w * float(1 << _PRECISION_BITS)
    taps = np.where(q < 0, (-0.5 + q).astype(np.int64), (0.5 + q).astype(np.int64))
    return first.astype(np.int32), count.astype(np.int32), taps.astype(np.int32)


def normalize_table():
    """[3,256] fp32: network-input value of output channel c for the 8-bit value v.  mmcv's
    imnormalize hands cv2 a float32 image and float64 scalars (mean, 1 / std, both widened
    from float32 arrays); cv2 then subtracts and multiplies in double and rounds once."""
    v = np.arange(256, dtype=np.float64)[None, :]
    mean = np.float64(np.array(MEAN, dtype=np.float32))[:, None]
    stdinv = (1 / np.float64(np.array(STD, dtype=np.float32)))[:, None]
    return np.ascontiguousarray(((v - mean) * stdinv).astype(np.float32))


class PixelPipeline:
    """uint8 HWC camera image on the device -> the fp32 CHW network input of one view."""

    def __init__(self, view, device):
        self.view, self.device = view, torch.device(device)
        h, w = view.src_hw
        nw, nh = view.resized_wh
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        fx, cx, tx = resample_tables(w, nw)
        fy, cy, ty = resample_tables(h, nh)
        self.x_first, self.x_count, self.x_taps = dev(fx), dev(cx), dev(tx)
        self.y_first, self.y_count, self.y_taps = dev(fy), dev(cy), dev(ty)
        # rows of the resized image the crop can touch -> source rows the first pass needs
        x0, y0, x1, y1 = view.crop
        if view.rotate == 0:
            ry0, ry1 = max(y0, 0), min(y1, nh)
        else:
            ry0, ry1 = 0, nh
        if ry1 > ry0:
            self.row0 = int(fy[ry0])
            self.rows = int(fy[ry1 - 1] + cy[ry1 - 1]) - self.row0
        else:
            self.row0, self.rows = 0, 1
        self.tmp = torch.empty((self.rows, nw, 3), device=self.device, dtype=torch.uint8)
        fixed = view.rotation_fixed_point()
        self.affine = (ctypes_int6(fixed) if fixed is not None else None)
        self.lut = torch.from_numpy(normalize_table()).to(self.device)
        # PIL skips a pass whose axis keeps its size
        self.identity_x, self.identity_y = nw == w, nh == h

    def __call__(self, img_u8, out=None):
        """img_u8 [H,W,3] uint8 (RGB as PIL decodes it) -> [3,fH,fW] fp32, the tensor
        ``mmlabNormalize(img_transform_core(img, ...))`` returns."""
        _require_cuda(img_u8, out)
        v = self.view
        h, w = v.src_hw
        assert img_u8.dtype == torch.uint8 and tuple(img_u8.shape) == (h, w, 3) \
            and img_u8.stride(2) == 1 and img_u8.stride(1) == 3
        fh, fw = v.out_hw
        if out is None:
            out = torch.empty((3, fh, fw), device=img_u8.device, dtype=torch.float32)
        assert out.shape == (3, fh, fw) and out.is_contiguous()
        nw, nh = v.resized_wh
        L = _lib.lib()
        check(L.pw_resample_rows_u8(
            _ptr(img_u8), img_u8.stride(0), h, w, self.row0, self.rows, _ptr(self.x_first),
            _ptr(self.x_count), _ptr(self.x_taps), self.x_taps.shape[1], int(self.identity_x),
            _ptr(self.tmp), nw, _stream()), 'pw_resample_rows_u8')
        x0, y0, x1, y1 = v.crop
        check(L.pw_resample_view_norm(
            _ptr(self.tmp), self.row0, self.rows, nw, nh, _ptr(self.y_first), _ptr(self.y_count),
            _ptr(self.y_taps), self.y_taps.shape[1], int(self.identity_y), x0, y0,
            int(v.flip), self.affine, _ptr(self.lut), _ptr(out), fh, fw,
            _stream()), 'pw_resample_view_norm')
        return out


def ctypes_int6(vals):
    import ctypes
    return (ctypes.c_int * 6)(*[int(v) for v in vals])
