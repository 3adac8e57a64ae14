"""Device-side pixel pipeline (csrc/pixels.cu through the C ABI) against the reference's own
arithmetic -- PIL's resize / crop / transpose / rotate and cv2's normalisation, called by
oracle/pixel_ref.py exactly as loading.py:847-854,954-961 calls them.  Bar: bit-exact (8-bit
fixed-point resampling, table-driven normalisation)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import pixel_ref                          # noqa: E402  (checker only)
from preworld_b200.pixels import PixelPipeline, View  # noqa: E402

DATA_CONFIG = dict(cams=['a'] * 6, Ncams=6, input_size=(256, 704), src_size=(900, 1600),
                   resize=(-0.06, 0.11), rot=(-5.4, 5.4), flip=True, crop_h=(0.0, 0.0),
                   resize_test=0.0)


def _views(hw, cfg):
    vs = [View.for_test(hw, cfg), View.for_test(hw, cfg, flip=True, scale=0.04)]
    for seed in (1, 2, 3, 4):
        np.random.seed(seed)
        vs.append(View.for_train(hw, cfg))
    return vs


def test_nuscenes_views_match_pil_bit_exact():
    """1600x900 camera image -> 3x256x704 network input: the deterministic test view, a
    flipped + rescaled test view and four seeded training views (resize, crop, flip, rotation)."""
    img = pixel_ref.synthetic_photo(900, 1600, 0)
    dev = torch.from_numpy(img).cuda()
    for v in _views((900, 1600), DATA_CONFIG):
        got = PixelPipeline(v, 'cuda')(dev)
        want = pixel_ref.network_input(img, v)
        assert torch.equal(got.cpu(), want), vars(v)


def test_shipped_resolution_and_edge_geometries():
    """512x1408 (the shipped config's input size), enlargement, odd sizes, a crop box
    reaching beyond the resized image (zero fill), no resize at all."""
    cfg = dict(DATA_CONFIG, input_size=(512, 1408))
    img = pixel_ref.synthetic_photo(900, 1600, 5)
    dev = torch.from_numpy(img).cuda()
    for v in _views((900, 1600), cfg)[:4]:
        assert torch.equal(PixelPipeline(v, 'cuda')(dev).cpu(), pixel_ref.network_input(img, v))
    small = pixel_ref.synthetic_photo(97, 131, 1)
    sdev = torch.from_numpy(small).cuda()
    for v in (View((97, 131), 1.37, (-5, 10, 155, 140), False, 0.0),
              View((97, 131), 0.61, (3, -7, 75, 60), True, 3.3),
              View((97, 131), 1.0, (0, 0, 131, 97), False, -4.9),
              View((97, 131), 1.0, (10, 5, 100, 80), True, 0.0)):
        got = PixelPipeline(v, 'cuda')(sdev)
        assert torch.equal(got.cpu(), pixel_ref.network_input(small, v)), vars(v)


def test_pipeline_feeds_the_model_input_layout():
    """18 camera images of one sample written straight into the loader's camera-major batch
    [1, 18, 3, 256, 704] (out= slices), equal to the host pipeline's tensor."""
    v = View.for_test((900, 1600), DATA_CONFIG)
    pipe = PixelPipeline(v, 'cuda')
    batch = torch.empty((1, 18, 3, 256, 704), device='cuda')
    imgs = [pixel_ref.synthetic_photo(900, 1600, 10 + i) for i in range(3)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    devs = [torch.from_numpy(imgs[i % 3]).cuda() for i in range(18)]
    pipe(devs[0], out=batch[0, 0])
    torch.cuda.synchronize()
    t0.record()
    for i in range(18):
        pipe(devs[i], out=batch[0, i])
    t1.record()
    torch.cuda.synchronize()
    print(f'pixel pipeline: {t0.elapsed_time(t1) / 18 * 1e3:.1f} us per 1600x900 image '
          f'({18 * 900 * 1600 * 3 / 1e6:.0f} MB uint8 in, {batch.numel() * 4 / 1e6:.0f} MB fp32 out)')
    for i in (0, 7, 17):
        assert torch.equal(batch[0, i].cpu(), pixel_ref.network_input(imgs[i % 3], v))
    rot, tran = v.post_homography()
    assert torch.allclose(rot, torch.diag(torch.tensor([0.44, 0.44, 1.0])))
    assert torch.equal(tran, torch.tensor([0., -140., 0.]))
