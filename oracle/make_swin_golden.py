"""Generates tests/golden/tiny_swin.npz by running the REFERENCE's own SwinTransformer
(backbones/swin.py) and FPN_LSS (necks/lss_fpn.py) from their files (oracle/swin_shim.py) on
a seeded input with seeded weights.  Run in the build container:

    python oracle/make_swin_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import swin_ref, swin_shim  # noqa: E402


def reference_outputs():
    swin, fpn = swin_shim.load()
    bb = swin.SwinTransformer(with_cp=False, **swin_shim.TINY_SWIN)
    neck = fpn.FPN_LSS(**swin_shim.TINY_NECK)
    bb.eval()           # SwinTransformer.train() returns None (swin.py:972-976)
    neck.eval()
    swin_shim.seeded_init_(bb, 11)
    swin_shim.seeded_init_(neck, 12)
    x = swin_ref.tiny_input()
    with torch.no_grad():
        outs = bb(x)
        n = neck(outs[1:])
    return bb, neck, x, outs, n


if __name__ == '__main__':
    bb, neck, x, outs, n = reference_outputs()
    path = os.path.join(ROOT, 'tests', 'golden', 'tiny_swin.npz')
    np.savez_compressed(path, stereo=outs[0].numpy(), out2=outs[1].numpy(), out3=outs[2].numpy(),
                        neck=n.numpy(),
                        rel_index=bb.stages[0].blocks[0].attn.w_msa.relative_position_index.numpy(),
                        keys_backbone=np.array(sorted(bb.state_dict().keys())),
                        keys_neck=np.array(sorted(neck.state_dict().keys())))
    print('wrote', path, [tuple(o.shape) for o in outs], tuple(n.shape),
          [float(o.abs().mean()) for o in outs], float(n.abs().mean()))
