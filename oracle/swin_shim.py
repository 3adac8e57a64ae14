"""TEST INFRASTRUCTURE ONLY (build container).  Loads the reference's own
mmdet3d/models/backbones/swin.py (SwinTransformer, :679-976) and FPN_LSS
(necks/lss_fpn.py:13-99) VERBATIM from /root/reference under oracle/ref_shim.py plus the
stand-ins this file adds for what swin.py imports from mmcv 1.6.0 / mmseg (absent from the
container; semantics restated -- the un-pinned assumption of this oracle, as for ConvModule):

  mmcv.cnn.bricks.transformer.FFN    Sequential(Sequential(Linear, act, Dropout), Linear,
                                      Dropout); identity + dropout_layer(layers(x))
  mmcv.cnn.bricks.transformer.build_dropout / DropPath   identity in eval mode
  mmcv.cnn.trunc_normal_init, mmcv.cnn.utils.weight_init.constant_init   (init only)
  mmcv.runner._load_checkpoint, mmcv.runner.base_module.{BaseModule, ModuleList}
  mmseg.ops.resize                    F.interpolate (checkpoint conversion only)
"""
import importlib
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ref_shim
from .swin_ref import TINY_INPUT, TINY_NECK, TINY_SWIN, seeded_init_  # noqa: F401


class _DropPath(nn.Module):
    def __init__(self, drop_prob=0.):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        assert not self.training, 'the oracle runs in eval mode'
        return x


def build_dropout(cfg, default_args=None):
    cfg_ = dict(cfg)
    typ = cfg_.pop('type')
    return _DropPath(**cfg_) if typ == 'DropPath' else nn.Dropout(**cfg_)


class FFN(ref_shim.BaseModule):
    """mmcv/cnn/bricks/transformer.py (1.6.0) FFN."""

    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2,
                 act_cfg=dict(type='ReLU', inplace=True), ffn_drop=0., dropout_layer=None,
                 add_identity=True, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        layers, in_channels = [], embed_dims
        for _ in range(num_fcs - 1):
            layers.append(nn.Sequential(nn.Linear(in_channels, feedforward_channels),
                                        ref_shim.build_activation_layer(act_cfg),
                                        nn.Dropout(ffn_drop)))
            in_channels = feedforward_channels
        layers.append(nn.Linear(feedforward_channels, embed_dims))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = nn.Sequential(*layers)
        self.dropout_layer = build_dropout(dropout_layer) if dropout_layer else nn.Identity()
        self.add_identity = add_identity

    def forward(self, x, identity=None):
        out = self.layers(x)
        if not self.add_identity:
            return self.dropout_layer(out)
        if identity is None:
            identity = x
        return identity + self.dropout_layer(out)


_loaded = None


def load():
    """-> (swin module, lss_fpn module) of the reference, imported from their files."""
    global _loaded
    if _loaded is not None:
        return _loaded
    ref_shim.install()
    mm = sys.modules
    mm['mmcv.cnn'].trunc_normal_init = lambda *a, **k: None
    ref_shim._install('mmcv.cnn.bricks.transformer', FFN=FFN, build_dropout=build_dropout)
    ref_shim._install('mmcv.cnn.bricks.registry', ATTENTION=ref_shim.Registry('attention'))
    ref_shim._install('mmcv.cnn.utils')
    ref_shim._install('mmcv.cnn.utils.weight_init', constant_init=lambda *a, **k: None)
    mm['mmcv.runner']._load_checkpoint = lambda *a, **k: {}
    ref_shim._install('mmcv.runner.base_module', BaseModule=ref_shim.BaseModule,
                      ModuleList=nn.ModuleList)
    ref_shim._install('mmseg.ops', resize=lambda x, size=None, mode='bilinear', **k:
                      F.interpolate(x, size=size, mode=mode))
    ref_shim._install('mmdet3d.utils', get_root_logger=lambda *a, **k: None)
    mm.pop('mmdet3d.models.backbones.swin', None)          # ref_shim's isinstance() stub
    swin = importlib.import_module('mmdet3d.models.backbones.swin')
    fpn = importlib.import_module('mmdet3d.models.necks.lss_fpn')
    _loaded = (swin, fpn)
    return _loaded
