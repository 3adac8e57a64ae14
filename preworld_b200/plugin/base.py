"""Shared host-side machinery of the plugin modules.

Modules keep their parameters in ``torch.nn`` containers with the reference's
attribute names (so reference checkpoints load by key) but never call those
containers' ``forward``: at first use each module *packs* its weights into the
kernel layout (tap-major [K, Cout], BatchNorm folded into a per-channel
affine) and from then on only launches C-ABI kernels.
"""
import torch
import torch.nn as nn

from .. import ops


class BaseModule(nn.Module):
    """mmcv.runner.BaseModule surface used by the configs (init_cfg /
    init_weights) + pack caching."""

    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg
        self._pw_packs = None
        self._pw_key = None

    def init_weights(self):
        pass

    # -- pack cache ---------------------------------------------------------
    # names of the child modules whose tensors this module packs itself
    # (None == everything below this module)
    _pack_children = None

    def _pack_key(self):
        ver = 0
        dev = None
        mods = [self] if self._pack_children is None else \
            [getattr(self, n) for n in self._pack_children if hasattr(self, n)]
        for m in mods:
            for t in m.parameters():
                ver += t._version
                dev = t.device
            for t in m.buffers():
                ver += t._version
        return (dev, ver, self.training)

    def packs(self):
        key = self._pack_key()
        if self._pw_packs is None or self._pw_key != key:
            if self.training:
                raise RuntimeError(
                    f'{type(self).__name__}: the B200 path is forward-only '
                    '(eval mode); call .eval() first')
            with torch.no_grad():
                self._pw_packs = self._build_packs()
            self._pw_key = key
        return self._pw_packs

    def _build_packs(self):
        raise NotImplementedError

    def _apply(self, fn, *a, **k):
        self._pw_packs = None
        return super()._apply(fn, *a, **k)


def bn_tuple(bn):
    return (bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)


def pack_conv(conv, bn=None, **kw):
    """nn.ConvNd (+ following BatchNorm) -> ops.PackedConv."""
    return ops.PackedConv(conv.weight, conv.bias,
                          bn_tuple(bn) if bn is not None else None,
                          stride=conv.stride, padding=conv.padding,
                          dilation=conv.dilation, **kw)


def pack_linear(lin, **kw):
    return ops.PackedConv(lin.weight, lin.bias, None, **kw)


_CONV = {None: nn.Conv2d, 'Conv2d': nn.Conv2d, 'Conv3d': nn.Conv3d,
         'Conv1d': nn.Conv1d}
_NORM = {'BN': nn.BatchNorm2d, 'BN1d': nn.BatchNorm1d, 'BN2d': nn.BatchNorm2d,
         'BN3d': nn.BatchNorm3d,
         # SyncBN only differs in training; plain BatchNorm3d holds the same
         # parameter / buffer keys (OccHead uses it on 5-D tensors,
         # preworld-7frame-finetune.py:39)
         'SyncBN': nn.BatchNorm3d}


def build_conv_layer(cfg, *args, **kwargs):
    """mmcv.cnn.build_conv_layer for the conv types on the path."""
    cfg_ = {} if cfg is None else dict(cfg)
    typ = cfg_.pop('type', None)
    if typ not in _CONV:
        raise KeyError(f'conv type {typ} is not on the PreWorld forward path')
    return _CONV[typ](*args, **kwargs, **cfg_)


def build_norm_layer(cfg, num_features, postfix='', dims=None):
    """mmcv.cnn.build_norm_layer: returns (name, layer); 'bn' abbreviation."""
    cfg_ = dict(cfg)
    typ = cfg_.pop('type')
    requires_grad = cfg_.pop('requires_grad', True)
    cfg_.setdefault('eps', 1e-5)
    cls = _NORM[typ]
    if typ == 'BN' and dims == 3:
        cls = nn.BatchNorm3d
    layer = cls(num_features, **cfg_)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return 'bn' + str(postfix), layer


class ConvModule(nn.Module):
    """Parameter container with mmcv 1.6.0 ConvModule's keys and semantics
    (conv -> norm -> act, ``bias='auto'`` == no conv bias when a norm follows,
    default ``act_cfg=dict(type='ReLU')``).  Used at reference
    backbones/resnet.py:92-111,153-162, necks/lss_fpn.py:120-129,
    necks/fpn.py:111-131, detectors/preworld.py:72-79."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1,
                 padding=0, dilation=1, groups=1, bias='auto', conv_cfg=None,
                 norm_cfg=None, act_cfg=dict(type='ReLU'), inplace=True,
                 **kwargs):
        super().__init__()
        assert groups == 1
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == 'auto':
            bias = not self.with_norm
        self.conv = build_conv_layer(conv_cfg, in_channels, out_channels,
                                     kernel_size, stride=stride,
                                     padding=padding, dilation=dilation,
                                     bias=bias)
        if self.with_norm:
            dims = 3 if isinstance(self.conv, nn.Conv3d) else 2
            _, self.bn = build_norm_layer(norm_cfg, out_channels, dims=dims)
        if self.with_activation:
            if act_cfg['type'] != 'ReLU':
                raise KeyError('only ReLU ConvModules are on the path')
        self.act = 'relu' if self.with_activation else None

    def pack(self, **kw):
        return pack_conv(self.conv, self.bn if self.with_norm else None, **kw)
