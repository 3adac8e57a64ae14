"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Container-only loader that executes the reference's own ``.py`` files
*verbatim from /root/reference* on CPU.  The reference hard-depends on
mmcv-full 1.6.0 / mmdet 2.24.0 / mmseg / torch_scatter /
torch_efficient_distloss / matplotlib / IPython / tkinter, none of which are
installed here, so this file supplies the thinnest possible stand-ins for the
*containers* those packages provide (Registry, ConvModule, BaseModule,
build_conv/norm_layer, mmdet BasicBlock/Bottleneck/ResNet) and CPU restatements
of the reference's CUDA-only ops (bev_pool_v2, raw2alpha, alpha2weight,
cumdist_thres, segment_coo).  All arithmetic that is *in* the reference is
executed from the reference's own source.

It is used by ``oracle/make_golden.py`` (golden-vector generation) and by
``tests/test_oracle_vs_reference.py`` (skipped when /root/reference is absent,
i.e. on the GPU box).  The semantics restated here follow:

* mmcv 1.6.0 ``mmcv/cnn/bricks/conv_module.py`` (ConvModule: conv -> norm ->
  act, ``bias='auto'`` == no bias when a norm layer follows, default
  ``act_cfg=dict(type='ReLU')``, norm sub-module named by ``build_norm_layer``
  abbreviation: 'bn' for BN/BN1d/BN2d/BN3d/SyncBN, 'gn' for GN).
* mmcv 1.6.0 ``mmcv/utils/registry.py`` (Registry.build: pop ``type``, call).
* mmdet 2.24.0 ``mmdet/models/backbones/resnet.py`` (BasicBlock, Bottleneck,
  ResNet -- same topology and parameter names as torchvision's resnet).
Those packages are pinned in the reference's requirements.txt:14-15.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get('PREWORLD_REFERENCE_ROOT', '/root/reference')


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'mmdet3d'))


# --------------------------------------------------------------------------
# mmcv.utils.Registry
# --------------------------------------------------------------------------
class Registry:
    def __init__(self, name, build_func=None, parent=None, scope=None):
        self._name = name
        self._module_dict = {}
        self.parent = parent

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key):
        if key in self._module_dict:
            return self._module_dict[key]
        if self.parent is not None:
            return self.parent.get(key)
        return None

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            self._module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return _register(module)
        return _register

    def build(self, cfg, default_args=None):
        args = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        typ = args.pop('type')
        cls = typ if not isinstance(typ, str) else self.get(typ)
        if cls is None:
            raise KeyError(f'{typ} is not in the {self._name} registry')
        return cls(**args)


# --------------------------------------------------------------------------
# mmcv.cnn builders
# --------------------------------------------------------------------------
_CONV = {None: nn.Conv2d, 'Conv1d': nn.Conv1d, 'Conv2d': nn.Conv2d,
         'Conv3d': nn.Conv3d, 'Conv': nn.Conv2d,
         'deconv3d': nn.ConvTranspose3d}
_NORM = {'BN': ('bn', nn.BatchNorm2d), 'BN1d': ('bn', nn.BatchNorm1d),
         'BN2d': ('bn', nn.BatchNorm2d), 'BN3d': ('bn', nn.BatchNorm3d),
         'SyncBN': ('bn', nn.SyncBatchNorm), 'GN': ('gn', nn.GroupNorm),
         'LN': ('ln', nn.LayerNorm)}


def build_conv_layer(cfg, *args, **kwargs):
    cfg_ = {} if cfg is None else dict(cfg)
    typ = cfg_.pop('type', None)
    return _CONV[typ](*args, **kwargs, **cfg_)


def build_norm_layer(cfg, num_features, postfix=''):
    cfg_ = dict(cfg)
    typ = cfg_.pop('type')
    abbr, layer_cls = _NORM[typ]
    requires_grad = cfg_.pop('requires_grad', True)
    cfg_.setdefault('eps', 1e-5)
    if typ == 'GN':
        layer = layer_cls(num_channels=num_features, **cfg_)
    else:
        layer = layer_cls(num_features, **cfg_)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return abbr + str(postfix), layer


def build_upsample_layer(cfg, *args, **kwargs):
    raise NotImplementedError('not on the PreWorld forward path')


def build_activation_layer(cfg):
    cfg_ = dict(cfg)
    typ = cfg_.pop('type')
    return getattr(nn, typ)(**cfg_)


class ConvModule(nn.Module):
    """conv -> norm -> act; mmcv 1.6.0 conv_module.py semantics."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1,
                 padding=0, dilation=1, groups=1, bias='auto', conv_cfg=None,
                 norm_cfg=None, act_cfg=dict(type='ReLU'), inplace=True,
                 with_spectral_norm=False, padding_mode='zeros',
                 order=('conv', 'norm', 'act')):
        super().__init__()
        assert order == ('conv', 'norm', 'act') and padding_mode == 'zeros'
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == 'auto':
            bias = not self.with_norm
        self.conv = build_conv_layer(
            conv_cfg, in_channels, out_channels, kernel_size, stride=stride,
            padding=padding, dilation=dilation, groups=groups, bias=bias)
        if self.with_norm:
            self.norm_name, norm = build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        if self.with_activation:
            act_cfg_ = dict(act_cfg)
            if act_cfg_['type'] not in ('Tanh', 'PReLU', 'Sigmoid', 'HSigmoid',
                                        'Swish', 'GELU'):
                act_cfg_.setdefault('inplace', inplace)
            self.activate = build_activation_layer(act_cfg_)

    @property
    def norm(self):
        return getattr(self, self.norm_name) if self.with_norm else None

    def forward(self, x, activate=True, norm=True):
        x = self.conv(x)
        if norm and self.with_norm:
            x = self.norm(x)
        if activate and self.with_activation:
            x = self.activate(x)
        return x


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg

    def init_weights(self):
        pass


def _identity_decorator(*dargs, **dkwargs):
    if len(dargs) == 1 and callable(dargs[0]) and not dkwargs:
        return dargs[0]
    return lambda f: f


# --------------------------------------------------------------------------
# mmdet.models.backbones.resnet
# --------------------------------------------------------------------------
class BasicBlock(BaseModule):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None,
                 style='pytorch', with_cp=False, conv_cfg=None,
                 norm_cfg=dict(type='BN'), dcn=None, plugins=None,
                 init_cfg=None):
        super().__init__(init_cfg)
        _, norm1 = build_norm_layer(norm_cfg, planes, postfix=1)
        _, norm2 = build_norm_layer(norm_cfg, planes, postfix=2)
        self.conv1 = build_conv_layer(conv_cfg, inplanes, planes, 3,
                                      stride=stride, padding=dilation,
                                      dilation=dilation, bias=False)
        self.add_module('bn1', norm1)
        self.conv2 = build_conv_layer(conv_cfg, planes, planes, 3, padding=1,
                                      bias=False)
        self.add_module('bn2', norm2)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        identity = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        if self.downsample is not None:
            identity = self.downsample(x)
        out = out + identity
        return self.relu(out)


class Bottleneck(BaseModule):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None,
                 style='pytorch', with_cp=False, conv_cfg=None,
                 norm_cfg=dict(type='BN'), dcn=None, plugins=None,
                 init_cfg=None):
        super().__init__(init_cfg)
        s1, s2 = (1, stride) if style == 'pytorch' else (stride, 1)
        self.conv1 = build_conv_layer(conv_cfg, inplanes, planes, 1, stride=s1,
                                      bias=False)
        self.add_module('bn1', build_norm_layer(norm_cfg, planes, 1)[1])
        self.conv2 = build_conv_layer(conv_cfg, planes, planes, 3, stride=s2,
                                      padding=dilation, dilation=dilation,
                                      bias=False)
        self.add_module('bn2', build_norm_layer(norm_cfg, planes, 2)[1])
        self.conv3 = build_conv_layer(conv_cfg, planes, planes * 4, 1,
                                      bias=False)
        self.add_module('bn3', build_norm_layer(norm_cfg, planes * 4, 3)[1])
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        identity = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        if self.downsample is not None:
            identity = self.downsample(x)
        out = out + identity
        return self.relu(out)


class ResNet(BaseModule):
    """mmdet 2.24 ResNet (style='pytorch', no deep stem / DCN / plugins)."""
    arch_settings = {18: (BasicBlock, (2, 2, 2, 2)),
                     34: (BasicBlock, (3, 4, 6, 3)),
                     50: (Bottleneck, (3, 4, 6, 3)),
                     101: (Bottleneck, (3, 4, 23, 3)),
                     152: (Bottleneck, (3, 8, 36, 3))}

    def __init__(self, depth, in_channels=3, stem_channels=None,
                 base_channels=64, num_stages=4, strides=(1, 2, 2, 2),
                 dilations=(1, 1, 1, 1), out_indices=(0, 1, 2, 3),
                 style='pytorch', deep_stem=False, avg_down=False,
                 frozen_stages=-1, conv_cfg=None,
                 norm_cfg=dict(type='BN', requires_grad=True), norm_eval=True,
                 dcn=None, stage_with_dcn=(False, False, False, False),
                 plugins=None, with_cp=False, zero_init_residual=True,
                 pretrained=None, init_cfg=None):
        super().__init__(init_cfg)
        assert not deep_stem and not avg_down and dcn is None
        block, stage_blocks = self.arch_settings[depth]
        self.deep_stem = deep_stem
        self.out_indices = out_indices
        stem_channels = stem_channels or base_channels
        self.conv1 = nn.Conv2d(in_channels, stem_channels, 7, 2, 3, bias=False)
        self.norm1_name, norm1 = build_norm_layer(norm_cfg, stem_channels, 1)
        self.add_module(self.norm1_name, norm1)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.res_layers = []
        inplanes = stem_channels
        for i, nb in enumerate(stage_blocks[:num_stages]):
            planes = base_channels * 2 ** i
            layers = []
            for b in range(nb):
                stride = strides[i] if b == 0 else 1
                downsample = None
                if b == 0 and (stride != 1
                               or inplanes != planes * block.expansion):
                    downsample = nn.Sequential(
                        nn.Conv2d(inplanes, planes * block.expansion, 1,
                                  stride=stride, bias=False),
                        build_norm_layer(norm_cfg,
                                         planes * block.expansion)[1])
                layers.append(block(inplanes, planes, stride=stride,
                                    dilation=dilations[i],
                                    downsample=downsample, style=style,
                                    norm_cfg=norm_cfg))
                inplanes = planes * block.expansion
            name = f'layer{i + 1}'
            self.add_module(name, nn.Sequential(*layers))
            self.res_layers.append(name)

    @property
    def norm1(self):
        return getattr(self, self.norm1_name)

    def forward(self, x):
        x = self.relu(self.norm1(self.conv1(x)))
        x = self.maxpool(x)
        outs = []
        for i, name in enumerate(self.res_layers):
            x = getattr(self, name)(x)
            if i in self.out_indices:
                outs.append(x)
        return tuple(outs)


# --------------------------------------------------------------------------
# CPU restatements of the reference's CUDA-only ops
# --------------------------------------------------------------------------
def bev_pool_v2_cpu(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                    bev_feat_shape, interval_starts, interval_lengths):
    """Follows mmdet3d/ops/bev_pool_v2/bev_pool.py:17-41,86-92 and the kernel
    mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:21-48 via the C oracle."""
    from oracle import c_ref
    depth = depth.contiguous().float()
    feat = feat.contiguous().float()
    out = c_ref.bev_pool_v2_fwd(
        depth.numpy().ravel(), feat.numpy().reshape(-1, feat.shape[-1]),
        ranks_depth.int().numpy(), ranks_feat.int().numpy(),
        ranks_bev.int().numpy(), interval_starts.int().numpy(),
        interval_lengths.int().numpy(), int(np.prod(bev_feat_shape[:-1])))
    out = torch.from_numpy(out).view(*bev_feat_shape)
    return out.permute(0, 4, 1, 2, 3).contiguous()


class _Raw2Alpha:
    @staticmethod
    def apply(density, shift, interval):
        from oracle import c_ref
        a = c_ref.raw2alpha(density.detach().float().numpy().ravel(),
                            float(shift), float(interval))
        return torch.from_numpy(a).view(density.shape)


class _Alphas2Weights:
    @staticmethod
    def apply(alpha, ray_id, n_rays):
        from oracle import c_ref
        w, last = c_ref.alpha2weight(alpha.detach().float().numpy(),
                                     ray_id.long().numpy(), int(n_rays))
        return torch.from_numpy(w), torch.from_numpy(last)


class _UB360:
    @staticmethod
    def cumdist_thres(dist, thres):
        from oracle import c_ref
        m = c_ref.cumdist_thres(dist.detach().float().numpy(), float(thres))
        return torch.from_numpy(m)


def segment_coo(src, index, out=None, reduce='sum'):
    """torch_scatter.segment_coo(reduce='sum') on a sorted index ==
    index_add_ (reference call sites nerf_head.py:332-353)."""
    assert reduce == 'sum'
    return out.index_add_(0, index.to(out.device).long(), src)


def _install(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []  # behave like a package
        sys.modules[name] = m
        if '.' in name:
            parent, child = name.rsplit('.', 1)
            setattr(_install(parent), child, m)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


_installed = False


def install():
    """Install the stand-in packages and stub the reference's package
    ``__init__``s (they import compiled mmcv ops) so that single reference
    files can be imported verbatim by dotted name."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')
    _installed = True

    MMCV_MODELS = Registry('model')
    _install('mmcv', __version__='1.6.0')
    _install('mmcv.utils', Registry=Registry)
    _install('mmcv.cnn', MODELS=MMCV_MODELS, ConvModule=ConvModule,
             build_conv_layer=build_conv_layer,
             build_norm_layer=build_norm_layer,
             build_upsample_layer=build_upsample_layer)
    _install('mmcv.cnn.bricks')
    _install('mmcv.cnn.bricks.conv_module', ConvModule=ConvModule)
    _install('mmcv.runner', BaseModule=BaseModule,
             force_fp32=_identity_decorator, auto_fp16=_identity_decorator)

    MMDET = Registry('models', parent=MMCV_MODELS)
    MMDET.register_module(module=ResNet)

    class _NullLoss(nn.Module):
        def __init__(self, **kw):
            super().__init__()

    for n in ('CrossEntropyLoss', 'CustomFocalLoss'):
        MMDET.register_module(name=n, module=_NullLoss)
    _install('mmdet')
    _install('mmdet.core', reduce_mean=lambda t: t)
    _install('mmdet.models', BACKBONES=MMDET, NECKS=MMDET, HEADS=MMDET,
             DETECTORS=MMDET, LOSSES=MMDET)
    _install('mmdet.models.builder', BACKBONES=MMDET, NECKS=MMDET, HEADS=MMDET,
             DETECTORS=MMDET, LOSSES=MMDET, ROI_EXTRACTORS=MMDET,
             SHARED_HEADS=MMDET, build_loss=lambda cfg: MMDET.build(cfg))
    _install('mmdet.models.backbones')
    _install('mmdet.models.backbones.resnet', BasicBlock=BasicBlock,
             Bottleneck=Bottleneck, ResNet=ResNet)
    _install('mmseg')
    _install('mmseg.models')
    _install('mmseg.models.builder', LOSSES=MMDET)

    _install('torch_scatter', segment_coo=segment_coo)
    _install('torch_efficient_distloss', flatten_eff_distloss=None)
    _install('IPython', embed=lambda *a, **k: None)
    _install('matplotlib', cm=None)
    _install('matplotlib.pyplot')

    # the reference package tree, with empty package __init__s
    ref = os.path.join(REFERENCE_ROOT, 'mmdet3d')
    for pkg in ('mmdet3d', 'mmdet3d.models', 'mmdet3d.models.necks',
                'mmdet3d.models.backbones', 'mmdet3d.models.detectors',
                'mmdet3d.models.heads', 'mmdet3d.models.nerf', 'mmdet3d.ops',
                'mmdet3d.ops.bev_pool_v2', 'mmdet3d.core'):
        m = _install(pkg)
        m.__path__ = [os.path.join(ref, *pkg.split('.')[1:])]
    _install('mmdet3d.ops.bev_pool_v2.bev_pool', bev_pool_v2=bev_pool_v2_cpu,
             TRTBEVPoolv2=None)
    _install('mmdet3d.core.bbox', Box3DMode=types.SimpleNamespace(LIDAR=0),
             Coord3DMode=None, LiDARInstance3DBoxes=None)

    class SwinTransformer(nn.Module):  # isinstance() target only
        pass
    _install('mmdet3d.models.backbones.swin', SwinTransformer=SwinTransformer)

    # nerf/utils.py JIT-compiles CUDA at import (utils.py:12-24) and imports
    # tkinter; its three autograd wrappers + two losses are restated instead.
    class silog_loss(nn.Module):
        def __init__(self, variance_focus=0.85):
            super().__init__()
            self.variance_focus = variance_focus

    class l1_loss(nn.Module):
        pass
    _install('mmdet3d.models.nerf.utils', Raw2Alpha=_Raw2Alpha,
             Alphas2Weights=_Alphas2Weights, ub360_utils_cuda=_UB360,
             silog_loss=silog_loss, l1_loss=l1_loss, cos_sim_loss=l1_loss)

    # real reference builder (mmdet3d/models/builder.py) on top of the shims
    builder = importlib.import_module('mmdet3d.models.builder')
    sys.modules['mmdet3d.models'].builder = builder

    # Base3DDetector/MVXTwoStageDetector/CenterPoint pull in the LiDAR stack
    # (mmcv.ops Voxelization, bbox coders ...).  Only three behaviours of that
    # chain are on the camera path and they are restated here:
    #   mvx_two_stage.py:65-68  build img_backbone / img_neck
    #   mvx_two_stage.py 'with_img_neck' property
    #   base.py:47-62           forward(return_loss) dispatch
    class CenterPoint(BaseModule):
        def __init__(self, img_backbone=None, img_neck=None,
                     pts_bbox_head=None, train_cfg=None, test_cfg=None,
                     pretrained=None, init_cfg=None, **kw):
            super().__init__(init_cfg)
            if img_backbone:
                self.img_backbone = builder.build_backbone(img_backbone)
            if img_neck is not None:
                self.img_neck = builder.build_neck(img_neck)

        @property
        def with_img_neck(self):
            return hasattr(self, 'img_neck') and self.img_neck is not None

        def forward(self, return_loss=True, **kwargs):
            if return_loss:
                return self.forward_train(**kwargs)
            return self.forward_test(**kwargs)
    _install('mmdet3d.models.detectors.centerpoint', CenterPoint=CenterPoint)

    heads = importlib.import_module('mmdet3d.models.heads.occupancy_head')
    sys.modules['mmdet3d.models.heads'].DownScaleModule3DCustom = \
        heads.DownScaleModule3DCustom
    sys.modules['mmdet3d.models.heads'].OccHead = heads.OccHead


def load(dotted):
    """Import a reference module verbatim, e.g.
    ``load('mmdet3d.models.necks.view_transformer')``."""
    install()
    return importlib.import_module(dotted)


def load_all():
    """Import every reference file on the hot path (registers the classes in
    the reference's own MODELS registry) and return that registry's builder."""
    install()
    for mod in ('necks.view_transformer', 'necks.fpn', 'necks.lss_fpn',
                'backbones.resnet', 'heads.occupancy_head', 'nerf.nerf_head',
                'detectors.bevdet', 'detectors.bevdet_occ',
                'detectors.preworld', 'detectors.preworld_temporal_traj'):
        importlib.import_module('mmdet3d.models.' + mod)
    return sys.modules['mmdet3d.models.builder']
