"""Registry surface of the reference (mmdet3d/models/builder.py:16-28,81-122).

One ``MODELS`` registry aliased as BACKBONES / NECKS / HEADS / LOSSES /
DETECTORS; classes self-register with ``@NECKS.register_module()`` and config
dicts select them by their ``type`` string, exactly as
``configs/preworld/**`` expect.  mmcv is not a dependency: this is the ~40
lines of ``mmcv.utils.Registry`` the path uses.
"""
import torch.nn as nn


class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key):
        return self._module_dict.get(key)

    def __contains__(self, key):
        return key in self._module_dict

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            key = name or cls.__name__
            if key in self._module_dict and not force:
                raise KeyError(f'{key} is already registered in {self._name}')
            self._module_dict[key] = cls
            return cls
        return _register(module) if module is not None else _register

    def build(self, cfg, default_args=None):
        if not isinstance(cfg, dict) or 'type' not in cfg:
            raise TypeError(f'cfg must be a dict with a "type" key, got {cfg}')
        args = dict(cfg)
        for k, v in (default_args or {}).items():
            args.setdefault(k, v)
        typ = args.pop('type')
        cls = self.get(typ) if isinstance(typ, str) else typ
        if cls is None:
            raise KeyError(f'{typ} is not in the {self._name} registry')
        return cls(**args)


MODELS = Registry('models')
BACKBONES = MODELS
NECKS = MODELS
HEADS = MODELS
LOSSES = MODELS
DETECTORS = MODELS


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_neck(cfg):
    return NECKS.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)


def build_loss(cfg):
    return LOSSES.build(cfg)


def build_detector(cfg, train_cfg=None, test_cfg=None):
    return DETECTORS.build(
        cfg, default_args=dict(train_cfg=train_cfg, test_cfg=test_cfg))


def build_model(cfg, train_cfg=None, test_cfg=None):
    return build_detector(cfg, train_cfg=train_cfg, test_cfg=test_cfg)


class _TrainingOnlyLoss(nn.Module):
    """Placeholder for the training loss ``loss_occ`` the base config names
    (bevstereo-occ.py:109-112, unused by PreWorld's own heads); calling it
    raises.  ``CustomFocalLoss`` (preworld.py:117) is the real thing below."""

    def __init__(self, **kwargs):
        super().__init__()
        self.cfg = kwargs

    def forward(self, *a, **k):
        raise NotImplementedError(
            'training losses are out of scope of the forward-only path')


for _n in ('CrossEntropyLoss',):
    LOSSES.register_module(name=_n, module=type(_n, (_TrainingOnlyLoss,), {}))

# preworld.py:117 builds this one; it is implemented
from ..losses import CustomFocalLoss as _CustomFocalLoss  # noqa: E402
LOSSES.register_module(name='CustomFocalLoss', module=_CustomFocalLoss)
