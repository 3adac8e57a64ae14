"""Minimal loader for mmcv-style python config files.

The reference selects every module of the hot path through config dicts
(`configs/preworld/nuscenes/*.py`, loaded by `mmcv.Config.fromfile` in
reference tools/train.py:125 / tools/test.py:162).  This loader implements the
subset those files use: plain-python evaluation, `_base_` inheritance
(string or list, relative paths), recursive dict merge with `_delete_`, and
attribute access.  It lets the reference's config files -- and the derived R50
configs under ./configs -- be loaded unchanged.
"""
import copy
import os
import types


class ConfigDict(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value


def _to_config_dict(obj):
    if isinstance(obj, dict):
        return ConfigDict({k: _to_config_dict(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_config_dict(v) for v in obj)
    return obj


def _merge(base, child):
    """mmcv Config._merge_a_into_b: child keys override; dicts merge
    recursively unless the child carries `_delete_=True`."""
    out = copy.deepcopy(base)
    for k, v in child.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            if v.get('_delete_', False):
                out[k] = {kk: vv for kk, vv in v.items() if kk != '_delete_'}
            else:
                out[k] = _merge(out[k], v)
        else:
            out[k] = copy.deepcopy(v)
    return out


def _load_file(path):
    path = os.path.abspath(path)
    scope = {'__file__': path}
    with open(path) as f:
        exec(compile(f.read(), path, 'exec'), scope)
    cfg = {k: v for k, v in scope.items()
           if not k.startswith('__') and not isinstance(
               v, (types.ModuleType, types.FunctionType, type))}
    bases = cfg.pop('_base_', [])
    if isinstance(bases, str):
        bases = [bases]
    merged = {}
    for b in bases:
        merged = _merge(merged, _load_file(
            os.path.join(os.path.dirname(path), b)))
    return _merge(merged, cfg)


class Config:
    @staticmethod
    def fromfile(path):
        return _to_config_dict(_load_file(path))
