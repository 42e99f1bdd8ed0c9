"""Minimal `Config.fromfile` / `ConfigDict` / `Registry` / `build_from_cfg`.

The reference's plugin API is mmcv's string->class registry driven by python-dict config files
(`mmdet/models/builder.py:6-34`, `tools/train.py:89`).  mmcv is a third-party dependency that is
neither vendored in the reference nor installed on the B200 box, so the host side restates the
published semantics the LD configs rely on:

* config files are executed as python; `_base_` (str or list) files are loaded first and the child's
  dicts are merged key-by-key on top, `_delete_=True` replaces instead of merging;
* `ConfigDict` gives attribute access (`train_cfg.assigner`, `gfl_head.py:94`);
* `build_from_cfg(cfg, registry, default_args)` pops `type`, looks the class up by `__name__`,
  passes the remaining keys as kwargs; unknown type -> KeyError, unknown kwarg -> TypeError.
"""
import copy
import inspect
import os

DELETE_KEY = '_delete_'
BASE_KEY = '_base_'


class ConfigDict(dict):
    def __init__(self, *a, **kw):
        super().__init__()
        for k, v in dict(*a, **kw).items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(f"'ConfigDict' object has no attribute '{k}'")

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def to_dict(self):
        def conv(v):
            if isinstance(v, dict):
                return {k: conv(x) for k, x in v.items()}
            if isinstance(v, (list, tuple)):
                return type(v)(conv(x) for x in v)
            return v
        return conv(self)


def _merge_a_into_b(a, b):
    b = dict(b)
    for k, v in a.items():
        if isinstance(v, dict) and k in b and not v.get(DELETE_KEY, False):
            if not isinstance(b[k], dict):
                raise TypeError(f'{k}={v} in child config cannot inherit from base because {k} '
                                f'is a dict in the child config but is of type {type(b[k])} in base; '
                                f'set `{DELETE_KEY}=True` to ignore the base config')
            b[k] = _merge_a_into_b(v, b[k])
        else:
            if isinstance(v, dict):
                v = {kk: vv for kk, vv in v.items() if kk != DELETE_KEY}
            b[k] = v
    return b


def _file2dict(filename):
    filename = os.path.abspath(os.path.expanduser(filename))
    if not os.path.isfile(filename):
        raise FileNotFoundError(f'config file {filename} does not exist')
    if not filename.endswith('.py'):
        raise IOError('Only py type configs are supported')
    ns = {'__file__': filename}
    with open(filename) as f:
        exec(compile(f.read(), filename, 'exec'), ns)
    cfg = {k: v for k, v in ns.items()
           if not k.startswith('__') and not inspect.ismodule(v) and not inspect.isfunction(v)
           and not inspect.isclass(v)}
    if BASE_KEY in cfg:
        base = cfg.pop(BASE_KEY)
        base = base if isinstance(base, list) else [base]
        merged = {}
        for b in base:
            bd = _file2dict(os.path.join(os.path.dirname(filename), b))
            dup = merged.keys() & bd.keys()
            if dup:
                raise KeyError(f'Duplicate key is not allowed among bases: {sorted(dup)}')
            merged.update(bd)
        cfg = _merge_a_into_b(cfg, merged)
    return cfg


class Config:
    """Attribute-access view of a python config file with `_base_` inheritance."""

    def __init__(self, cfg_dict=None, filename=None):
        object.__setattr__(self, '_cfg_dict', ConfigDict(cfg_dict or {}))
        object.__setattr__(self, '_filename', filename)

    @staticmethod
    def fromfile(filename):
        return Config(_file2dict(filename), filename=filename)

    @property
    def filename(self):
        return self._filename

    def __getattr__(self, k):
        return getattr(self._cfg_dict, k)

    def __getitem__(self, k):
        return self._cfg_dict[k]

    def __contains__(self, k):
        return k in self._cfg_dict

    def get(self, k, default=None):
        return self._cfg_dict.get(k, default)

    def __setattr__(self, k, v):
        self._cfg_dict[k] = v

    def merge_from_dict(self, options):
        """`--cfg-options a.b.c=v` semantics (tools/train.py:55-64)."""
        nested = {}
        for full, v in options.items():
            d = nested
            keys = full.split('.')
            for kk in keys[:-1]:
                d = d.setdefault(kk, {})
            d[keys[-1]] = v
        object.__setattr__(self, '_cfg_dict',
                           ConfigDict(_merge_a_into_b(nested, self._cfg_dict.to_dict())))


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

    def __len__(self):
        return len(self._module_dict)

    def __repr__(self):
        return f'Registry(name={self._name}, items={sorted(self._module_dict)})'

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            key = name or cls.__name__
            if not force and key in self._module_dict:
                raise KeyError(f'{key} is already registered in {self._name}')
            self._module_dict[key] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict):
        raise TypeError(f'cfg must be a dict, but got {type(cfg)}')
    if 'type' not in cfg and not (default_args and 'type' in default_args):
        raise KeyError(f'`cfg` or `default_args` must contain the key "type", but got {cfg}')
    args = dict(cfg)
    if default_args is not None:
        for k, v in default_args.items():
            args.setdefault(k, v)
    obj_type = args.pop('type')
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError(f'{obj_type} is not in the {registry.name} registry')
    elif inspect.isclass(obj_type):
        obj_cls = obj_type
    else:
        raise TypeError(f'type must be a str or valid type, but got {type(obj_type)}')
    return obj_cls(**args)
