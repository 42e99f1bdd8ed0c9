"""Minimal stand-in for `mmcv` 1.2.x so the UNMODIFIED reference (`/root/reference/mmdet`)
can be imported in the build container and run on CPU.

TEST INFRASTRUCTURE ONLY.  Nothing in `ld_b200/` may import this file.  It is used by
`oracle/make_golden.py` (golden-vector generation, build container only: `/root/reference`
does not exist on the GPU box) and by the `-m "not gpu"` tests that cross-check the oracle
restatement against the reference when `/root/reference` is present.

mmcv is a third-party dependency of the reference (`mmdet/__init__.py:19-20` pins
mmcv-full >=1.2.4,<=1.3) whose source is NOT under `/root/reference`.  What is restated here
is the published behaviour of the handful of mmcv symbols the LD train path touches
(SURVEY.md Appendix C): Registry/build_from_cfg, ConvModule, Scale, build_conv_layer,
build_norm_layer, the nn.init helpers, and identity decorators for force_fp32/auto_fp16/jit.
Everything else (`mmcv.ops`, `mmcv.parallel`, pycocotools, terminaltables, matplotlib ...)
is a permissive dummy that only has to survive `import mmdet`.
"""
import inspect
import math
import sys
import types

import numpy as np
import torch
import torch.nn as nn

_INSTALLED = False


# --------------------------------------------------------------------------- registry
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
        return self._module_dict.get(key, None)

    def __contains__(self, key):
        return key in self._module_dict

    def __len__(self):
        return len(self._module_dict)

    def _register(self, cls, name=None, force=False):
        name = name or cls.__name__
        if not force and name in self._module_dict:
            raise KeyError(f'{name} is already registered in {self._name}')
        self._module_dict[name] = cls

    def register_module(self, name=None, force=False, module=None):
        if isinstance(name, type):  # legacy @R.register_module usage
            self._register(name)
            return name
        if module is not None:
            self._register(module, name, force)
            return module

        def _deco(cls):
            self._register(cls, name, force)
            return cls

        return _deco


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict):
        raise TypeError(f'cfg must be a dict, but got {type(cfg)}')
    if 'type' not in cfg and not (default_args and 'type' in default_args):
        raise KeyError('`cfg` or `default_args` must contain the key "type"')
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
        raise TypeError(f'type must be a str or class, got {type(obj_type)}')
    return obj_cls(**args)


class ConfigDict(dict):
    """dict with attribute access (addict-like), nested dicts converted on the way in."""

    def __init__(self, *a, **kw):
        super().__init__()
        for k, v in dict(*a, **kw).items():
            self[k] = v

    @classmethod
    def _conv(cls, v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._conv(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._conv(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


# --------------------------------------------------------------------------- init helpers
def constant_init(module, val, bias=0):
    if hasattr(module, 'weight') and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def xavier_init(module, gain=1, bias=0, distribution='normal'):
    if distribution == 'uniform':
        nn.init.xavier_uniform_(module.weight, gain=gain)
    else:
        nn.init.xavier_normal_(module.weight, gain=gain)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def normal_init(module, mean=0, std=1, bias=0):
    nn.init.normal_(module.weight, mean, std)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def uniform_init(module, a=0, b=1, bias=0):
    nn.init.uniform_(module.weight, a, b)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def kaiming_init(module, a=0, mode='fan_out', nonlinearity='relu', bias=0,
                 distribution='normal'):
    if distribution == 'uniform':
        nn.init.kaiming_uniform_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    else:
        nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def caffe2_xavier_init(module, bias=0):
    kaiming_init(module, a=1, mode='fan_in', nonlinearity='leaky_relu', distribution='uniform')


def bias_init_with_prob(prior_prob):
    return float(-np.log((1 - prior_prob) / prior_prob))


# --------------------------------------------------------------------------- layers
class DeformConv2dPack(nn.Module):
    """mmcv.ops.DeformConv2dPack (conv type 'DCN', mmcv 1.2.x deform_conv.py) restated on torchvision.ops.deform_conv2d,
    which shares mmcv's offset layout (channel 2k = dy, 2k+1 = dx of kernel point k, per deformable group) and sampling
    rule (zero outside (-1, H) x (-1, W)).  Parameters: `weight`, `conv_offset.{weight,bias}` (zero-initialised).  mmcv's
    own kernel is not available here: parity for DCN is pinned to this restatement only."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deform_groups=1, bias=False, **kw):
        super().__init__()
        assert not bias
        import math
        self.stride, self.padding, self.dilation, self.groups, self.deform_groups = stride, padding, dilation, groups, \
            deform_groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, kernel_size, kernel_size))
        n = in_channels * kernel_size * kernel_size
        nn.init.uniform_(self.weight, -1.0 / math.sqrt(n), 1.0 / math.sqrt(n))
        self.conv_offset = nn.Conv2d(in_channels, deform_groups * 2 * kernel_size * kernel_size, kernel_size, stride,
                                     padding, dilation=dilation, bias=True)
        nn.init.zeros_(self.conv_offset.weight)
        nn.init.zeros_(self.conv_offset.bias)

    def forward(self, x):
        from torchvision.ops import deform_conv2d
        return deform_conv2d(x, self.conv_offset(x), self.weight, None, self.stride, self.padding, self.dilation)


def build_conv_layer(cfg, *args, **kwargs):
    if cfg is None or cfg.get('type', 'Conv2d') in ('Conv2d', 'Conv'):
        return nn.Conv2d(*args, **kwargs)
    if cfg.get('type') == 'DCN':
        extra = {k: v for k, v in cfg.items() if k != 'type'}
        return DeformConv2dPack(*args, **kwargs, **extra)
    raise NotImplementedError(f'conv type {cfg} not available in the mmcv stand-in')


def build_norm_layer(cfg, num_features, postfix=''):
    cfg = dict(cfg)
    t = cfg.pop('type')
    requires_grad = cfg.pop('requires_grad', True)
    cfg.setdefault('eps', 1e-5)
    if t == 'BN':
        name, layer = 'bn', nn.BatchNorm2d(num_features, **cfg)
    elif t == 'GN':
        name, layer = 'gn', nn.GroupNorm(num_channels=num_features, **cfg)
    else:
        raise NotImplementedError(t)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return name + str(postfix), layer


class Scale(nn.Module):
    def __init__(self, scale=1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(scale, dtype=torch.float))

    def forward(self, x):
        return x * self.scale


class ConvModule(nn.Module):
    """conv -> norm -> act, bias='auto' => bias iff no norm (mmcv/cnn/bricks/conv_module.py)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0,
                 dilation=1, groups=1, bias='auto', conv_cfg=None, norm_cfg=None,
                 act_cfg=dict(type='ReLU'), inplace=True, with_spectral_norm=False,
                 padding_mode='zeros', order=('conv', 'norm', 'act')):
        super().__init__()
        assert order == ('conv', 'norm', 'act') and padding_mode == 'zeros'
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == 'auto':
            bias = not self.with_norm
        self.with_bias = bias
        self.conv = build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size,
                                     stride=stride, padding=padding, dilation=dilation,
                                     groups=groups, bias=bias)
        self.in_channels, self.out_channels = in_channels, out_channels
        if self.with_norm:
            self.norm_name, norm = build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        if self.with_activation:
            assert act_cfg['type'] == 'ReLU'
            self.activate = nn.ReLU(inplace=inplace)
        self.init_weights()

    @property
    def norm(self):
        return getattr(self, self.norm_name)

    def init_weights(self):
        kaiming_init(self.conv, a=0, nonlinearity='relu')
        if self.with_norm:
            constant_init(self.norm, 1, bias=0)

    def forward(self, x, activate=True, norm=True):
        x = self.conv(x)
        if norm and self.with_norm:
            x = self.norm(x)
        if activate and self.with_activation:
            x = self.activate(x)
        return x


# --------------------------------------------------------------------------- decorators
def _identity_decorator_factory(*dargs, **dkwargs):
    if len(dargs) == 1 and callable(dargs[0]) and not dkwargs:
        return dargs[0]

    def _wrap(f):
        return f

    return _wrap


def load_checkpoint(model, filename, map_location=None, strict=False, logger=None):
    ckpt = torch.load(filename, map_location=map_location)
    sd = ckpt.get('state_dict', ckpt)
    model.load_state_dict(sd, strict=strict)
    return ckpt


def get_dist_info():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def is_tuple_of(seq, expected_type):
    return isinstance(seq, tuple) and all(isinstance(x, expected_type) for x in seq)


def is_list_of(seq, expected_type):
    return isinstance(seq, list) and all(isinstance(x, expected_type) for x in seq)


def is_seq_of(seq, expected_type, seq_type=None):
    return isinstance(seq, (list, tuple)) and all(isinstance(x, expected_type) for x in seq)


def is_str(x):
    return isinstance(x, str)


# --------------------------------------------------------------------------- permissive stubs
class _Dummy:
    """Subclassable / callable placeholder for symbols that only need to exist."""

    def __init__(self, *a, **kw):
        pass

    def __call__(self, *a, **kw):
        if len(a) == 1 and callable(a[0]) and not kw:  # used as a decorator
            return a[0]
        raise NotImplementedError('mmcv stand-in: symbol is a placeholder')

    @classmethod
    def register_module(cls, *a, **kw):
        return _identity_decorator_factory(*a, **kw)


class _PermissiveModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        val = type(name, (_Dummy,), {})
        setattr(self, name, val)
        return val


# --------------------------------------------------------------------------- mmcv.image (cv2 / numpy)
def rescale_size(old_size, scale, return_scale=False):
    """mmcv/image/geometric.py rescale_size."""
    w, h = old_size
    if isinstance(scale, (float, int)):
        scale_factor = scale
    else:
        max_long_edge, max_short_edge = max(scale), min(scale)
        scale_factor = min(max_long_edge / max(h, w), max_short_edge / min(h, w))
    new_size = (int(w * float(scale_factor) + 0.5), int(h * float(scale_factor) + 0.5))
    return (new_size, scale_factor) if return_scale else new_size


def imresize(img, size, return_scale=False, interpolation='bilinear', out=None, backend=None):
    import cv2
    h, w = img.shape[:2]
    codes = dict(nearest=cv2.INTER_NEAREST, bilinear=cv2.INTER_LINEAR, bicubic=cv2.INTER_CUBIC, area=cv2.INTER_AREA)
    resized = cv2.resize(img, size, dst=out, interpolation=codes[interpolation])
    if not return_scale:
        return resized
    return resized, size[0] / w, size[1] / h


def imrescale(img, scale, return_scale=False, interpolation='bilinear', backend=None):
    h, w = img.shape[:2]
    new_size, scale_factor = rescale_size((w, h), scale, return_scale=True)
    rescaled = imresize(img, new_size, interpolation=interpolation)
    return (rescaled, scale_factor) if return_scale else rescaled


def imflip(img, direction='horizontal'):
    assert direction in ('horizontal', 'vertical', 'diagonal')
    return np.flip(img, axis=1) if direction == 'horizontal' else (
        np.flip(img, axis=0) if direction == 'vertical' else np.flip(img, axis=(0, 1)))


def imnormalize(img, mean, std, to_rgb=True):
    import cv2
    img = img.copy().astype(np.float32)
    assert img.dtype != np.uint8
    mean = np.float64(mean.reshape(1, -1))
    stdinv = 1 / np.float64(std.reshape(1, -1))
    if to_rgb:
        cv2.cvtColor(img, cv2.COLOR_BGR2RGB, img)
    cv2.subtract(img, mean, img)
    cv2.multiply(img, stdinv, img)
    return img


def impad(img, *, shape=None, padding=None, pad_val=0, padding_mode='constant'):
    import cv2
    assert (shape is not None) ^ (padding is not None)
    if shape is not None:
        padding = (0, 0, shape[1] - img.shape[1], shape[0] - img.shape[0])
    return cv2.copyMakeBorder(img, padding[1], padding[3], padding[0], padding[2], cv2.BORDER_CONSTANT, value=pad_val)


def impad_to_multiple(img, divisor, pad_val=0):
    pad_h = int(np.ceil(img.shape[0] / divisor)) * divisor
    pad_w = int(np.ceil(img.shape[1] / divisor)) * divisor
    return impad(img, shape=(pad_h, pad_w), pad_val=pad_val)


def batched_nms(boxes, scores, idxs, nms_cfg, class_agnostic=False):
    """mmcv.ops.batched_nms (mmcv 1.2.x ops/nms.py:229-292) restated: per-class NMS through the coordinate-offset trick
    (boxes + class * (max coordinate + 1)), greedy IoU > threshold suppression (torchvision.ops.nms has the same rule),
    results sorted by score.  Above `split_thr` boxes mmcv loops over the classes instead; both give the same keep set."""
    from torchvision.ops import nms as tv_nms
    nms_cfg_ = dict(nms_cfg)
    class_agnostic = nms_cfg_.pop('class_agnostic', class_agnostic)
    assert nms_cfg_.pop('type', 'nms') == 'nms'
    split_thr = nms_cfg_.pop('split_thr', 10000)
    thr = nms_cfg_['iou_threshold']
    if class_agnostic:
        boxes_for_nms = boxes
    else:
        max_coordinate = boxes.max()
        offsets = idxs.to(boxes) * (max_coordinate + torch.tensor(1).to(boxes))
        boxes_for_nms = boxes + offsets[:, None]
    if len(boxes_for_nms) < split_thr:
        keep = tv_nms(boxes_for_nms, scores, thr)
        boxes, scores = boxes[keep], scores[keep]
    else:
        total_mask = scores.new_zeros(scores.size(), dtype=torch.bool)
        for cid in torch.unique(idxs):
            mask = (idxs == cid).nonzero(as_tuple=False).view(-1)
            total_mask[mask[tv_nms(boxes_for_nms[mask], scores[mask], thr)]] = True
        keep = total_mask.nonzero(as_tuple=False).view(-1)
        keep = keep[scores[keep].argsort(descending=True)]
        boxes, scores = boxes[keep], scores[keep]
    return torch.cat([boxes, scores[:, None]], -1), keep


def _mod(name, permissive=True, **attrs):
    m = (_PermissiveModule if permissive else types.ModuleType)(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so `import a.b` works
    sys.modules[name] = m
    return m


def install():
    """Inject the stand-in modules into sys.modules (idempotent)."""
    global _INSTALLED
    if _INSTALLED:
        return
    _INSTALLED = True
    hooks_registry = Registry('hook')

    mmcv = _mod('mmcv', __version__='1.2.7', jit=_identity_decorator_factory,
                ConfigDict=ConfigDict, is_tuple_of=is_tuple_of, is_list_of=is_list_of,
                is_seq_of=is_seq_of, is_str=is_str, imrescale=imrescale, imresize=imresize, imflip=imflip,
                imnormalize=imnormalize, impad=impad, impad_to_multiple=impad_to_multiple, rescale_size=rescale_size)
    mmcv.utils = _mod('mmcv.utils', Registry=Registry, build_from_cfg=build_from_cfg,
                      print_log=lambda *a, **k: None,
                      get_logger=lambda name, **k: __import__('logging').getLogger(name),
                      get_git_hash=lambda *a, **k: 'standin', is_tuple_of=is_tuple_of)
    mmcv.cnn = _mod('mmcv.cnn', ConvModule=ConvModule, Scale=Scale,
                    build_conv_layer=build_conv_layer, build_norm_layer=build_norm_layer,
                    normal_init=normal_init, constant_init=constant_init,
                    xavier_init=xavier_init, kaiming_init=kaiming_init,
                    uniform_init=uniform_init, caffe2_xavier_init=caffe2_xavier_init,
                    bias_init_with_prob=bias_init_with_prob,
                    is_norm=lambda m: isinstance(m, (nn.modules.batchnorm._BatchNorm,
                                                     nn.GroupNorm)))
    mmcv.cnn.bricks = _mod('mmcv.cnn.bricks')
    _mod('mmcv.cnn.bricks.transformer')
    mmcv.runner = _mod('mmcv.runner', force_fp32=_identity_decorator_factory,
                       auto_fp16=_identity_decorator_factory,
                       load_checkpoint=load_checkpoint, get_dist_info=get_dist_info,
                       HOOKS=hooks_registry)
    mmcv.runner.hooks = _mod('mmcv.runner.hooks', HOOKS=hooks_registry)
    mmcv.ops = _mod('mmcv.ops')
    for sub in ('nms', 'roi_align', 'carafe', 'merge_cells'):
        setattr(mmcv.ops, sub, _mod('mmcv.ops.' + sub))
    mmcv.ops.nms.batched_nms = batched_nms
    mmcv.ops.batched_nms = batched_nms
    mmcv.parallel = _mod('mmcv.parallel')
    mmcv.image = _mod('mmcv.image')

    pc = _mod('pycocotools', __version__='12.0.2')      # mmdet/datasets/coco.py:21 asserts on it at import
    for sub in ('mask', 'coco', 'cocoeval'):
        setattr(pc, sub, _mod('pycocotools.' + sub))
    _mod('terminaltables')
    mpl = _mod('matplotlib')
    for sub in ('pyplot', 'collections', 'patches'):
        setattr(mpl, sub, _mod('matplotlib.' + sub))
    # kd_one_stage.py:2 has a stray `from numpy.lib.twodim_base import tri` (gone in NumPy 2)
    if 'numpy.lib.twodim_base' not in sys.modules:
        try:
            import numpy.lib.twodim_base  # noqa: F401
        except Exception:
            _mod('numpy.lib.twodim_base', permissive=False, tri=np.tri)


def import_reference(ref_root='/root/reference'):
    """Return the reference's `mmdet` package, imported unmodified through the stand-in."""
    import os
    if not os.path.isdir(os.path.join(ref_root, 'mmdet')):
        raise FileNotFoundError(f'{ref_root}/mmdet not found (reference only exists in the build container)')
    install()
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    import mmdet  # noqa: F401
    return mmdet
