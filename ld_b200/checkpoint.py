"""mmcv-format checkpoint I/O for the LD path: what mmcv's CheckpointHook / load_checkpoint and
`tools/model_converters/publish_model.py:16-32` do, so that checkpoints move between the reference and this build.

A checkpoint is `dict(meta=..., state_dict=OrderedDict, optimizer=...)`; `state_dict` keys are the reference's (the
modules here keep the same parameter names; the teacher is not part of the detector's state_dict, kd_one_stage.py:97-108).
"""
import hashlib
import os
import time
from collections import OrderedDict

import torch


def weights_to_cpu(state_dict):
    out = OrderedDict()
    for k, v in state_dict.items():
        out[k] = v.detach().cpu().contiguous()      # channels-last parameter storage is an implementation detail
    return out


def save_checkpoint(model, filename, optimizer=None, meta=None):
    """mmcv.runner.save_checkpoint: meta gets `time` (+ CLASSES when the model carries them); `optimizer` may be an
    `LDTrainStep` (its `state_dict()` = momentum buffer + iteration) or anything with `state_dict()`."""
    meta = dict(meta or {})
    meta.update(time=time.asctime())
    if hasattr(model, 'CLASSES') and model.CLASSES is not None:
        meta.update(CLASSES=model.CLASSES)
    module = model.module if hasattr(model, 'module') else model
    ckpt = dict(meta=meta, state_dict=weights_to_cpu(module.state_dict()))
    if optimizer is not None:
        ckpt['optimizer'] = optimizer.state_dict()
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    torch.save(ckpt, filename)
    return filename


def load_checkpoint(model, filename, map_location='cpu', strict=False, optimizer=None):
    """mmcv.runner.load_checkpoint: accepts a bare state_dict or a checkpoint dict, strips a DataParallel `module.`
    prefix, raises FileNotFoundError on a missing file, reports (strict=False) or raises on (strict=True) key mismatches.
    Returns the checkpoint dict."""
    if not os.path.isfile(filename):
        raise FileNotFoundError(f'{filename} is not a checkpoint file')
    ckpt = torch.load(filename, map_location=map_location, weights_only=False)
    sd = ckpt['state_dict'] if isinstance(ckpt, dict) and 'state_dict' in ckpt else ckpt
    if not isinstance(sd, dict):
        raise RuntimeError(f'No state_dict found in checkpoint file {filename}')
    if list(sd.keys()) and list(sd.keys())[0].startswith('module.'):
        sd = OrderedDict((k[7:], v) for k, v in sd.items())
    res = model.load_state_dict(sd, strict=strict)
    if res.missing_keys or res.unexpected_keys:
        import warnings
        warnings.warn(f'load_checkpoint({filename}): missing keys {res.missing_keys[:8]} unexpected keys '
                      f'{res.unexpected_keys[:8]} ({len(res.missing_keys)} / {len(res.unexpected_keys)})')
    for p in model.parameters():          # cached operand copies of the tcgen05 convs follow the weights
        for k in ('_ldb_packed', '_ldb_dgrad', '_ldb_packed32', '_ldb_dgrad32'):
            p.__dict__.pop(k, None)
    if optimizer is not None and isinstance(ckpt, dict) and 'optimizer' in ckpt:
        optimizer.load_state_dict(ckpt['optimizer'])
    return ckpt


def publish_model(in_file, out_file):
    """tools/model_converters/publish_model.py:16-32: drop the optimizer state, save, append the first 8 hex digits of
    the file's sha256 to the name.  Returns the final file name."""
    ckpt = torch.load(in_file, map_location='cpu', weights_only=False)
    if 'optimizer' in ckpt:
        del ckpt['optimizer']
    torch.save(ckpt, out_file)
    h = hashlib.sha256()
    with open(out_file, 'rb') as f:
        for chunk in iter(lambda: f.read(1 << 20), b''):
            h.update(chunk)
    stem = out_file[:-4] if out_file.endswith('.pth') else out_file
    final = f'{stem}-{h.hexdigest()[:8]}.pth'
    os.replace(out_file, final)
    return final
