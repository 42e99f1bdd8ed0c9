"""Run the UNMODIFIED reference (`/root/reference/mmdet`) on CPU through the mmcv stand-in.

TEST INFRASTRUCTURE ONLY (build container only; `/root/reference` is absent on the GPU box).
Used by `oracle/make_golden.py` to freeze golden vectors under `tests/golden/` and by CPU tests
that pin the oracle restatement (`oracle/ld_oracle.py`) against the real reference code.
"""
import os

import torch

from . import mmcv_standin

REF_ROOT = '/root/reference'


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, 'mmdet'))


def load_model_cfg(config_rel='configs/ld/ld_r50_gflv1_r101_fpn_coco_1x.py', **head_overrides):
    """`model` section of a reference config, teacher config resolved to a dict.

    The reference resolves `teacher_config='configs/gfl/...'` against CWD
    (`kd_one_stage.py:39-40` -> `mmcv.Config.fromfile`); we resolve it against REF_ROOT and hand
    the detector the loaded dict (the `isinstance(teacher_config, str)` branch is skipped)."""
    from ld_b200.config import Config, ConfigDict
    cfg = Config.fromfile(os.path.join(REF_ROOT, config_rel))
    model = ConfigDict(cfg.model.to_dict())
    model['pretrained'] = None
    if 'teacher_config' in model:
        tc = model['teacher_config']
        if isinstance(tc, str):
            tcfg = Config.fromfile(os.path.join(REF_ROOT, tc))
            tmodel = tcfg.model.to_dict()
            tmodel['pretrained'] = None
            model['teacher_config'] = ConfigDict(model=tmodel)
        model['teacher_ckpt'] = None
    for k, v in head_overrides.items():
        model['bbox_head'][k] = v
    return model


def build_reference_detector(model_cfg):
    mmcv_standin.import_reference(REF_ROOT)
    from mmdet.models import build_detector
    model = build_detector(model_cfg)
    model.train()
    return model


def reference_step(model, batch, backward=True):
    """One forward(+backward) of the reference detector. Returns (losses, loss, log_vars)."""
    for p in model.parameters():
        p.grad = None
    losses = model(img=batch['img'], img_metas=batch['img_metas'],
                   gt_bboxes=batch['gt_bboxes'], gt_labels=batch['gt_labels'],
                   return_loss=True)
    loss, log_vars = model._parse_losses(losses)
    if backward:
        loss.backward()
    return losses, loss, log_vars


def reference_head(cfg_overrides=None, config_rel='configs/ld/ld_r50_gflv1_r101_fpn_coco_1x.py'):
    """A bare reference `LDHead` / `LDv2Head` (no backbone) for target / loss_single goldens."""
    mmcv_standin.import_reference(REF_ROOT)
    from mmdet.models import build_head
    model_cfg = load_model_cfg(config_rel, **(cfg_overrides or {}))
    head_cfg = model_cfg['bbox_head']
    head_cfg['train_cfg'] = model_cfg['train_cfg']
    head_cfg['test_cfg'] = model_cfg['test_cfg']
    head = build_head(head_cfg)
    head.train()
    return head
