"""CPU: the reference's own config files load through ld_b200.Config and their `model` section builds
UNCHANGED through the ld_b200 registries (build container only: needs /root/reference); the presets used
on the GPU box are the same dictionaries."""
import os

import pytest
import torch

REF = '/root/reference'
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'configs')), reason='reference tree absent')


def _norm(d):
    if isinstance(d, dict):
        return {k: _norm(v) for k, v in d.items()}
    if isinstance(d, (list, tuple)):
        return [_norm(v) for v in d]
    return d


@needs_ref
def test_reference_config_builds_unchanged(monkeypatch):
    import ld_b200
    from ld_b200 import presets
    monkeypatch.chdir(REF)    # the reference resolves teacher_config relative to CWD (kd_one_stage.py:39-40)
    cfg = ld_b200.Config.fromfile('configs/ld/ld_r50_gflv1_r101_fpn_coco_1x.py')
    assert cfg.optimizer.lr == 0.0025 and cfg.data.samples_per_gpu == 2       # child overrides base
    assert cfg.lr_config.warmup_iters == 500                                   # inherited through two levels
    model = ld_b200.build_detector(cfg.model, train_cfg=cfg.get('train_cfg'), test_cfg=cfg.get('test_cfg'))
    assert type(model).__name__ == 'KnowledgeDistillationSingleStageDetector'
    assert type(model.bbox_head).__name__ == 'LDHead' and model.bbox_head.imitation_method == 'finegrained'
    assert model.bbox_head.loss_im.loss_weight == 2.0 and model.bbox_head.loss_kd.T == 2
    assert model.teacher_model.backbone.depth == 101 and model.backbone.depth == 50
    assert 'teacher_model' not in dict(model.named_modules()) and not any('teacher' in k for k in model.state_dict())
    n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)
    assert n_train == 32215193
    # presets == the reference files
    want = _norm(cfg.model.to_dict())
    got = _norm(presets.ld_model('r50'))
    for k in ('backbone', 'neck', 'bbox_head', 'train_cfg', 'test_cfg', 'output_feature', 'type'):
        assert got[k] == want[k], k
    tcfg = ld_b200.Config.fromfile('configs/gfl/gfl_r101_fpn_mstrain_2x_coco.py')
    tw = _norm(tcfg.model.to_dict())
    tg = _norm(presets.gfl_model(101))
    for k in ('backbone', 'neck', 'bbox_head', 'train_cfg', 'test_cfg', 'type'):
        assert tg[k] == tw[k], k


@needs_ref
def test_r18_config_and_registry_errors(monkeypatch):
    import ld_b200
    monkeypatch.chdir(REF)
    cfg = ld_b200.Config.fromfile('configs/ld/ld_r18_gflv1_r101_fpn_coco_1x.py')
    model = ld_b200.build_detector(cfg.model)
    assert model.bbox_head.imitation_method == 'gibox'       # LDHead ctor default (ld_head.py:62)
    with pytest.raises(KeyError):
        ld_b200.build_loss(dict(type='NoSuchLoss'))
    with pytest.raises(TypeError):
        ld_b200.build_loss(dict(type='IMLoss', bogus=1))


@needs_ref
def test_ldv2_config_builds_with_reference_keys(monkeypatch):
    """configs/ldv2/ld_r50_gflv2_r101_fpn_1x.py (LDv2Head student, GFocalHead teacher): the registry
    surface, parameters and checkpoint keys are in place (the compute path is CUDA-only: tests/test_gpu_ldv2.py)."""
    import ld_b200
    from oracle import ld_oracle as lo
    monkeypatch.chdir(REF)
    cfg = ld_b200.Config.fromfile('configs/ldv2/ld_r50_gflv2_r101_fpn_1x.py')
    model = ld_b200.build_detector(cfg.model)
    head = model.bbox_head
    assert type(head).__name__ == 'LDv2Head' and type(model.teacher_model.bbox_head).__name__ == 'GFocalHead'
    assert head.cls_out_channels == 81 and head.imitation_method == 'gibox' and head.loss_im.loss_weight == 2
    assert (head.reg_topk, head.reg_channels, head.add_mean) == (4, 64, True)
    o = lo.OracleLDDetector(50, 101, v2=True)
    assert set(model.state_dict()) == set(o.state_dict())
    assert all(model.state_dict()[k].shape == v.shape for k, v in o.state_dict().items())
    w1, b1, w2, b2 = head.quality_params()
    assert w1.shape == (64, 20) and b1.shape == (64,) and w2.shape == (64,) and b2.shape == (1,)
    with pytest.raises(NotImplementedError):      # plain GFLv2 training is not on the LD hot path
        model.teacher_model.bbox_head.forward_train(None)


def test_state_dict_keys_match_oracle_port():
    import ld_b200
    from ld_b200 import presets
    from ld_b200.config import ConfigDict
    from oracle import ld_oracle as lo
    m = ld_b200.build_detector(ConfigDict(presets.ld_model('r18', teacher_depth=18)))
    o = lo.OracleLDDetector(18, 18)
    assert set(m.state_dict()) == set(o.state_dict())
    assert set(m.teacher_model.state_dict()) == set(o.teacher_model.state_dict())
    for k, v in o.state_dict().items():
        assert m.state_dict()[k].shape == v.shape, k
    # frozen stem + layer1 (resnet.py:572-588), BN always eval (norm_eval)
    m.train()
    assert not m.backbone.conv1.weight.requires_grad and not m.backbone.layer1[0].conv1.weight.requires_grad
    assert m.backbone.layer2[0].conv1.weight.requires_grad and not m.backbone.layer2[0].bn1.training
    assert not m.teacher_model.training


def test_warmup_schedule():
    from ld_b200.detectors import LDTrainStep
    s = LDTrainStep.__new__(LDTrainStep)
    s.base_lr, s.warmup_iters, s.warmup_ratio = 0.01, 500, 0.001
    assert abs(s.lr_at(0) - 0.01 * 0.001) < 1e-12 and abs(s.lr_at(250) - 0.01 * (1 - 0.5 * 0.999)) < 1e-12
    assert s.lr_at(500) == 0.01


def test_missing_checkpoints_are_errors_by_default(monkeypatch):
    """mmcv's load_checkpoint raises on a missing file (kd_one_stage.py:40-44); so does the port unless random
    initialisation is explicitly allowed (a silently random teacher trains against noise)."""
    import pytest
    import ld_b200
    from ld_b200 import detectors, presets
    from ld_b200.config import ConfigDict
    monkeypatch.delenv('LDB200_ALLOW_RANDOM_INIT', raising=False)
    cfg = ConfigDict(presets.ld_model('r18', teacher_depth=18))
    cfg.teacher_config.model.pretrained = None
    cfg.teacher_ckpt = '/nonexistent/gfl_r101.pth'
    with pytest.raises(FileNotFoundError):
        ld_b200.build_detector(cfg)
    cfg.teacher_ckpt = None
    cfg.pretrained = 'torchvision://resnet18'
    monkeypatch.setenv('LD_B200_PRETRAINED_DIR', '/nonexistent')
    monkeypatch.setattr('torch.hub.get_dir', lambda: '/nonexistent')
    with pytest.raises(FileNotFoundError):
        ld_b200.build_detector(cfg)
    with detectors.allow_random_init(), pytest.warns(UserWarning, match='RANDOM'):
        ld_b200.build_detector(cfg)


def test_train_step_schedule_and_state_dict():
    """LrUpdaterHook arithmetic (linear warm-up x step decay) and optimizer-state round trip of LDTrainStep."""
    import torch
    import ld_b200
    from ld_b200 import presets
    from ld_b200.config import ConfigDict
    m = ld_b200.build_detector(ConfigDict(presets.ld_model('r18', teacher_depth=18)))
    ts = ld_b200.LDTrainStep(m, lr=0.01, warmup_iters=500, warmup_ratio=0.001, lr_steps=(8 * 7330, 11 * 7330))
    assert abs(ts.lr_at(0) - 0.01 * 0.001) < 1e-12 and abs(ts.lr_at(250) - 0.01 * (1 - 0.5 * 0.999)) < 1e-12
    assert ts.lr_at(500) == 0.01 and abs(ts.lr_at(8 * 7330) - 0.001) < 1e-12 and abs(ts.lr_at(11 * 7330) - 1e-4) < 1e-12
    ts.flat_m.normal_()
    ts.iter = 1234
    sd = ts.state_dict()
    ts2 = ld_b200.LDTrainStep(m, lr=0.5)
    ts2.load_state_dict(sd)
    assert ts2.iter == 1234 and ts2.base_lr == 0.01 and torch.equal(ts2.flat_m, sd['momentum_buffer'])
    assert ts2.lr_at(9 * 7330) == ts.lr_at(9 * 7330)


def test_x101_dcn_teacher_builds_with_reference_keys_and_oracle_matches_reference():
    """BASELINE configs[3] teacher (configs/imv2/gflv2_x101_fpn_2x_coco.py: ResNeXt-101 32x4d, DCN in stages 3-4,
    GFocalHead): the registry builds it from the reference's own file with the reference's state-dict keys, and the
    oracle's ResNeXt/DCN port reproduces the UNMODIFIED reference backbone (run through the mmcv stand-in, whose
    DeformConv2dPack restates mmcv on torchvision.ops.deform_conv2d) on a seeded input."""
    import os
    import pytest
    import torch
    from oracle import ref_harness as rh
    if not rh.reference_available():
        pytest.skip('needs /root/reference')
    import ld_b200
    from ld_b200 import synth
    from ld_b200.config import Config, ConfigDict
    from oracle import ld_oracle as lo, mmcv_standin
    cfg = Config.fromfile(os.path.join(rh.REF_ROOT, 'configs/imv2/gflv2_x101_fpn_2x_coco.py'))
    mine = ld_b200.build_detector(ConfigDict(cfg.model.to_dict()))
    mmcv_standin.import_reference(rh.REF_ROOT)
    from mmdet.models import build_backbone
    bcfg = cfg.model.backbone.to_dict()
    ref_bb = build_backbone(dict(bcfg))
    ref_keys = sorted('backbone.' + k for k in ref_bb.state_dict())
    assert sorted(k for k in mine.state_dict() if k.startswith('backbone.')) == ref_keys
    # a shallower variant of the same family for the numeric check (depth 50, same groups / DCN stages)
    bcfg.update(depth=50)
    ref50 = build_backbone(dict(bcfg))
    ref50.eval()                     # (the reference's ResNet.train() does not return self)
    ora50 = lo.OracleResNet(50, groups=32, base_width=4, stage_with_dcn=(False, False, True, True))
    ora50.eval()
    synth.seeded_init(ref50, 7)
    synth.seeded_init(ora50, 7)
    assert sorted(ref50.state_dict()) == sorted(ora50.state_dict())
    x = torch.randn(1, 3, 96, 128, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        a, b = ref50(x), ora50(x)
    for u, v in zip(a, b):
        assert torch.allclose(u, v, rtol=1e-5, atol=1e-6)
