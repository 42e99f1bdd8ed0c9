"""CPU: dry run of the HOST logic -- registries, module graph, autograd wiring, row-space bookkeeping, argument
marshalling of every C-ABI call -- with the shared library replaced by ctypes callbacks of the SAME signatures
(`ld_b200.lib._SIGS`) that do nothing and return 0.  ctypes still converts and checks every argument, the autograd
functions still have to hand back one gradient per input with the right shape, so Python-level mistakes in the train
step surface here without a GPU.  Kernel outputs are uninitialised memory: nothing numeric is asserted (that is what
the `-m gpu` parity tests are for)."""
import contextlib
import ctypes as C

import pytest
import torch

import ld_b200
from ld_b200 import lib as L, presets, synth
from ld_b200.config import ConfigDict


class _FakeStream:
    cuda_stream = 0
    device = torch.device('cpu')

    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass

    def record_event(self, *a):
        return None


@pytest.fixture
def fake_backend(monkeypatch):
    calls = []
    keep = []

    def make(name, restype, argtypes):
        def impl(*args):
            calls.append(name)
            if name.endswith('workspace_bytes'):
                return 1 << 16
            if name == 'ldb200_last_error':
                return b''
            return 0
        cb = C.CFUNCTYPE(restype, *argtypes)(impl)
        keep.append(cb)
        return cb

    class FakeLib:
        pass
    fake = FakeLib()
    for name, (restype, argtypes) in L._SIGS.items():
        setattr(fake, name, make(name, restype, argtypes))
    monkeypatch.setattr(L, '_lib', fake)
    monkeypatch.setattr(L, 'stream_ptr', lambda: C.c_void_p(0))
    stream = _FakeStream()
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a, **k: stream)
    monkeypatch.setattr(torch.cuda, 'Stream', lambda *a, **k: stream)
    monkeypatch.setattr(torch.cuda, 'stream', lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, 'current_device', lambda: 0)
    monkeypatch.setattr(torch.Tensor, 'record_stream', lambda self, s: None)
    return calls


def _batch(n=2, h=128, w=160, seed=9):
    b = synth.make_batch(n, h, w, seed=seed)
    return dict(img=b['img'], img_metas=b['img_metas'], gt_bboxes=b['gt_bboxes'], gt_labels=b['gt_labels'])


def _check_grads(model):
    missing = [k for k, p in model.named_parameters() if p.requires_grad and p.grad is None]
    assert not missing, missing[:8]
    for k, p in model.named_parameters():
        if p.grad is not None:
            assert p.grad.shape == p.shape, k


@pytest.mark.parametrize('imitation,im_weight', [('finegrained', 2.0), ('gibox', 2.0), ('gibox', 0.0)])
def test_ld_train_step_host_logic(fake_backend, imitation, im_weight):
    cfg = ConfigDict(presets.ld_model('r18', teacher_depth=18))
    cfg.bbox_head.update(imitation_method=imitation, loss_im=dict(type='IMLoss', loss_weight=im_weight))
    m = ld_b200.build_detector(cfg).train()
    losses = m(**_batch())
    assert list(losses) == ['loss_cls', 'loss_bbox', 'loss_dfl', 'loss_ld', 'loss_ld_vlr', 'loss_kd', 'loss_kd_neg',
                            'loss_im']
    assert all(len(v) == 5 for v in losses.values())
    loss, log = m._parse_losses(losses)
    loss.backward()
    _check_grads(m)
    assert ('ldb200_gi_region' in fake_backend) == (imitation == 'gibox' and im_weight != 0)
    for name in ('ldb200_ld_targets', 'ldb200_ld_loss_prepass', 'ldb200_ld_loss_fused', 'ldb200_conv2d_wgrad_tc',
                 'ldb200_stem_tc'):
        assert name in fake_backend, name
    assert 'ldb200_stem_im2col' not in fake_backend and not any(n.endswith('_v2') for n in fake_backend)


def test_ld_train_step_im2col_stem_host_logic(fake_backend, monkeypatch):
    """The fused dual stem is the default; LDB200_NO_FUSED_STEM=1 selects the im2col + GEMM route."""
    monkeypatch.setenv('LDB200_NO_FUSED_STEM', '1')
    m = ld_b200.build_detector(ConfigDict(presets.ld_model('r18', teacher_depth=18))).train()
    loss, _ = m._parse_losses(m(**_batch()))
    loss.backward()
    _check_grads(m)
    assert fake_backend.count('ldb200_stem_im2col') == 1 and 'ldb200_stem_tc' not in fake_backend


def test_flat_sgd_step_host_logic(fake_backend):
    m = ld_b200.build_detector(ConfigDict(presets.ld_model('r18', teacher_depth=18))).train()
    ts = ld_b200.LDTrainStep(m, lr=0.001, warmup_iters=4)
    for _ in range(2):
        ts.step(_batch(), sync_logs=False)
    assert ts.iter == 2 and fake_backend.count('ldb200_sgd_flat') == 2
    p = m.bbox_head.gfl_cls.weight
    assert p.data_ptr() >= ts.flat_w.data_ptr() and p.grad.data_ptr() >= ts.flat_g.data_ptr()


def test_ldv2_train_step_host_logic(fake_backend):
    """configs/ldv2/ld_r50_gflv2_r101_fpn_1x.py structure (R18 backbones): LDv2Head student, GFocalHead teacher."""
    cfg = ConfigDict(presets.ld_model('r18', teacher_depth=18))
    v2 = dict(reg_topk=4, reg_channels=64, add_mean=True,
              loss_cls=dict(type='QualityFocalLoss', use_sigmoid=False, beta=2.0, loss_weight=1.0))
    cfg.bbox_head.update(type='LDv2Head', imitation_method='gibox', loss_im=dict(type='IMLoss', loss_weight=2), **v2)
    cfg.teacher_config.model.bbox_head.update(type='GFocalHead', **v2)
    m = ld_b200.build_detector(cfg).train()
    losses = m(**_batch())
    assert all(len(v) == 5 for v in losses.values())
    loss, _ = m._parse_losses(losses)
    loss.backward()
    _check_grads(m)
    for name in ('ldb200_dgqp_forward', 'ldb200_dgqp_backward', 'ldb200_ld_loss_prepass_v2', 'ldb200_ld_loss_fused_v2',
                 'ldb200_gi_region_v2'):
        assert fake_backend.count(name) == 1, name
    assert 'ldb200_ld_loss_fused' not in fake_backend


def test_head_loss_apis_host_logic(fake_backend):
    from oracle import cases
    d = cases.loss_case_inputs(2, 128, 160, 3)
    cfg = ConfigDict(presets.ld_model('r50'))
    hc = cfg.bbox_head
    hc['train_cfg'], hc['test_cfg'] = cfg.train_cfg, cfg.test_cfg
    head = ld_b200.build_head(hc).train()
    leaves = {k: [t.clone().requires_grad_(True) for t in d[k]] for k in ('cls', 'reg', 'x')}
    losses = head.loss(leaves['cls'], leaves['reg'], d['gt_bboxes'], d['gt_labels'], (d['t_cls'], d['t_reg']),
                       leaves['x'], d['t_x'], d['img_metas'])
    sum(sum(v) for v in losses.values()).backward()
    assert all(t.grad is not None and t.grad.shape == t.shape for grp in leaves.values() for t in grp)
    # LDv2Head.loss(cls_scores, bbox_preds, cls_feat, ..., soft_teacher = (scores, bbox_preds, cls_feats), ...)
    d2 = cases.loss_case_inputs_v2(2, 128, 160, 4)
    hc2 = ConfigDict(presets.ld_model('r50')).bbox_head
    hc2.update(type='LDv2Head', reg_topk=4, reg_channels=64, add_mean=True, imitation_method='gibox',
               loss_cls=dict(type='QualityFocalLoss', use_sigmoid=False, beta=2.0, loss_weight=1.0),
               loss_im=dict(type='IMLoss', loss_weight=2))
    hc2['train_cfg'], hc2['test_cfg'] = cfg.train_cfg, cfg.test_cfg
    head2 = ld_b200.build_head(hc2).train()
    lv = {k: [t.clone().requires_grad_(True) for t in d2[k]] for k in ('score', 'cls_feat', 'reg', 'x')}
    losses = head2.loss(lv['score'], lv['reg'], lv['cls_feat'], d2['gt_bboxes'], d2['gt_labels'],
                        (None, d2['t_reg'], d2['t_cls']), lv['x'], d2['t_x'], d2['img_metas'])
    sum(sum(v) for v in losses.values()).backward()
    assert all(t.grad is not None and t.grad.shape == t.shape for grp in lv.values() for t in grp)


@pytest.mark.parametrize('student', ['r18', 'r50'])
def test_parity_mode_train_step_host_logic(fake_backend, student):
    """fp32 ("parity") mode: fp32 activations, every conv / wgrad on split-bf16 operands (ldb200_split_bf16 before
    each launch), same module graph."""
    from ld_b200 import engine
    m = ld_b200.build_detector(ConfigDict(presets.ld_model(student, teacher_depth=18))).train()
    ts = ld_b200.LDTrainStep(m, lr=0.001, warmup_iters=4, precision='fp32')
    ts.step(_batch(), sync_logs=False)
    assert not engine.parity()                      # the mode is scoped to the step
    convs = fake_backend.count('ldb200_conv2d_tc') + fake_backend.count('ldb200_conv2d_tc_multi')
    # every conv splits its activation (and, once, its weights); the leading plane product runs as several short
    # accumulation chains (one launch per tap and 256 input channels), so there are more launches than splits
    assert fake_backend.count('ldb200_split_bf16') > 100 and convs > fake_backend.count('ldb200_split_bf16')
    _check_grads(m)


def test_overlapped_gradient_exchange_covers_flat_buffer_once(fake_backend, monkeypatch):
    """Data-parallel exchange of LDTrainStep with the overlap opted in (LDB200_AR_OVERLAP=1): the backward
    pass runs in pieces (head, neck, layer4, layer3, rest), each piece's segment of the flat gradient buffer is
    all-reduced as soon as it is complete and the remainder after the last piece; together they cover every element
    exactly once, and the gradients are those of one backward call."""
    import torch.distributed as dist
    m = ld_b200.build_detector(ConfigDict(presets.ld_model('r50', teacher_depth=18))).train()
    calls = []
    monkeypatch.setattr(dist, 'all_reduce', lambda t, *a, **k: calls.append((t.data_ptr(), t.numel())))
    monkeypatch.setenv('LDB200_AR_OVERLAP', '1')
    ts = ld_b200.LDTrainStep(m, lr=0.001, warmup_iters=4)
    ts.world = 2                                  # pretend to be one of two ranks
    ts._init_overlap()
    assert ts._hook_handle is not None
    order = []
    orig = ts._reduce_span
    monkeypatch.setattr(ts, '_reduce_span', lambda seg: (order.append(seg), orig(seg))[1])
    ts.step(_batch(), sync_logs=False)
    _check_grads(m)
    base, es = ts.flat_g.data_ptr(), ts.flat_g.element_size()
    spans = sorted(((p - base) // es, (p - base) // es + n) for p, n in calls if n > 8)    # (8-byte normaliser excluded)
    assert spans[0][0] == 0 and spans[-1][1] == ts.flat_g.numel()
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:])), spans                       # disjoint, no gaps
    first = [s for i, s in enumerate(order) if s not in order[:i]]
    assert first == ['bbox_head.', 'neck.', 'backbone.layer4.', 'backbone.layer3.'], first
    assert len(spans) == 5                         # + the remainder (layer2) after the backward pass


def test_ldv2_x101_dcn_teacher_host_logic(fake_backend):
    """BASELINE configs[3] structure: LDv2Head student, GFocalHead teacher on ResNeXt-101 32x4d with DCN in stages 3-4
    (grouped convs as dense block-diagonal launches, DCN = offset conv + deform_im2col + GEMM), forward-only teacher."""
    m = ld_b200.build_detector(ConfigDict(presets.ldv2_model('r18'))).train()
    loss, _ = m._parse_losses(m(**_batch()))
    loss.backward()
    _check_grads(m)
    # 23 + 3 DCN bottlenecks in stages 3-4 of the X101 teacher
    assert fake_backend.count('ldb200_deform_im2col') == 26
    assert fake_backend.count('ldb200_ld_loss_fused_v2') == 1
