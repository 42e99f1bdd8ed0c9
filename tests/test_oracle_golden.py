"""CPU: the oracle restatement (oracle/ld_oracle.py) against the golden vectors frozen from the
UNMODIFIED reference (oracle/make_golden.py).  Integer targets bit-exact, floats <= 1e-5."""
import os

import pytest
import torch

from ld_b200 import synth
from oracle import cases, ld_oracle as lo


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)


@pytest.mark.parametrize('case', cases.TARGET_CASES, ids=lambda c: c[0])
def test_targets_bit_exact(golden_dir, case):
    name, n, h, w, seed = case
    gold = _load(golden_dir, name)
    gb, gl = synth.make_gts(n, h, w, seed)
    sizes = cases.featmap_sizes(h, w)
    metas = synth.make_img_metas(n, h, w)
    t = lo.ld_targets(sizes, cases.STRIDES, metas, gb, gl, imitation='finegrained')
    tc = lo.ld_targets(sizes, cases.STRIDES, metas, gb, gl, imitation='fitnet')
    assert torch.equal(torch.cat(t['labels'], 1), gold['labels'].long())
    assert torch.equal(torch.cat(t['label_weights'], 1), gold['label_weights'].float())
    assert torch.equal(torch.cat(t['bbox_targets'], 1), gold['bbox_targets'])
    assert torch.equal(torch.cat(t['vlr'], 1), gold['vlr'])          # float equality
    assert torch.equal(torch.cat(t['im'], 1), gold['im_finegrained'].float())
    assert torch.equal(torch.cat(tc['im'], 1), gold['im_center'].float())
    assert t['num_total_pos'] == gold['num_total_pos']
    for a, a0 in zip(t['anchors'], gold['level_anchor0']):
        assert torch.equal(a[:2], a0)   # first two anchors of the level


@pytest.mark.parametrize('case', cases.LOSS_CASES, ids=lambda c: c[0])
def test_loss_and_grads(golden_dir, case):
    name, n, h, w, seed = case
    gold = _load(golden_dir, name)
    d = cases.loss_case_inputs(n, h, w, seed)
    leaves = {k: [t.clone().requires_grad_(True) for t in d[k]] for k in ('cls', 'reg', 'x')}
    tg = lo.ld_targets(d['sizes'], cases.STRIDES, d['img_metas'], d['gt_bboxes'], d['gt_labels'])
    losses = lo.ld_loss(leaves['cls'], leaves['reg'], d['t_cls'], d['t_reg'], leaves['x'],
                        d['t_x'], tg, cases.STRIDES)
    total = sum(sum(v) for v in losses.values())
    total.backward()
    for k, v in gold['losses'].items():
        got = torch.stack([t.detach() for t in losses[k]])
        assert torch.allclose(got, v, rtol=1e-5, atol=1e-6), (k, got, v)
    for l in range(5):
        assert torch.allclose(leaves['cls'][l].grad, gold['grad_cls'][l], rtol=1e-5, atol=1e-7)
        assert torch.allclose(leaves['reg'][l].grad, gold['grad_reg'][l], rtol=1e-5, atol=1e-7)
        gx = leaves['x'][l].grad
        gx = torch.zeros_like(leaves['x'][l]) if gx is None else gx
        assert torch.allclose(gx.sum(1), gold['grad_x_rowsum'][l], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize('case', cases.LOSS_CASES, ids=lambda c: c[0] + '_gibox')
def test_loss_gibox(golden_dir, case):
    """imitation_method='gibox' (LDHead default, ld_head.py:170-176, 613-638): the GI-box NMS picks (indices,
    bit-exact) and the resulting loss_im / feature gradients against the unmodified reference."""
    name, n, h, w, seed = case
    gold = _load(golden_dir, name + '_gibox')
    d = cases.loss_case_inputs(n, h, w, seed)
    leaves = {k: [t.clone().requires_grad_(True) for t in d[k]] for k in ('cls', 'reg', 'x')}
    tg = lo.ld_targets(d['sizes'], cases.STRIDES, d['img_metas'], d['gt_bboxes'], d['gt_labels'])
    picked = []
    losses = lo.ld_loss(leaves['cls'], leaves['reg'], d['t_cls'], d['t_reg'], leaves['x'],
                        d['t_x'], tg, cases.STRIDES, imitation='gibox', gi_out=picked)
    sum(sum(v) for v in losses.values()).backward()
    assert len(picked) == 5
    for got, want in zip(picked, gold['gi_idx']):
        assert torch.equal(got.to(torch.int32), want), (got, want)
    for k, v in gold['losses'].items():
        got = torch.stack([t.detach() for t in losses[k]])
        assert torch.allclose(got, v, rtol=1e-5, atol=1e-6), (k, got, v)
    for l in range(5):
        gx = leaves['x'][l].grad
        gx = torch.zeros_like(leaves['x'][l]) if gx is None else gx
        assert torch.allclose(gx.sum(1), gold['grad_x_rowsum'][l], rtol=1e-4, atol=1e-6)
        assert torch.allclose(leaves['cls'][l].grad.norm(), gold['grad_cls_norm'][l], rtol=1e-5)
        assert torch.allclose(leaves['reg'][l].grad.norm(), gold['grad_reg_norm'][l], rtol=1e-5)


@pytest.mark.parametrize('case', cases.STEP_CASES + [cases.STEP_CASE_CONFIG1], ids=lambda c: c[0])
def test_full_step(golden_dir, case):
    name, sd, td, n, h, w, seed = case
    gold = _load(golden_dir, name)
    m = lo.OracleLDDetector(sd, td).train()
    synth.seeded_init(m, seed)
    synth.seeded_init(m.teacher_model, seed + 1, prefix='teacher.')
    b = synth.make_batch(n, h, w, seed=1000 + seed)
    losses, loss, log = m.step(b)
    for k, v in gold['log_vars'].items():
        assert abs(log[k] - v) <= 1e-5 * max(1.0, abs(v)), (k, log[k], v)
    named = dict(m.named_parameters())
    assert sorted(k for k, p in named.items() if p.grad is None) == sorted(gold['no_grad'])
    for k, st in gold['grad_stats'].items():
        g = named[k].grad
        assert abs(float(g.double().norm()) - float(st[0])) <= 1e-4 * float(st[0]) + 1e-9, k
    for k, g in gold['grads'].items():
        assert torch.allclose(named[k].grad, g, rtol=1e-4, atol=1e-6 * float(g.abs().max())), k
    for k, pr in gold.get('grad_proj', {}).items():     # fingerprint of every element of every gradient tensor
        nrm = float(gold['grad_stats'][k][0])
        assert abs(cases.grad_projection(k, named[k].grad) - pr) <= 1e-4 * nrm * named[k].numel() ** 0.5 + 1e-9, k


@pytest.mark.parametrize('case', cases.LOSS_CASES_V2, ids=lambda c: c[0])
def test_loss_ldv2(golden_dir, case):
    """`LDv2Head.loss` (ld_gflv2.py, configs/ldv2/ld_r50_gflv2_r101_fpn_1x.py): QFL on joint scores, KD on cls_feat,
    raw-score GI boxes, 81 class channels -- restatement against the unmodified reference."""
    name, n, h, w, seed = case
    gold = _load(golden_dir, name)
    d = cases.loss_case_inputs_v2(n, h, w, seed)
    leaves = {k: [t.clone().requires_grad_(True) for t in d[k]] for k in ('score', 'cls_feat', 'reg', 'x')}
    tg = lo.ld_targets(d['sizes'], cases.STRIDES, d['img_metas'], d['gt_bboxes'], d['gt_labels'])
    picked = []
    losses = lo.ld_loss(leaves['score'], leaves['reg'], d['t_cls'], d['t_reg'], leaves['x'], d['t_x'], tg,
                        cases.STRIDES, imitation='gibox', gi_out=picked, cls_feats=leaves['cls_feat'], num_classes=80)
    sum(sum(v) for v in losses.values()).backward()
    for got, want in zip(picked, gold['gi_idx']):
        assert torch.equal(got.to(torch.int32), want), (got, want)
    for k, v in gold['losses'].items():
        got = torch.stack([t.detach() for t in losses[k]])
        assert torch.allclose(got, v, rtol=1e-5, atol=1e-6), (k, got, v)

    def g(t):
        return torch.zeros_like(t) if t.grad is None else t.grad
    for l in range(5):
        assert torch.allclose(g(leaves['score'][l]), gold['grad_score'][l], rtol=1e-5, atol=1e-7)
        assert torch.allclose(g(leaves['reg'][l]), gold['grad_reg'][l], rtol=1e-5, atol=1e-7)
        assert torch.allclose(g(leaves['cls_feat'][l]).sum(1), gold['grad_cls_feat_rowsum'][l], rtol=1e-4, atol=1e-7)
        assert torch.allclose(g(leaves['x'][l]).sum(1), gold['grad_x_rowsum'][l], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize('case', cases.STEP_CASES_V2, ids=lambda c: c[0])
def test_full_step_ldv2(golden_dir, case):
    """Whole LDv2 iteration (GFocalHead student + teacher with the DGQP branch, gfocal_head.py:180-217)."""
    name, sd, td, n, h, w, seed = case
    gold = _load(golden_dir, name)
    m = lo.OracleLDDetector(sd, td, imitation='gibox', v2=True,
                            loss_cfg=dict(w_ld_vlr=0.25, T_ld_vlr=10.0, w_kd=10.0, T_kd=2.0, w_im=2.0)).train()
    synth.seeded_init(m, seed)
    synth.seeded_init(m.teacher_model, seed + 1, prefix='teacher.')
    b = synth.make_batch(n, h, w, seed=1000 + seed)
    losses, loss, log = m.step(b)
    for k, v in gold['log_vars'].items():
        assert abs(log[k] - v) <= 1e-5 * max(abs(v), 1.0), (k, log[k], v)
    named = dict(m.named_parameters())
    assert set(gold['no_grad']) == {k for k, p in named.items() if p.grad is None}
    for k, ref in gold['grads'].items():
        assert torch.allclose(named[k].grad, ref, rtol=1e-4, atol=1e-7), k
    for k, st in gold['grad_stats'].items():
        gn = float(named[k].grad.double().norm())
        assert abs(gn - float(st[0])) <= 1e-4 * max(float(st[0]), 1e-6), (k, gn, float(st[0]))


def test_loss_curve_first_iterations(golden_dir):
    """training-trajectory parity.  The oracle port under torch SGD + mmcv's linear warm-up follows the
    unmodified reference's 100-iteration loss curve (tests/golden/c_*.pt); the first 25 iterations are replayed here
    (the GPU test replays all 100 through LDTrainStep)."""
    name, sd, td, n, h, w, seed, iters, base_lr, warm = cases.CURVE_CASE
    gold = _load(golden_dir, name)
    m = lo.OracleLDDetector(sd, td).train()
    synth.seeded_init(m, seed)
    synth.seeded_init(m.teacher_model, seed + 1, prefix='teacher.')
    opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=base_lr, momentum=0.9, weight_decay=1e-4)
    for it in range(25):
        lr = base_lr * (1 - (1 - it / warm) * (1 - 0.001)) if it < warm else base_lr
        for gp in opt.param_groups:
            gp['lr'] = lr
        losses, loss, log = m.step(synth.make_batch(n, h, w, seed=2000 + it))
        opt.step()
        want = float(gold['loss'][it])
        assert abs(float(loss) - want) <= 1e-3 * abs(want), (it, float(loss), want)


def test_reference_known_answers():
    """The only known-answer vectors the reference's own tests/doctests hold on this path
    (SURVEY.md §8c): tests/test_anchor.py:22-41 `test_strides` and the bbox_overlaps examples of
    iou2d_calculator.py:66-87 (values probed from the reference)."""
    # test_strides: AnchorGenerator([10],[1.],[1.],[10]) on a 2x2 map -> base size 10, i.e. half = 5.
    # Our generator is the octave_base_scale=8 special case: side 8*stride; check layout + centre rule
    a = lo.level_anchors(2, 2, 10)
    exp = torch.tensor([[-5., -5., 5., 5.], [5., -5., 15., 5.], [-5., 5., 5., 15.], [5., 5., 15., 15.]])
    assert torch.equal((a[:, :2] + a[:, 2:]) / 2, (exp[:, :2] + exp[:, 2:]) / 2)   # centres (x*s, y*s), x fastest
    assert torch.equal(a[:, 2:] - a[:, :2], torch.full((4, 2), 80.))               # side 8*stride
    b1 = torch.tensor([[0., 0., 10., 10.], [10., 10., 20., 20.], [32., 32., 38., 42.]])
    b2 = torch.tensor([[0., 0., 10., 20.], [0., 10., 10., 19.], [10., 10., 20., 20.]])
    iou = lo.pair_overlaps(b1, b2, 'iou')
    assert torch.allclose(iou, torch.tensor([[.5, 0, 0], [0, 0, 1.], [0, 0, 0]]))
    assert torch.allclose(lo.aligned_overlaps(b1, b2, 'giou'), torch.tensor([.5, -.05, -.8214]), atol=1e-4)
    d = torch.diagonal(lo.pair_overlaps(b1, b2, 'diou'))
    assert torch.allclose(d, torch.tensor([.95, -.2005, -.4889]), atol=1e-4)       # IoF-based quirk


@pytest.mark.parametrize('case', cases.INFER_CASES, ids=lambda c: c[0])
def test_get_bboxes(golden_dir, case):
    """Inference restatement (`lo.get_bboxes`: top-k, Integral, distance2bbox + clip, score threshold, class-offset NMS)
    against `GFLHead.get_bboxes` of the unmodified reference (gfl_head.py:354-451, bbox_nms.py:70-193)."""
    name, n, h, w, seed, rescale = case
    gold = _load(golden_dir, name)
    d = cases.infer_case_inputs(n, h, w, seed)
    res = lo.get_bboxes(d['cls'], d['reg'], cases.STRIDES, [m['img_shape'] for m in d['img_metas']],
                        [m['scale_factor'] for m in d['img_metas']] if rescale else None)
    for (det, lab), gd, gl in zip(res, gold['dets'], gold['labels']):
        assert torch.equal(lab, gl)
        assert torch.allclose(det, gd, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize('case', cases.PIPELINE_CASES, ids=lambda c: c[0])
def test_input_pipeline_restatement(golden_dir, case):
    """oracle/pipeline_ref.py (cv2 / numpy restatement of mmcv.imrescale / imflip / imnormalize / impad in the
    reference's transform order) is BIT-identical to the reference's own Resize / RandomFlip / Normalize / Pad classes."""
    import hashlib
    import numpy as np
    from oracle import pipeline_ref as pr
    name, sizes, seed, flips, img_scale = case
    gold = _load(golden_dir, name)
    imgs, boxes, _ = pr.synth_images(sizes, seed)
    batch, metas, bb = pr.train_pipeline(imgs, boxes, flips, img_scale)
    for i in range(len(imgs)):
        ph, pw = gold['pad_shape'][i][:2]
        chw = np.ascontiguousarray(batch[i, :, :ph, :pw].numpy())
        assert hashlib.sha256(chw.tobytes()).hexdigest() == gold['sha'][i]
        assert float(batch[i, :, ph:].abs().sum()) == 0 and float(batch[i, :, :, pw:].abs().sum()) == 0
        assert tuple(metas[i]['img_shape']) == gold['img_shape'][i] and tuple(metas[i]['pad_shape']) == gold['pad_shape'][i]
        assert torch.equal(torch.from_numpy(bb[i]), gold['boxes'][i])
        assert torch.equal(torch.from_numpy(metas[i]['scale_factor']), gold['scale_factor'][i])
