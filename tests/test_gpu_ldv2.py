"""GPU: GFLv2 quality branch (ldb200_dgqp_forward / _backward), the LDv2 loss path (V2 prepass / GI region / fused
loss, `LDv2Head.loss`, the LDv2 train step), the fused dual stem and the 100-iteration loss curve -- against torch
autograd of gfocal_head.py:198-215 and the golden vectors of the unmodified reference (tests/golden/l2_*.pt, s2_*.pt,
c_*.pt).  The row arithmetic is also pinned on CPU (tests/test_v2_math_host.py) and the oracle against the reference
(tests/test_oracle_golden.py).  First GPU run: round 2 (gpurun_out/r02_pending.log, 11/13 green; the two failures
were test tolerances on cancelling sums, fixed below)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu]


def dgqp_ref(reg, cls_feat, W1, b1, W2, b2, bins=17, topk=4):
    prob = F.softmax(reg.reshape(-1, 4, bins), dim=2)
    top = prob.topk(topk, dim=2)[0]
    stat = torch.cat([top, top.mean(dim=2, keepdim=True)], dim=2).reshape(reg.shape[0], -1)
    q = torch.sigmoid(F.linear(F.relu(F.linear(stat, W1, b1)), W2[None], b2)).squeeze(1)
    return q, cls_feat.sigmoid() * q[:, None]


@pytest.mark.parametrize('rows', [1, 63, 64, 1000, 4321])
def test_dgqp_forward_backward(rows):
    from ld_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(rows)
    dev, C, bins, topk, hid = 'cuda', 81, 17, 4, 64
    reg = (torch.randn(rows, 4 * bins, device=dev, generator=g) * 2).requires_grad_(True)
    f = (torch.randn(rows, C, device=dev, generator=g) * 1.5 - 2).requires_grad_(True)
    W1 = (torch.randn(hid, 4 * (topk + 1), device=dev, generator=g) * 0.5).requires_grad_(True)
    b1 = (torch.randn(hid, device=dev, generator=g) * 0.2).requires_grad_(True)
    W2 = (torch.randn(hid, device=dev, generator=g) * 0.5).requires_grad_(True)
    b2 = torch.full((1,), 0.1, device=dev, requires_grad=True)
    q_ref, s_ref = dgqp_ref(reg, f, W1, b1, W2, b2, bins, topk)
    d_score = torch.randn(rows, C, device=dev, generator=g)
    d_kd = torch.randn(rows, C, device=dev, generator=g) * 0.1
    d_reg0 = torch.randn(rows, 4 * bins, device=dev, generator=g)
    (s_ref * d_score).sum().backward()

    det = [t.detach().contiguous() for t in (reg, f, W1, b1, W2, b2)]
    q, score = ops.dgqp_forward(*det, reg_topk=topk, add_mean=True)
    assert torch.allclose(q, q_ref.detach(), rtol=1e-5, atol=1e-6)
    assert torch.allclose(score, s_ref.detach(), rtol=1e-5, atol=1e-7)
    d_reg = d_reg0.clone()
    d_f, dW1, db1, dW2, db2 = ops.dgqp_backward(*det, q, d_score, d_reg, d_kd=d_kd, reg_topk=topk, add_mean=True)
    torch.cuda.synchronize()

    def close(got, want, name, tol=1e-4):
        err = float((got - want).abs().max())
        assert err <= tol * float(want.abs().max()) + 1e-7, (name, err)
    close(d_f, f.grad + d_kd, 'd_cls_feat')
    close(d_reg - d_reg0, reg.grad, 'd_reg')
    close(dW1, W1.grad, 'dW1', 5e-4)
    close(db1, b1.grad, 'db1', 5e-4)
    close(dW2, W2.grad, 'dW2', 5e-4)
    close(db2, b2.grad, 'db2', 5e-4)


def test_dgqp_backward_bf16_grads():
    from ld_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(7)
    dev, rows, C, bins, topk, hid = 'cuda', 777, 81, 17, 4, 64
    reg = torch.randn(rows, 4 * bins, device=dev, generator=g) * 2
    f = torch.randn(rows, C, device=dev, generator=g) * 1.5 - 2
    W1 = torch.randn(hid, 4 * (topk + 1), device=dev, generator=g) * 0.5
    b1 = torch.randn(hid, device=dev, generator=g) * 0.2
    W2 = torch.randn(hid, device=dev, generator=g) * 0.5
    b2 = torch.full((1,), 0.1, device=dev)
    q, _ = ops.dgqp_forward(reg, f, W1, b1, W2, b2)
    d_score = torch.randn(rows, 88, device=dev, generator=g)      # padded leading dimension, as the loss kernel emits
    d_reg32, d_reg16 = torch.zeros(rows, 72, device=dev), torch.zeros(rows, 72, device=dev, dtype=torch.bfloat16)
    out32 = ops.dgqp_backward(reg, f, W1, b1, W2, b2, q, d_score, d_reg32)
    out16 = ops.dgqp_backward(reg, f, W1, b1, W2, b2, q, d_score.bfloat16(), d_reg16)
    torch.cuda.synchronize()
    assert float((out16[0][:, :C].float() - out32[0][:, :C]).abs().max()) <= 2e-2 * float(out32[0][:, :C].abs().max())
    assert float((d_reg16[:, :68].float() - d_reg32[:, :68]).abs().max()) <= 2e-2 * float(d_reg32.abs().max()) + 1e-6
    # parameter gradients are sums over all rows of bf16-rounded contributions of both signs (db2 is ONE such sum):
    # bound the difference by the rounding of the summands, not by the (cancelling) total
    for a, b in zip(out16[1:], out32[1:]):
        assert float((a - b).abs().max()) <= 5e-2 * float(b.abs().max()) + 1e-6


# ---------------------------------------------------------------------------------------------------------
# LDv2Head.loss through the V2 loss kernels against the unmodified reference (tests/golden/l2_*.pt)
def make_v2_head():
    import ld_b200
    from ld_b200 import presets
    from ld_b200.config import ConfigDict
    cfg = ConfigDict(presets.ld_model('r50'))
    head_cfg = cfg.bbox_head
    # configs/ldv2/ld_r50_gflv2_r101_fpn_1x.py:27-56
    head_cfg.update(type='LDv2Head', reg_topk=4, reg_channels=64, add_mean=True, imitation_method='gibox',
                    loss_cls=dict(type='QualityFocalLoss', use_sigmoid=False, beta=2.0, loss_weight=1.0),
                    loss_im=dict(type='IMLoss', loss_weight=2))
    head_cfg['train_cfg'] = cfg.train_cfg
    head_cfg['test_cfg'] = cfg.test_cfg
    return ld_b200.build_head(head_cfg).cuda().train()


def test_ldv2_head_loss_api_matches_reference(golden_dir):
    import os
    from oracle import cases
    for name, n, h, w, seed in cases.LOSS_CASES_V2:
        gold = torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)
        d = cases.loss_case_inputs_v2(n, h, w, seed)
        head = make_v2_head()
        dev = torch.device('cuda')
        leaves = {k: [t.to(dev).requires_grad_(True) for t in d[k]] for k in ('score', 'cls_feat', 'reg', 'x')}
        losses = head.loss(leaves['score'], leaves['reg'], leaves['cls_feat'], [b.to(dev) for b in d['gt_bboxes']],
                           [b.to(dev) for b in d['gt_labels']],
                           (None, [t.to(dev) for t in d['t_reg']], [t.to(dev) for t in d['t_cls']]),
                           leaves['x'], [t.to(dev) for t in d['t_x']], d['img_metas'])
        sum(sum(v) for v in losses.values()).backward()
        torch.cuda.synchronize()
        tg = head._last_im_targets
        counts, rows = tg.counts.cpu(), tg.fg_rows.cpu()
        o = 0
        for l, want in enumerate(gold['gi_idx']):
            c = int(counts[4 + 3 * l + 2])
            r0 = sum(tg.pyr.level_anchors[:l]) * n
            assert c == want.numel() and torch.equal(rows[o:o + c] - r0, want), (name, l)
            o += c
        for k, v in gold['losses'].items():
            got = torch.stack([t.detach().cpu() for t in losses[k]])
            assert torch.allclose(got, v, rtol=1e-4, atol=1e-6), (name, k, got, v)

        def g(t):
            return torch.zeros_like(t) if t.grad is None else t.grad
        for l in range(5):
            for key, gk in (('score', 'grad_score'), ('reg', 'grad_reg')):
                got, ref = g(leaves[key][l]).cpu(), gold[gk][l]
                assert float((got - ref).abs().max()) <= 1e-4 * float(ref.abs().max()) + 1e-9, (name, key, l)
            # channel sums cancel (the KD term's gradient sums to zero over classes): the error scale of a row sum is
            # the sum of the magnitudes, not the sum itself
            for key, gk in (('cls_feat', 'grad_cls_feat_rowsum'), ('x', 'grad_x_rowsum')):
                got, ref = g(leaves[key][l]).cpu(), gold[gk][l]
                scale = max(float(ref.abs().max()), float(got.abs().sum(1).max()), 1e-6)
                assert float((got.sum(1) - ref).abs().max()) <= 2e-4 * scale, (name, key, l)


def test_ldv2_step_vs_golden(golden_dir):
    """Whole LDv2 iteration (configs/ldv2/ld_r50_gflv2_r101_fpn_1x.py with R18 backbones): LDv2Head student with the
    DGQP branch, GFocalHead teacher, gibox imitation -- bf16 convs against the fp32 reference (tolerances of
    tests/test_gpu_step.py)."""
    import os
    import ld_b200
    from ld_b200 import presets, synth
    from ld_b200.config import ConfigDict
    from oracle import cases
    name, sd, td, n, h, w, seed = cases.STEP_CASES_V2[0]
    gold = torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)
    cfg = ConfigDict(presets.ld_model(f'r{sd}', teacher_depth=td))
    v2 = dict(reg_topk=4, reg_channels=64, add_mean=True,
              loss_cls=dict(type='QualityFocalLoss', use_sigmoid=False, beta=2.0, loss_weight=1.0))
    cfg.bbox_head.update(type='LDv2Head', imitation_method='gibox', loss_im=dict(type='IMLoss', loss_weight=2), **v2)
    cfg.teacher_config.model.bbox_head.update(type='GFocalHead', **v2)
    m = ld_b200.build_detector(cfg)
    synth.seeded_init(m, seed)
    synth.seeded_init(m.teacher_model, seed + 1, prefix='teacher.')
    m = m.cuda().train()
    b = synth.make_batch(n, h, w, seed=1000 + seed)
    data = dict(img=b['img'].cuda(), img_metas=b['img_metas'], gt_bboxes=[x.cuda() for x in b['gt_bboxes']],
                gt_labels=[x.cuda() for x in b['gt_labels']])
    loss, log = m._parse_losses(m(**data))
    loss.backward()
    torch.cuda.synchronize()
    for k, v in gold['log_vars'].items():
        # loss_im rides on the 10 GI picks per level: with bf16 features a different pick changes it by O(1/10)
        tol = 0.15 if k in ('loss_im', 'loss') else 0.03
        assert abs(log[k] - v) <= tol * abs(v) + 2e-3, (k, log[k], v)
    named = dict(m.named_parameters())
    assert sorted(k for k, p in named.items() if p.grad is None) == sorted(gold['no_grad'])
    for k, g in gold['grads'].items():
        mine, ref = named[k].grad.float().cpu().flatten(), g.flatten()
        cos = float(torch.dot(mine, ref) / (mine.norm() * ref.norm() + 1e-30))
        assert cos > 0.98, (k, cos)


# ---------------------------------------------------------------------------------------------------------
# fused dual stem (ldb200_stem_tc, the default; LDB200_NO_FUSED_STEM=1 = im2col + GEMM route) against that route.
# Unlike the kernels above this one synchronises warps through mbarriers: a first-run bug could spin forever -- keep
# the `timeout` in front of pytest.
@pytest.mark.parametrize('shape', [(2, 128, 160), (1, 97, 131), (3, 64, 64)])
def test_fused_stem_matches_im2col_route(shape):
    import ld_b200
    from ld_b200 import ops, presets, synth
    from ld_b200.config import ConfigDict
    n, h, w = shape
    m = ld_b200.build_detector(ConfigDict(presets.ld_model('r18', teacher_depth=18)))
    synth.seeded_init(m, 5)
    synth.seeded_init(m.teacher_model, 6, prefix='teacher.')
    m = m.cuda().train()
    img = torch.randn(n, 3, h, w, device='cuda', generator=torch.Generator(device='cuda').manual_seed(h))
    with torch.no_grad():
        want_s = m.backbone.stem(img)                    # im2col + GEMM + pool
        want_t = m.teacher_model.backbone.stem(img)
        ws, afs = m.backbone.stem_params()
        wt, aft = m.teacher_model.backbone.stem_params()
        got_s, got_t = ops.stem_tc(img.contiguous(), torch.cat([ws, wt]).contiguous(), [afs, aft])
        one, = ops.stem_tc(img.contiguous(), wt.contiguous(), [aft])
        got_s, got_t, one = ops.maxpool3x3s2(got_s), ops.maxpool3x3s2(got_t), ops.maxpool3x3s2(one)
    torch.cuda.synchronize()
    for got, want in ((got_s, want_s), (got_t, want_t), (one, want_t)):
        assert got.shape == want.shape
        assert float((got.float() - want.float()).abs().max()) <= 1e-2 * float(want.float().abs().max())


def test_fused_stem_train_step_equals_default(monkeypatch):
    import ld_b200
    from ld_b200 import presets, synth
    from ld_b200.config import ConfigDict
    m = ld_b200.build_detector(ConfigDict(presets.ld_model('r18', teacher_depth=18)))
    synth.seeded_init(m, 3)
    synth.seeded_init(m.teacher_model, 4, prefix='teacher.')
    m = m.cuda().train()
    b = synth.make_batch(2, 128, 160, seed=77)
    data = dict(img=b['img'].cuda(), img_metas=b['img_metas'], gt_bboxes=[x.cuda() for x in b['gt_bboxes']],
                gt_labels=[x.cuda() for x in b['gt_labels']])
    _, log1 = m._parse_losses(m(**data))              # default: fused dual stem
    monkeypatch.setenv('LDB200_NO_FUSED_STEM', '1')
    _, log0 = m._parse_losses(m(**data))              # im2col + GEMM route
    for k, v in log0.items():
        assert abs(log1[k] - v) <= 2e-2 * abs(v) + 1e-3, (k, log1[k], v)


# ---------------------------------------------------------------------------------------------------------
def test_loss_curve_100_iterations(golden_dir):
    """SURVEY.md §8d / north_star: the student's loss curve over 100 synthetic iterations against the unmodified
    reference (tests/golden/c_*.pt: fp32 CPU, torch SGD + mmcv warm-up).  LDTrainStep computes the convolutions in
    bf16, so the bound here is the bf16 one (5 % per point, 2 % on the mean of the last 20 iterations); the measured
    maximum deviation is printed for DESIGN.md."""
    import os
    import ld_b200
    from ld_b200 import presets, synth
    from ld_b200.config import ConfigDict
    from oracle import cases
    name, sd, td, n, h, w, seed, iters, base_lr, warm = cases.CURVE_CASE
    gold = torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)
    m = ld_b200.build_detector(ConfigDict(presets.ld_model(f'r{sd}', teacher_depth=td)))
    synth.seeded_init(m, seed)
    synth.seeded_init(m.teacher_model, seed + 1, prefix='teacher.')
    m = m.cuda().train()
    ts = ld_b200.LDTrainStep(m, lr=base_lr, warmup_iters=warm, warmup_ratio=0.001)
    curve = []
    for it in range(iters):
        b = synth.make_batch(n, h, w, seed=2000 + it)
        data = dict(img=b['img'].cuda(), img_metas=b['img_metas'], gt_bboxes=[x.cuda() for x in b['gt_bboxes']],
                    gt_labels=[x.cuda() for x in b['gt_labels']])
        loss, _ = ts.step(data)
        curve.append(float(loss))
    got, want = torch.tensor(curve), gold['loss']
    rel = ((got - want).abs() / want.abs())
    print('loss-curve max rel deviation', float(rel.max()), 'mean', float(rel.mean()))
    assert float(rel.max()) <= 0.05, (float(rel.max()), int(rel.argmax()))
    assert abs(float(got[-20:].mean()) - float(want[-20:].mean())) <= 0.02 * float(want[-20:].mean())
