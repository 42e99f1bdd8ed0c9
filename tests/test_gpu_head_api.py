"""GPU: the reference-shaped public API -- `LDHead.loss(...)` with the reference's argument list
(ld_head.py:284-294) and the stand-alone `KnowledgeDistillationKLDivLoss` module (kd_loss.py:39-88) -- against
the golden vectors of the unmodified reference / the oracle (fp32, 1e-4)."""
import os

import pytest
import torch

from oracle import cases, ld_oracle as lo

pytestmark = pytest.mark.gpu


def make_head():
    import ld_b200
    from ld_b200 import presets
    from ld_b200.config import ConfigDict
    cfg = ConfigDict(presets.ld_model('r50'))
    head_cfg = cfg.bbox_head
    head_cfg['train_cfg'] = cfg.train_cfg
    head_cfg['test_cfg'] = cfg.test_cfg
    return ld_b200.build_head(head_cfg).cuda().train()


@pytest.mark.parametrize('case', cases.LOSS_CASES[:2], ids=lambda c: c[0])
def test_ldhead_loss_api_matches_reference(golden_dir, case):
    name, n, h, w, seed = case
    gold = torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)
    d = cases.loss_case_inputs(n, h, w, seed)
    head = make_head()
    dev = torch.device('cuda')
    leaves = {k: [t.to(dev).requires_grad_(True) for t in d[k]] for k in ('cls', 'reg', 'x')}
    losses = head.loss(leaves['cls'], leaves['reg'], [b.to(dev) for b in d['gt_bboxes']],
                       [b.to(dev) for b in d['gt_labels']],
                       ([t.to(dev) for t in d['t_cls']], [t.to(dev) for t in d['t_reg']]),
                       leaves['x'], [t.to(dev) for t in d['t_x']], d['img_metas'])
    assert list(losses) == ['loss_cls', 'loss_bbox', 'loss_dfl', 'loss_ld', 'loss_ld_vlr', 'loss_kd', 'loss_kd_neg',
                            'loss_im']
    total = sum(sum(v) for v in losses.values())
    total.backward()
    torch.cuda.synchronize()
    for k, v in gold['losses'].items():
        got = torch.stack([t.detach().cpu() for t in losses[k]])
        assert torch.allclose(got, v, rtol=1e-4, atol=1e-6), (k, got, v)
    for l in range(5):
        for key, gk in (('cls', 'grad_cls'), ('reg', 'grad_reg')):
            g, ref = leaves[key][l].grad.cpu(), gold[gk][l]
            assert float((g - ref).abs().max()) <= 1e-4 * float(ref.abs().max()) + 1e-9, (key, l)
        gx = leaves['x'][l].grad
        gx = torch.zeros_like(leaves['x'][l]) if gx is None else gx
        ref = gold['grad_x_rowsum'][l]
        assert float((gx.cpu().sum(1) - ref).abs().max()) <= 2e-4 * max(float(ref.abs().max()), 1e-6), l


@pytest.mark.parametrize('case', cases.LOSS_CASES, ids=lambda c: c[0] + '_gibox')
def test_ldhead_loss_gibox_matches_reference(golden_dir, case):
    """imitation_method='gibox' (the LDHead default; ld_head.py:170-176, 613-638): GI-box NMS picks bit-exact
    (`ldb200_gi_region`), loss terms and gradients against the unmodified reference."""
    name, n, h, w, seed = case
    gold = torch.load(os.path.join(golden_dir, name + '_gibox.pt'), weights_only=False)
    d = cases.loss_case_inputs(n, h, w, seed)
    head = make_head()
    head.imitation_method = 'gibox'
    dev = torch.device('cuda')
    leaves = {k: [t.to(dev).requires_grad_(True) for t in d[k]] for k in ('cls', 'reg', 'x')}
    losses = head.loss(leaves['cls'], leaves['reg'], [b.to(dev) for b in d['gt_bboxes']],
                       [b.to(dev) for b in d['gt_labels']],
                       ([t.to(dev) for t in d['t_cls']], [t.to(dev) for t in d['t_reg']]),
                       leaves['x'], [t.to(dev) for t in d['t_x']], d['img_metas'])
    sum(sum(v) for v in losses.values()).backward()
    torch.cuda.synchronize()
    # picked rows: row space = levels back to back, the reference's indices are level-local
    tg = head._last_im_targets
    counts = tg.counts.cpu()
    rows = tg.fg_rows.cpu()[:int(counts[1])]
    o = 0
    for l, want in enumerate(gold['gi_idx']):
        c = int(counts[4 + 3 * l + 2])
        assert c == want.numel(), (l, c, want.numel())
        r0 = sum(tg.pyr.level_anchors[:l]) * n
        assert torch.equal(rows[o:o + c] - r0, want), (l, rows[o:o + c] - r0, want)
        o += c
    for k, v in gold['losses'].items():
        got = torch.stack([t.detach().cpu() for t in losses[k]])
        assert torch.allclose(got, v, rtol=1e-4, atol=1e-6), (k, got, v)
    for l in range(5):
        gx = leaves['x'][l].grad
        gx = torch.zeros_like(leaves['x'][l]) if gx is None else gx
        ref = gold['grad_x_rowsum'][l]
        assert float((gx.cpu().sum(1) - ref).abs().max()) <= 2e-4 * max(float(ref.abs().max()), 1e-6), l
        for key, gk in (('cls', 'grad_cls_norm'), ('reg', 'grad_reg_norm')):
            assert abs(float(leaves[key][l].grad.norm()) - float(gold[gk][l])) <= 1e-4 * float(gold[gk][l]) + 1e-9


def test_kd_kl_module_standalone():
    import ld_b200
    g = torch.Generator().manual_seed(0)
    pred = (torch.randn(4000, 17, generator=g) * 2).requires_grad_(True)
    soft = torch.randn(4000, 17, generator=g) * 2
    w = torch.rand(4000, generator=g)
    ref = 0.25 * (lo.kd_kl_rows(pred, soft, 10.0) * w).sum() / 4.0
    ref.backward()
    mod = ld_b200.build_loss(dict(type='KnowledgeDistillationKLDivLoss', loss_weight=0.25, T=10))
    p2 = pred.detach().cuda().requires_grad_(True)
    out = mod(p2, soft.cuda(), weight=w.cuda(), avg_factor=4.0)
    out.backward()
    torch.cuda.synchronize()
    assert abs(float(out) - float(ref)) <= 1e-4 * abs(float(ref))
    assert float((p2.grad.cpu() - pred.grad).abs().max()) <= 1e-4 * float(pred.grad.abs().max())
    with pytest.raises(ValueError):
        mod(p2, soft.cuda(), avg_factor=4.0, reduction_override='sum')
    with pytest.raises(TypeError):
        ld_b200.build_loss(dict(type='KnowledgeDistillationKLDivLoss', T=10, bogus=1))


def test_imloss_module_standalone():
    import ld_b200
    g = torch.Generator().manual_seed(1)
    x = torch.randn(300, 256, generator=g).requires_grad_(True)
    t = torch.randn(300, 256, generator=g)
    ref = 2.0 * torch.nn.functional.mse_loss(x, t)
    ref.backward()
    mod = ld_b200.build_loss(dict(type='IMLoss', loss_weight=2.0))
    xg = x.detach().cuda().requires_grad_(True)
    out = mod(xg, t.cuda())
    out.backward()
    torch.cuda.synchronize()
    assert abs(float(out) - float(ref)) <= 1e-5 * abs(float(ref))
    assert float((xg.grad.cpu() - x.grad).abs().max()) <= 1e-5 * float(x.grad.abs().max())


def test_ldhead_loss_honours_upstream_gradients():
    """the loss node precomputes its input gradients for the plain sum of all terms; a scaled loss, a
    re-weighted sum or a subset of the returned terms must still give the right gradients (oracle autograd)."""
    n, h, w, seed = 2, 192, 256, 17
    d = cases.loss_case_inputs(n, h, w, seed)
    dev = torch.device('cuda')
    tg = lo.ld_targets(d['sizes'], cases.STRIDES, d['img_metas'], d['gt_bboxes'], d['gt_labels'])

    def weights(losses):     # 3 x (cls + kd of level 0 only) + 0.5 x bbox, everything else dropped
        return 3.0 * (losses['loss_cls'][0] + losses['loss_kd'][0]) + 0.5 * sum(losses['loss_bbox'])

    ref = {k: [t.clone().requires_grad_(True) for t in d[k]] for k in ('cls', 'reg', 'x')}
    weights(lo.ld_loss(ref['cls'], ref['reg'], d['t_cls'], d['t_reg'], ref['x'], d['t_x'], tg, cases.STRIDES)).backward()
    head = make_head()
    leaves = {k: [t.to(dev).requires_grad_(True) for t in d[k]] for k in ('cls', 'reg', 'x')}
    losses = head.loss(leaves['cls'], leaves['reg'], [b.to(dev) for b in d['gt_bboxes']], [b.to(dev) for b in d['gt_labels']],
                       ([t.to(dev) for t in d['t_cls']], [t.to(dev) for t in d['t_reg']]),
                       leaves['x'], [t.to(dev) for t in d['t_x']], d['img_metas'])
    weights(losses).backward()
    torch.cuda.synchronize()
    for key in ('cls', 'reg'):
        for l in range(5):
            want = ref[key][l].grad
            want = torch.zeros_like(ref[key][l]) if want is None else want
            got = leaves[key][l].grad.cpu()
            assert float((got - want).abs().max()) <= 1e-4 * float(want.abs().max()) + 1e-9, (key, l)
    for l in range(5):       # the imitation term was not part of the sum: no gradient reaches x
        gx = leaves['x'][l].grad
        assert gx is None or float(gx.abs().max()) == 0.0
