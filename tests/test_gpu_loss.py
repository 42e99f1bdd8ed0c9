"""GPU: fused LDHead loss (C-ABI) vs golden vectors of the unmodified reference and the CPU oracle.
Tolerance: 1e-4 relative on every loss term and gradient (BASELINE.json north_star), fp32."""
import os

import pytest
import torch

from oracle import cases, ld_oracle as lo

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def run_fused(d, n, h, w, grad_dtype=torch.float32, feat_dtype=torch.float32, cfg_kw=None):
    from ld_b200 import lib as L, ops
    dev = torch.device('cuda')
    pyr = L.Pyramid.make(n, d['sizes'], cases.STRIDES)
    tg = ops.ld_targets(pyr, d['gt_bboxes'], d['gt_labels'], [(h, w)] * n, device=dev)
    rows = {k: ops.to_rows([t.to(dev) for t in d[k]]) for k in ('cls', 'reg', 't_cls', 't_reg', 'x', 't_x')}
    cfg = ops.make_loss_cfg(**(cfg_kw or {}))
    wsum, score = ops.ld_loss_prepass(cfg, rows['cls'], rows['reg'], tg, want_score=True)
    norm = ops.loss_normalisers(tg, wsum)
    losses, d_cls, d_reg, d_im = ops.ld_loss_fused(
        cfg, rows['cls'], rows['reg'], rows['t_cls'], rows['t_reg'], rows['x'].to(feat_dtype),
        rows['t_x'].to(feat_dtype), tg, norm, term_grad=True, grad_dtype=grad_dtype)
    d_x = torch.zeros(pyr.rows, 256, dtype=grad_dtype, device=dev)
    ops.scatter_add_rows(d_x, d_im, tg.fg_rows, tg.counts[1:2])
    torch.cuda.synchronize()
    return pyr, tg, losses.cpu(), d_cls, d_reg, d_x, score


def close(a, b, rtol=RTOL, what=''):
    scale = float(b.abs().max()) if b.numel() else 0.0
    err = float((a - b).abs().max()) if b.numel() else 0.0
    assert err <= rtol * max(scale, 1e-6), f'{what}: max err {err:.3e} vs scale {scale:.3e}'


@pytest.mark.parametrize('case', cases.LOSS_CASES, ids=lambda c: c[0])
def test_loss_golden(golden_dir, case):
    from ld_b200 import ops
    name, n, h, w, seed = case
    gold = torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)
    d = cases.loss_case_inputs(n, h, w, seed)
    pyr, tg, losses, d_cls, d_reg, d_x, _ = run_fused(d, n, h, w)
    for ti, k in enumerate(ops.TERMS):
        got, want = losses[ti, :5], gold['losses'][k]
        assert torch.allclose(got, want, rtol=RTOL, atol=1e-6), (k, got, want)
    gc = ops.from_rows(d_cls, pyr)
    gr = ops.from_rows(d_reg, pyr)
    gx = ops.from_rows(d_x, pyr)
    for l in range(5):
        close(gc[l].cpu(), gold['grad_cls'][l], what=f'd_cls L{l}')
        close(gr[l].cpu(), gold['grad_reg'][l], what=f'd_reg L{l}')
        close(gx[l].cpu().sum(1), gold['grad_x_rowsum'][l], rtol=2e-4, what=f'd_x L{l}')


@pytest.mark.parametrize('seed', [31, 32, 33])
def test_loss_vs_oracle(seed):
    """Live oracle comparison incl. full d_x, the score map and non-default loss weights."""
    from ld_b200 import ops
    n, h, w = 2, 320, 384
    d = cases.loss_case_inputs(n, h, w, seed)
    kw = dict(ld_T=5.0, ld_vlr_T=8.0, kd_T=3.0, im_weight=1.5, giou_weight=1.7) if seed == 33 else {}
    pyr, tg, losses, d_cls, d_reg, d_x, score = run_fused(d, n, h, w, cfg_kw=kw)
    leaves = {k: [t.clone().requires_grad_(True) for t in d[k]] for k in ('cls', 'reg', 'x')}
    otg = lo.ld_targets(d['sizes'], cases.STRIDES, d['img_metas'], d['gt_bboxes'], d['gt_labels'])
    ocfg = dict(lo.DEFAULT_LOSS_CFG)
    if kw:
        ocfg.update(T_ld=5.0, T_ld_vlr=8.0, T_kd=3.0, w_im=1.5, w_giou=1.7)
    ol = lo.ld_loss(leaves['cls'], leaves['reg'], d['t_cls'], d['t_reg'], leaves['x'], d['t_x'], otg,
                    cases.STRIDES, ocfg)
    sum(sum(v) for v in ol.values()).backward()
    for ti, k in enumerate(ops.TERMS):
        want = torch.stack([t.detach() for t in ol[k]])
        assert torch.allclose(losses[ti, :5], want, rtol=RTOL, atol=1e-6), (k, losses[ti, :5], want)
    gc = ops.from_rows(d_cls, pyr)
    gr = ops.from_rows(d_reg, pyr)
    gx = ops.from_rows(d_x, pyr)
    for l in range(5):
        close(gc[l].cpu(), leaves['cls'][l].grad, what=f'd_cls L{l}')
        close(gr[l].cpu(), leaves['reg'][l].grad, what=f'd_reg L{l}')
        g = leaves['x'][l].grad
        close(gx[l].cpu(), torch.zeros_like(leaves['x'][l]) if g is None else g, what=f'd_x L{l}')


def test_loss_bf16_outputs_and_no_grad_mode():
    """bf16 gradient / feature dtypes (the training configuration) agree with fp32 to bf16 rounding."""
    n, h, w = 2, 256, 320
    d = cases.loss_case_inputs(n, h, w, 41)
    pyr, tg, l32, c32, r32, x32, _ = run_fused(d, n, h, w)
    _, _, l16, c16, r16, x16, _ = run_fused(d, n, h, w, grad_dtype=torch.bfloat16)
    assert torch.allclose(l32, l16, rtol=1e-6, atol=1e-7)
    close(c16.float().cpu(), c32.cpu(), rtol=8e-3, what='bf16 d_cls')
    close(r16.float().cpu(), r32.cpu(), rtol=8e-3, what='bf16 d_reg')
    _, _, lb, _, _, xb, _ = run_fused(d, n, h, w, feat_dtype=torch.bfloat16)
    assert torch.allclose(lb[:7], l32[:7], rtol=1e-6, atol=1e-7)
    assert torch.allclose(lb[7], l32[7], rtol=2e-2, atol=1e-6)


def test_loss_linearity_full_size():
    """BASELINE size (16 x 800x1344 -> 358 400 rows): the gradient is linear in the upstream term
    weights and the losses do not depend on them (size-independent property)."""
    from ld_b200 import lib as L, ops
    from ld_b200 import synth
    dev = torch.device('cuda')
    n, h, w = 16, 800, 1344
    sizes = cases.featmap_sizes(h, w)
    pyr = L.Pyramid.make(n, sizes, cases.STRIDES)
    gb, gl = synth.make_gts(n, h, w, 9)
    tg = ops.ld_targets(pyr, gb, gl, [(h, w)] * n, device=dev)
    g = torch.Generator(device='cuda').manual_seed(3)
    R = pyr.rows
    cls = torch.randn(R, 80, device=dev, generator=g) * 1.5 - 2
    reg = torch.randn(R, 68, device=dev, generator=g)
    t_cls = torch.randn(R, 80, device=dev, generator=g) * 1.5 - 2
    t_reg = torch.randn(R, 68, device=dev, generator=g) * 1.5
    x = torch.randn(R, 256, device=dev, generator=g).bfloat16()
    t_x = torch.randn(R, 256, device=dev, generator=g).bfloat16()
    cfg = ops.make_loss_cfg()
    wsum, _ = ops.ld_loss_prepass(cfg, cls, reg, tg)
    norm = ops.loss_normalisers(tg, wsum)
    tg1 = torch.ones(8, 8)
    l1, c1, r1, i1 = ops.ld_loss_fused(cfg, cls, reg, t_cls, t_reg, x, t_x, tg, norm, term_grad=tg1)
    l2, c2, r2, i2 = ops.ld_loss_fused(cfg, cls, reg, t_cls, t_reg, x, t_x, tg, norm, term_grad=2 * tg1)
    torch.cuda.synchronize()
    assert torch.isfinite(l1).all() and torch.allclose(l1, l2, rtol=1e-5, atol=1e-7)
    assert torch.allclose(2 * c1, c2, rtol=1e-5, atol=1e-12) and torch.allclose(2 * r1, r2, rtol=1e-5, atol=1e-12)
    nfg = int(tg.counts[1])
    assert torch.allclose(2 * i1[:nfg], i2[:nfg], rtol=1e-5, atol=1e-12)
    # rows that are neither positive nor VLR get exactly zero box gradient
    dead = ~((tg.labels < 80) | (tg.vlr > 0))
    assert (r1[dead] == 0).all()
    assert abs(float(l1[8].sum()) - float(wsum)) <= 1e-4 * float(wsum)
