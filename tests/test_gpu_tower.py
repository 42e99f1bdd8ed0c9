"""GPU: one GFL tower layer (gfl_head.py:102-133, 177-180: level-shared 3x3 conv -> GroupNorm(32) -> ReLU) as the fused
multi-level node `engine.conv_gn_relu_levels`: GroupNorm statistics from the conv epilogue (`ldb_conv_desc.gn_stats`),
one normalise launch and two backward launches for all pyramid levels -- against torch autograd of the same layer,
in the bf16 mode (bf16 tolerances) and in the fp32-accurate mode (1e-4)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SIZES = [(25, 42), (13, 21), (7, 11), (4, 6), (2, 3)]


def test_conv_epilogue_gn_statistics():
    """Sum / sum of squares per (image, group of 8 channels) accumulated by the conv epilogue == reductions of the
    fp32 conv result (1-CTA and 2-CTA kernels, several levels in one launch, odd sizes -> partial tiles)."""
    from ld_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(11)
    for cout in (256, 128):
        wt = torch.randn(cout, 256, 3, 3, device='cuda', generator=g) * 0.03
        xs = [torch.randn(3, h, w, 256, device='cuda', generator=g).bfloat16() for h, w in SIZES[:4]]
        stats = [torch.zeros(3, cout // 8, 2, device='cuda') for _ in xs]
        ys = ops.conv2d_tc_multi(xs, ops.pack_conv_weight(wt), cout, 3, 3, 1, 1, gn_stats=stats)
        torch.cuda.synchronize()
        for x, y, st in zip(xs, ys, stats):
            ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.bfloat16().float(), None, 1, 1)       # [N, C, H, W]
            r = ref.reshape(3, cout // 8, -1)
            want = torch.stack([r.sum(-1), (r * r).sum(-1)], -1)
            scale = float((r * r).sum(-1).max())
            assert float((st - want).abs().max()) <= 2e-4 * scale, (cout, tuple(x.shape))
            assert float((y.float() - ref.permute(0, 2, 3, 1)).abs().max()) <= 8e-3 * float(ref.abs().max())


def _ref_layer(xs, wt, gamma, beta):
    outs = []
    for x in xs:
        c = F.conv2d(x.permute(0, 3, 1, 2), wt, None, 1, 1)
        outs.append(F.relu(F.group_norm(c, 32, gamma, beta, 1e-5)).permute(0, 2, 3, 1))
    return outs


@pytest.mark.parametrize('mode', ['bf16', 'fp32'])
def test_tower_layer_forward_backward(mode):
    from ld_b200 import engine
    g = torch.Generator(device='cuda').manual_seed(3)
    n = 2
    wt = (torch.randn(256, 256, 3, 3, device='cuda', generator=g) * 0.02)
    gamma = torch.rand(256, device='cuda', generator=g) + 0.5
    beta = torch.randn(256, device='cuda', generator=g) * 0.2
    xs32 = [torch.randn(n, h, w, 256, device='cuda', generator=g) for h, w in SIZES]
    dys32 = [torch.randn(n, h, w, 256, device='cuda', generator=g) for h, w in SIZES]
    dt = torch.bfloat16 if mode == 'bf16' else torch.float32
    # reference: torch autograd (fp64 in the fp32-accurate mode, fp32 on bf16-rounded operands otherwise)
    rt = torch.float64 if mode == 'fp32' else torch.float32
    rw = (wt if mode == 'fp32' else wt.bfloat16()).to(rt).requires_grad_(True)
    rg, rb = gamma.detach().clone().to(rt).requires_grad_(True), beta.detach().clone().to(rt).requires_grad_(True)
    rxs = [x.detach().clone().to(dt).to(rt).requires_grad_(True) for x in xs32]
    routs = _ref_layer(rxs, rw, rg, rb)
    sum((o * d.to(dt).to(rt)).sum() for o, d in zip(routs, dys32)).backward()

    w = wt.clone().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    ga, be = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    xs = [x.detach().clone().to(dt).requires_grad_(True) for x in xs32]
    with engine.precision(mode):
        ys = engine.conv_gn_relu_levels(w, ga, be, xs)
        torch.autograd.backward(ys, [d.to(dt) for d in dys32])
    torch.cuda.synchronize()
    tol_y, tol_g = (2e-2, 3e-2) if mode == 'bf16' else (1e-5, 1e-4)

    def rel(a, b):
        # fp32 mode: worst element relative to the largest; bf16 mode: rel-L2 (the bf16 rounding of the conv output
        # before the normalisation moves single elements -- and their ReLU masks -- by up to a bf16 ulp)
        a, b = a.double(), b.double()
        if mode == 'bf16':
            return float((a - b).norm() / (b.norm() + 1e-30))
        return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)
    for y, r in zip(ys, routs):
        assert rel(y, r) <= tol_y, ('y', rel(y, r))
    for x, r in zip(xs, rxs):
        assert rel(x.grad, r.grad) <= tol_g, ('dx', rel(x.grad, r.grad))
    assert rel(w.grad, rw.grad) <= tol_g, ('dw', rel(w.grad, rw.grad))
    assert rel(ga.grad, rg.grad) <= tol_g and rel(be.grad, rb.grad) <= tol_g, (rel(ga.grad, rg.grad), rel(be.grad, rb.grad))


def test_fused_tower_equals_per_level_path(monkeypatch):
    """The fused node against the per-level GroupNorm path it replaces (same kernels otherwise): head forward identical
    up to the rounding of the statistics, on a whole pyramid."""
    import ld_b200
    from ld_b200 import presets
    from ld_b200.config import ConfigDict
    cfg = ConfigDict(presets.ld_model('r50'))
    hc = cfg.bbox_head
    hc['train_cfg'], hc['test_cfg'] = cfg.train_cfg, cfg.test_cfg
    head = ld_b200.build_head(hc).cuda().train()
    g = torch.Generator(device='cuda').manual_seed(9)
    feats = [torch.randn(2, 256, h, w, device='cuda', generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
             for h, w in SIZES]
    with torch.no_grad():
        a_cls, a_reg = head(feats)
        monkeypatch.setenv('LDB200_GN_PER_LEVEL', '1')
        b_cls, b_reg = head(feats)
    torch.cuda.synchronize()
    for a, b in zip(a_cls + a_reg, b_cls + b_reg):
        assert float((a - b).abs().max()) <= 2e-2 * float(b.abs().max()) + 1e-3
