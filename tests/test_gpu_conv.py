"""GPU: tcgen05 implicit-GEMM convolution (C-ABI ldb200_conv2d_tc) vs torch conv2d in fp32 on the
same bf16-rounded operands.  Tolerance 1e-4 relative (fp32 accumulation order only) for fp32 outputs,
bf16 rounding (4e-3) for bf16 outputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def ref_conv(x_nhwc, w, stride, pad, scale=None, bias=None, add=None, relu=False):
    x = x_nhwc.float().permute(0, 3, 1, 2)
    y = F.conv2d(x, w.to(torch.bfloat16).float(), None, stride, pad)
    if scale is not None:
        y = y * scale.view(1, -1, 1, 1)
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    if add is not None:
        a = add.float().permute(0, 3, 1, 2)
        if a.shape[2:] != y.shape[2:]:
            a = F.interpolate(a, size=y.shape[2:], mode='nearest')
        y = y + a
    if relu:
        y = y.relu()
    return y.permute(0, 2, 3, 1).contiguous()


def check(got, want, rtol):
    err = float((got.float() - want).abs().max())
    scale = float(want.abs().max())
    assert err <= rtol * scale, f'max err {err:.3e}, scale {scale:.3e}'


CASES = [
    # n, h, w, cin, cout, k, stride, pad
    (2, 16, 24, 64, 64, 3, 1, 1),
    (1, 25, 42, 256, 256, 3, 1, 1),     # head tower shape (P5-ish, odd sizes)
    (2, 13, 21, 256, 80, 3, 1, 1),      # gfl_cls (Cout not a multiple of 32)
    (2, 7, 11, 256, 68, 3, 1, 1),       # gfl_reg
    (2, 20, 28, 128, 512, 1, 1, 0),     # bottleneck 1x1 (flat path)
    (1, 30, 44, 512, 128, 1, 1, 0),
    (2, 33, 47, 128, 128, 3, 2, 1),     # strided 3x3 (layer2.0.conv2) odd input
    (2, 32, 48, 256, 512, 1, 2, 0),     # strided 1x1 downsample
    (1, 13, 21, 256, 256, 3, 2, 1),     # FPN extra conv P6
    (1, 100, 168, 64, 64, 3, 1, 1),     # big map, many tiles per CTA
]


@pytest.mark.parametrize('case', CASES, ids=lambda c: 'x'.join(map(str, c)))
@pytest.mark.parametrize('out_dtype', [torch.float32, torch.bfloat16])
def test_conv_fprop(case, out_dtype):
    from ld_b200 import ops
    n, h, w, cin, cout, k, stride, pad = case
    g = torch.Generator(device='cuda').manual_seed(sum(case))
    x = torch.randn(n, h, w, cin, device='cuda', generator=g).bfloat16()
    wt = torch.randn(cout, cin, k, k, device='cuda', generator=g) * (2.0 / (cin * k * k)) ** 0.5
    y = ops.conv2d_tc(x, ops.pack_conv_weight(wt), cout, k, k, stride, pad, out_dtype=out_dtype)
    torch.cuda.synchronize()
    want = ref_conv(x, wt, stride, pad)
    assert y.shape == want.shape
    check(y, want, 1e-4 if out_dtype == torch.float32 else 5e-3)


def test_conv_epilogue_fusions():
    from ld_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(7)
    n, h, w, cin, cout = 2, 26, 38, 128, 256
    x = torch.randn(n, h, w, cin, device='cuda', generator=g).bfloat16()
    wt = torch.randn(cout, cin, 3, 3, device='cuda', generator=g) * 0.03
    scale = torch.rand(cout, device='cuda', generator=g) + 0.5
    bias = torch.randn(cout, device='cuda', generator=g)
    res = torch.randn(n, h, w, cout, device='cuda', generator=g).bfloat16()
    wp = ops.pack_conv_weight(wt)
    y = ops.conv2d_tc(x, wp, cout, 3, 3, 1, 1, out_dtype=torch.float32, scale=scale, bias=bias, add=res, relu=True)
    check(y, ref_conv(x, wt, 1, 1, scale, bias, res, True), 1e-4)
    # nearest-upsample add (FPN top-down, fpn.py:182-191) from a coarser map with odd size ratio
    coarse = torch.randn(n, 13, 19, cout, device='cuda', generator=g).bfloat16()
    y = ops.conv2d_tc(x, wp, cout, 3, 3, 1, 1, out_dtype=torch.float32, bias=bias, add=coarse)
    check(y, ref_conv(x, wt, 1, 1, None, bias, coarse), 1e-4)
    # accumulate into an existing output, and ReLU-mask
    base = torch.randn(n, h, w, cout, device='cuda', generator=g)
    out = base.clone()
    ops.conv2d_tc(x, wp, cout, 3, 3, 1, 1, out=out, accumulate=True)
    check(out, ref_conv(x, wt, 1, 1) + base, 1e-4)
    mask = torch.randn(n, h, w, cout, device='cuda', generator=g).bfloat16()
    y = ops.conv2d_tc(x, wp, cout, 3, 3, 1, 1, out_dtype=torch.float32, mask=mask)
    check(y, ref_conv(x, wt, 1, 1) * (mask.float() > 0), 1e-4)
    torch.cuda.synchronize()


TMA_EPI_CASES = [
    # n, h, w, cin, cout, k, pad  -- bf16 outputs with Cout % 64 == 0 leave through the TMA-staged epilogue
    (2, 20, 28, 128, 512, 1, 0),    # flat 1x1, short reduction (2 CTAs / SM kernel), 128-pixel rows
    (2, 20, 28, 512, 256, 1, 0),    # flat 1x1, 2-CTA kernel
    (1, 25, 42, 256, 256, 3, 1),    # 3x3, 2-CTA kernel, odd number of pixel tiles (dummy pair tile)
    (3, 7, 6, 256, 256, 3, 1),      # narrow map: 8-wide tiles, 8 x 4 pixel boxes, clipped on both edges
    (2, 9, 70, 128, 128, 3, 1),     # wide map: 32 x 1 boxes
    (2, 13, 21, 320, 192, 3, 1),    # Cin, Cout multiples of 64 only (BN = 256 tile half empty)
]


@pytest.mark.parametrize('case', TMA_EPI_CASES, ids=lambda c: 'x'.join(map(str, c)))
def test_conv_tma_epilogue_fusions(case):
    """bf16 output + folded-BN scale/bias + residual + ReLU, and + ReLU-mask, through the TMA-staged epilogue
    (residual / mask boxes TMA-loaded, result TMA-stored) against the fp32 reference."""
    from ld_b200 import ops
    n, h, w, cin, cout, k, pad = case
    g = torch.Generator(device='cuda').manual_seed(sum(case))
    x = torch.randn(n, h, w, cin, device='cuda', generator=g).bfloat16()
    wt = torch.randn(cout, cin, k, k, device='cuda', generator=g) * (2.0 / (cin * k * k)) ** 0.5
    wp = ops.pack_conv_weight(wt)
    scale = torch.rand(cout, device='cuda', generator=g) + 0.5
    bias = torch.randn(cout, device='cuda', generator=g)
    res = torch.randn(n, h, w, cout, device='cuda', generator=g).bfloat16()
    mask = torch.randn(n, h, w, cout, device='cuda', generator=g).bfloat16()
    mask[..., ::7] = 0       # exact zeros and negative zeros are "not > 0"
    mask[..., 3::11] = -0.0
    y = ops.conv2d_tc(x, wp, cout, k, k, 1, pad, scale=scale, bias=bias, add=res, relu=True)
    check(y, ref_conv(x, wt, 1, pad, scale, bias, res, True), 5e-3)
    y = ops.conv2d_tc(x, wp, cout, k, k, 1, pad, scale=scale, add=res, mask=mask)
    check(y, ref_conv(x, wt, 1, pad, scale, None, res) * (mask.float() > 0), 5e-3)
    y = ops.conv2d_tc(x, wp, cout, k, k, 1, pad, mask=mask)
    want = ref_conv(x, wt, 1, pad) * (mask.float() > 0)
    check(y, want, 5e-3)
    # the remaining compile-time feature sets of the lean epilogue (conv_tc.cu EF_*): BN + ReLU, BN alone, and the
    # dgrad form residual + ReLU-mask
    y = ops.conv2d_tc(x, wp, cout, k, k, 1, pad, scale=scale, bias=bias, relu=True)
    check(y, ref_conv(x, wt, 1, pad, scale, bias, None, True), 5e-3)
    y = ops.conv2d_tc(x, wp, cout, k, k, 1, pad, scale=scale, bias=bias)
    check(y, ref_conv(x, wt, 1, pad, scale, bias), 5e-3)
    y = ops.conv2d_tc(x, wp, cout, k, k, 1, pad, add=res, mask=mask)
    check(y, ref_conv(x, wt, 1, pad, None, None, res) * (mask.float() > 0), 5e-3)
    # output written into a channel slice of a wider buffer (row-space buffers): neighbours stay untouched
    wide = torch.full((n, h, w, cout + 64), 7.0, device='cuda', dtype=torch.bfloat16)
    ops.conv2d_tc(x, wp, cout, k, k, 1, pad, out=wide[..., :cout], bias=bias)
    check(wide[..., :cout], ref_conv(x, wt, 1, pad, None, bias), 5e-3)
    assert bool((wide[..., cout:] == 7.0).all())
    torch.cuda.synchronize()


def test_conv_dgrad_matches_autograd():
    """Data gradient of a stride-1 3x3 and a 1x1 conv through the same kernel with rotated weights."""
    from ld_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(11)
    n, h, w, cin, cout = 2, 19, 27, 128, 192
    for k, pad in ((3, 1), (1, 0)):
        wt = (torch.randn(cout, cin, k, k, device='cuda', generator=g) * 0.05).bfloat16().float()
        dy = torch.randn(n, h, w, cout, device='cuda', generator=g).bfloat16()
        x = torch.zeros(n, cin, h, w, device='cuda', requires_grad=True)
        F.conv2d(x, wt, None, 1, pad).backward(dy.float().permute(0, 3, 1, 2))
        want = x.grad.permute(0, 2, 3, 1)
        wd = ops.pack_conv_weight_dgrad(wt)   # [Cin, k*k*Cout]
        taps = [(pad - s, pad - r, (r * k + s) * cout) for r in range(k) for s in range(k)]
        dx = ops.conv2d_tc(dy, wd, cin, k, k, 1, pad, out_dtype=torch.float32, taps=taps, flat=False if k == 3 else None)
        torch.cuda.synchronize()
        check(dx, want, 1e-4)


WG_CASES = [
    # n, h, w, cin, cout, k, stride, pad
    (2, 16, 24, 64, 64, 3, 1, 1),
    (2, 25, 42, 256, 256, 3, 1, 1),
    (2, 13, 21, 256, 80, 3, 1, 1),      # gfl_cls: Cout < 128 (second 64-channel box partly out of range)
    (1, 20, 28, 128, 512, 1, 1, 0),
    (2, 33, 47, 128, 128, 3, 2, 1),
    (2, 32, 48, 256, 512, 1, 2, 0),
    (1, 50, 84, 512, 128, 1, 1, 0),
    (2, 20, 28, 512, 256, 1, 1, 0),     # 2-CTA pair tile, two input-channel tiles
    (3, 9, 11, 256, 768, 3, 1, 1),      # 2-CTA, three output-channel pair tiles, tiny maps (more splits than tiles)
]


@pytest.mark.parametrize('case', WG_CASES, ids=lambda c: 'x'.join(map(str, c)))
def test_conv_wgrad(case):
    from ld_b200 import ops
    n, h, w, cin, cout, k, stride, pad = case
    g = torch.Generator(device='cuda').manual_seed(sum(case) + 1)
    x = torch.randn(n, h, w, cin, device='cuda', generator=g).bfloat16()
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    dy = torch.randn(n, ho, wo, cout, device='cuda', generator=g).bfloat16()
    wt = torch.zeros(cout, cin, k, k, device='cuda', requires_grad=True)
    F.conv2d(x.float().permute(0, 3, 1, 2), wt, None, stride, pad).backward(dy.float().permute(0, 3, 1, 2))
    dw = torch.zeros(cout, k * k * cin, device='cuda')
    ops.conv2d_wgrad_tc(dy, x, dw, k, k, stride, pad)
    ops.conv2d_wgrad_tc(dy, x, dw, k, k, stride, pad)      # accumulates
    torch.cuda.synchronize()
    got = ops.unpack_conv_weight_grad(dw, cout, cin, k, k)
    check(got, 2 * wt.grad, 1e-4)


def test_glue_kernels():
    from ld_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(5)
    # stem im2col + flat GEMM == 7x7/2 conv
    img = torch.randn(2, 3, 70, 94, device='cuda', generator=g)
    wt = torch.randn(64, 3, 7, 7, device='cuda', generator=g) * 0.1
    a, ho, wo = ops.stem_im2col(img)
    wp = torch.zeros(64, 192, device='cuda', dtype=torch.bfloat16)
    wp[:, :147] = wt.permute(0, 2, 3, 1).reshape(64, 147).bfloat16()
    y = ops.conv2d_tc(a.view(1, 1, -1, 192), wp, 64, out_dtype=torch.float32, relu=True).view(2, ho, wo, 64)
    want = F.conv2d(img.bfloat16().float(), wt.bfloat16().float(), None, 2, 3).relu().permute(0, 2, 3, 1)
    check(y, want, 1e-4)
    # max pool
    yb = y.bfloat16()
    mp = ops.maxpool3x3s2(yb)
    want = F.max_pool2d(yb.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(mp.float(), want)
    # group norm + relu forward / backward vs torch
    x = (torch.randn(2, 13, 21, 256, device='cuda', generator=g) * 2 + 0.5).bfloat16()
    gamma = torch.rand(256, device='cuda', generator=g) + 0.5
    beta = torch.randn(256, device='cuda', generator=g) * 0.2
    yq, stats = ops.gn_relu_fwd(x, gamma, beta)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.relu(F.group_norm(xr, 32, gr, br, 1e-5))
    check(yq, yr.detach().permute(0, 2, 3, 1), 6e-3)
    dy = torch.randn(2, 13, 21, 256, device='cuda', generator=g).bfloat16()
    yr.backward(dy.float().permute(0, 3, 1, 2) * (yq.float().permute(0, 3, 1, 2) > 0) / (yr.detach() > 0).float().clamp(min=1e-9) * (yr.detach() > 0))
    dgam = torch.zeros(256, device='cuda')
    dbet = torch.zeros(256, device='cuda')
    dx = ops.gn_relu_bwd(dy, yq, x, stats, gamma, dgam, dbet)
    check(dx, xr.grad.permute(0, 2, 3, 1), 1.5e-2)
    check(dgam, gr.grad, 2e-3)
    check(dbet, br.grad, 2e-3)
    # upsample-add backward (13x21 -> 25x42 nearest)
    dfine = torch.randn(2, 25, 42, 64, device='cuda', generator=g).bfloat16()
    dco = torch.randn(2, 13, 21, 64, device='cuda', generator=g).bfloat16()
    c = dco.float().permute(0, 3, 1, 2).requires_grad_(True)
    F.interpolate(c, size=(25, 42), mode='nearest').backward(dfine.float().permute(0, 3, 1, 2))
    got = ops.upsample_add_bwd(dfine, dco.clone())
    check(got, dco.float() + c.grad.permute(0, 2, 3, 1), 1e-2)
    # colsum / relu_bwd / sgd
    m = torch.randn(1000, 80, device='cuda', generator=g)
    out = torch.zeros(80, device='cuda')
    ops.colsum(m, out)
    check(out, m.sum(0), 1e-5)
    mb = torch.randn(777, 256, device='cuda', generator=g).bfloat16()
    out = torch.zeros(256, device='cuda')
    ops.colsum(mb, out)
    check(out, mb.float().sum(0), 1e-5)
    rb = ops.relu_bwd(dy, yq)
    assert torch.equal(rb, torch.where(yq > 0, dy, torch.zeros_like(dy)))
    w0 = torch.randn(10001, device='cuda', generator=g)
    gg = torch.randn(10001, device='cuda', generator=g)
    p = torch.nn.Parameter(w0.clone())
    opt = torch.optim.SGD([p], lr=0.01, momentum=0.9, weight_decay=1e-4)
    wmine, buf = w0.clone(), torch.zeros_like(w0)
    for it in range(3):
        p.grad = gg * (it + 1)
        opt.step()
        ops.sgd_flat(wmine, gg * (it + 1), buf, 0.01, 0.9, 1e-4, 1.0, first_step=(it == 0))
    torch.cuda.synchronize()
    check(wmine, p.detach(), 1e-6)


def test_conv_multi_level_equals_per_level():
    """The five FPN levels of a level-shared head conv in ONE launch == five single launches (bit-identical)."""
    from ld_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(3)
    sizes = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
    for cout, out_dtype, bias in ((256, torch.bfloat16, False), (80, torch.float32, True)):
        wt = ops.pack_conv_weight(torch.randn(cout, 256, 3, 3, device='cuda', generator=g) * 0.03)
        b = torch.randn(cout, device='cuda', generator=g) if bias else None
        xs = [torch.randn(2, h, w, 256, device='cuda', generator=g).bfloat16() for h, w in sizes]
        ys = ops.conv2d_tc_multi(xs, wt, cout, 3, 3, 1, 1, out_dtype=out_dtype, bias=b)
        for x, y in zip(xs, ys):
            one = ops.conv2d_tc(x, wt, cout, 3, 3, 1, 1, out_dtype=out_dtype, bias=b)
            assert torch.equal(y, one)
    torch.cuda.synchronize()
