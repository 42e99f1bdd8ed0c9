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
