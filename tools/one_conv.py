"""One convolution (or its weight gradient) launched three times, for `ncu -s 2 -c 1`:
    python tools/one_conv.py n h w cin cout k stride pad [res|wgrad]
`res`: folded-BN scale/shift + identity residual + ReLU epilogue (bottleneck conv3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ld_b200 import ops
n, h, w, ci, co, k, st, pad = [int(v) for v in (sys.argv[1:9] if len(sys.argv) > 8 else (16, 100, 168, 128, 512, 1, 1, 0))]
mode = sys.argv[9] if len(sys.argv) > 9 else ''
ho, wo = (h + 2 * pad - k) // st + 1, (w + 2 * pad - k) // st + 1
x = torch.randn(n, h, w, ci, device='cuda').bfloat16()
if mode == 'wgrad':
    dy = torch.randn(n, ho, wo, co, device='cuda').bfloat16()
    dw = torch.zeros(co, k * k * ci, device='cuda')
    for _ in range(3):
        ops.conv2d_wgrad_tc(dy, x, dw, k, k, st, pad)
else:
    wt = ops.pack_conv_weight(torch.randn(co, ci, k, k, device='cuda') * 0.02)
    out = torch.empty(n, ho, wo, co, device='cuda', dtype=torch.bfloat16)
    kw = {}
    if mode == 'res':
        kw = dict(scale=torch.rand(co, device='cuda') + 0.5, bias=torch.randn(co, device='cuda'),
                  add=torch.randn(n, ho, wo, co, device='cuda').bfloat16(), relu=True)
    for _ in range(3):
        ops.conv2d_tc(x, wt, co, k, k, st, pad, out=out, **kw)
torch.cuda.synchronize()
