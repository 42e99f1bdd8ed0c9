"""The launches furthest from their roofline in tools/layer_roofline.py, each on its own, for `ncu --set full -k ...`:
    python tools/hot_ops.py [op ...]      ops: gn_bwd gn_apply expand reduce conv3
"""
import os as _os; _os.environ.setdefault("LDB200_ALLOW_RANDOM_INIT", "1")
import sys
sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import torch
from ld_b200 import ops

which = sys.argv[1:] or ['gn_bwd', 'gn_apply', 'expand', 'reduce', 'conv3']
dev = 'cuda'
n = 16
sizes = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
g = torch.Generator(device=dev).manual_seed(0)


def bf(*shape):
    return torch.randn(*shape, device=dev, generator=g).bfloat16()


for rep in range(3):
    if 'gn_bwd' in which or 'gn_apply' in which:
        xs = [bf(n, h, w, 256) for h, w in sizes]
        gamma, beta = torch.rand(256, device=dev) + 0.5, torch.randn(256, device=dev) * 0.1
        stats = []
        for x in xs:
            r = x.float().reshape(n, -1, 32, 8)
            stats.append(torch.stack([r.sum((1, 3)), (r * r).sum((1, 3))], -1).contiguous())
        if 'gn_apply' in which:
            ys = ops.gn_apply_multi(xs, stats, gamma, beta)
        if 'gn_bwd' in which:
            dys = [bf(*x.shape) for x in xs]
            dg, db = torch.zeros(256, device=dev), torch.zeros(256, device=dev)
            ops.gn_relu_bwd_multi(dys, None, xs, stats, gamma, dg, db, beta=beta)
    if 'expand' in which:      # layer3 expand conv: 1x1 256 -> 1024 + BN + residual + ReLU (SHORTK kernel, HBM-bound)
        x, res = bf(n, 50, 84, 256), bf(n, 50, 84, 1024)
        w = (torch.randn(1024, 256, device=dev, generator=g) * 0.05).bfloat16()
        sc, sh = torch.rand(1024, device=dev) + 0.5, torch.randn(1024, device=dev)
        ops.conv2d_tc(x, w, 1024, scale=sc, bias=sh, add=res, relu=True)
    if 'reduce' in which:      # layer3 reduce conv: 1x1 1024 -> 256 + BN + ReLU (2-CTA kernel, 16 k-blocks)
        x = bf(n, 50, 84, 1024)
        w = (torch.randn(256, 1024, device=dev, generator=g) * 0.03).bfloat16()
        sc, sh = torch.rand(256, device=dev) + 0.5, torch.randn(256, device=dev)
        ops.conv2d_tc(x, w, 256, scale=sc, bias=sh, relu=True)
    if 'conv3' in which:       # layer3 3x3 256 -> 256 + BN + ReLU
        x = bf(n, 50, 84, 256)
        w = (torch.randn(256, 9 * 256, device=dev, generator=g) * 0.02).bfloat16()
        sc, sh = torch.rand(256, device=dev) + 0.5, torch.randn(256, device=dev)
        ops.conv2d_tc(x, w, 256, 3, 3, 1, 1, scale=sc, bias=sh, relu=True)
    torch.cuda.synchronize()
print('done')
