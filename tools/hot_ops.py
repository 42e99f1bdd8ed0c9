"""The launches furthest from their roofline in tools/layer_roofline.py, each on its own, for `ncu --set full -k ...`:
    python tools/hot_ops.py [op ...]      ops: gn_bwd gn_apply expand reduce conv3
"""
import os as _os; _os.environ.setdefault("LDB200_ALLOW_RANDOM_INIT", "1")
import sys
sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import torch
from ld_b200 import ops

which = [] if 'convs' in sys.argv else (sys.argv[1:] or ['gn_bwd', 'gn_apply', 'expand', 'reduce', 'conv3'])
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

# ---- 1x1 / 3x3 conv shapes of the backbone only (kernel-choice A/B through the LDB200_* environment toggles):
#      python tools/hot_ops.py convs
if 'convs' in sys.argv:
    def t(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    print('env:', {k: v for k, v in _os.environ.items() if k.startswith('LDB200_') and k != 'LDB200_ALLOW_RANDOM_INIT'})
    for (cin, cout, hh, ww, res, k) in ((256, 1024, 50, 84, True, 1), (1024, 256, 50, 84, False, 1), (512, 128, 100, 168, False, 1),
                                        (128, 512, 100, 168, True, 1), (2048, 512, 25, 42, False, 1), (64, 256, 200, 336, True, 1),
                                        (512, 2048, 25, 42, True, 1), (256, 256, 50, 84, False, 3), (512, 512, 25, 42, False, 3),
                                        (128, 128, 100, 168, False, 3)):
        x = bf(n, hh, ww, cin)
        w = (torch.randn(cout, k * k * cin, device=dev, generator=g) * 0.05).bfloat16()
        sc, sh = torch.rand(cout, device=dev) + 0.5, torch.randn(cout, device=dev)
        r_ = bf(n, hh, ww, cout) if res else None
        o = torch.empty(n, hh, ww, cout, device=dev, dtype=torch.bfloat16)
        us = t(lambda: ops.conv2d_tc(x, w, cout, k, k, 1, k // 2, scale=sc, bias=sh, add=r_, relu=True, out=o))
        byts = (x.numel() + o.numel() * (2 if res else 1)) * 2
        fl = 2.0 * n * hh * ww * cin * cout * k * k
        print(f'conv {k}x{k} {cin:4d}->{cout:4d} @{hh}x{ww}{" +res" if res else "     "} {us:7.1f} us  {byts / us / 1e3:6.0f} GB/s  {fl / us / 1e6:6.0f} TF/s  '
              f'(bound {max(byts / 6568.7e3, fl / 1680.9e6):6.1f})')
    sys.exit(0)

# ---- plain timing (no profiler): python tools/hot_ops.py time
if 'time' in sys.argv:
    def t(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    xs = [bf(n, h, w, 256) for h, w in sizes]
    dys = [bf(*x.shape) for x in xs]
    gamma, beta = torch.rand(256, device=dev) + 0.5, torch.randn(256, device=dev) * 0.1
    stats = []
    for x in xs:
        r = x.float().reshape(n, -1, 32, 8)
        stats.append(torch.stack([r.sum((1, 3)), (r * r).sum((1, 3))], -1).contiguous())
    dg, db = torch.zeros(256, device=dev), torch.zeros(256, device=dev)
    ys = ops.gn_apply_multi(xs, stats, gamma, beta)
    el = sum(x.numel() for x in xs)
    print(f'gn_apply_multi           {t(lambda: ops.gn_apply_multi(xs, stats, gamma, beta)):7.1f} us  (4 B/elem bound {el * 4 / 6568.7e3:6.1f})')
    print(f'gn_relu_bwd_multi remask {t(lambda: ops.gn_relu_bwd_multi(dys, None, xs, stats, gamma, dg, db, beta=beta)):7.1f} us  (10 B/elem bound {el * 10 / 6568.7e3:6.1f})')
    print(f'gn_relu_bwd_multi read y {t(lambda: ops.gn_relu_bwd_multi(dys, ys, xs, stats, gamma, dg, db)):7.1f} us  (14 B/elem bound {el * 14 / 6568.7e3:6.1f})')
    for (cin, cout, hh, ww, res) in ((256, 1024, 50, 84, True), (1024, 256, 50, 84, False), (512, 128, 100, 168, False),
                                     (128, 512, 100, 168, True), (2048, 512, 25, 42, False), (64, 256, 200, 336, True),
                                     (512, 2048, 25, 42, True)):
        x = bf(n, hh, ww, cin)
        w = (torch.randn(cout, cin, device=dev, generator=g) * 0.05).bfloat16()
        sc, sh = torch.rand(cout, device=dev) + 0.5, torch.randn(cout, device=dev)
        r_ = bf(n, hh, ww, cout) if res else None
        o = torch.empty(n, hh, ww, cout, device=dev, dtype=torch.bfloat16)
        us = t(lambda: ops.conv2d_tc(x, w, cout, scale=sc, bias=sh, add=r_, relu=True, out=o))
        byts = (x.numel() + o.numel() * (2 if res else 1)) * 2
        fl = 2.0 * n * hh * ww * cin * cout
        print(f'conv 1x1 {cin:4d}->{cout:4d} @{hh}x{ww}{" +res" if res else "     "} {us:7.1f} us  {byts / us / 1e3:6.0f} GB/s  {fl / us / 1e6:6.0f} TF/s  '
              f'(bound {max(byts / 6568.7e3, fl / 1680.9e6):6.1f})')

# ---- stem routes: python tools/hot_ops.py stem
if 'stem' in sys.argv:
    def t(fn, reps=10):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    img = torch.randn(n, 3, 800, 1344, device=dev, generator=g)
    w2 = (torch.randn(128, 192, device=dev, generator=g) * 0.05).bfloat16()
    w2[:, 147:] = 0
    sc = [torch.rand(64, device=dev) + 0.5 for _ in range(2)]
    sh = [torch.randn(64, device=dev) for _ in range(2)]
    print(f'stem_im2col              {t(lambda: ops.stem_im2col(img)):8.1f} us')
    a, ho, wo = ops.stem_im2col(img)
    print(f'stem GEMM (one network)  {t(lambda: ops.conv2d_tc(a.view(1, 1, -1, 192), w2[:64].contiguous(), 64, scale=sc[0], bias=sh[0], relu=True)):8.1f} us')
    print(f'fused stem, 2 networks   {t(lambda: ops.stem_tc(img, w2, [(sc[0], sh[0]), (sc[1], sh[1])])):8.1f} us')
    print(f'fused stem, 1 network    {t(lambda: ops.stem_tc(img, w2[:64].contiguous(), [(sc[0], sh[0])])):8.1f} us')
    y = ops.stem_tc(img, w2[:64].contiguous(), [(sc[0], sh[0])])[0]
    print(f'maxpool                  {t(lambda: ops.maxpool3x3s2(y)):8.1f} us')
