"""Kernel micro-benchmarks on one B200 (CUDA events on the launching stream, warm-up, inputs > L2).
    python tools/microbench.py [kl] [loss] [targets]
Prints one JSON line per kernel with achieved GB/s against MEASURED_PEAKS.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ld_b200 import lib as L, ops, synth  # noqa: E402

PEAK = 6568.7
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


def bench_kl():
    dev = torch.device('cuda')
    rows = 5 * 200000 * 4
    s = torch.randn(rows, 17, device=dev) * 2
    t = torch.randn(rows, 17, device=dev) * 2
    w = torch.rand(rows, device=dev)
    loss = torch.zeros(1, device=dev)
    ds = torch.empty_like(s)
    med, best = timeit(lambda: ops.kl_dense_fwd_bwd(s, t, w, loss=loss, d_student=ds))
    byts = rows * 208
    print(json.dumps(dict(kernel='kl_dense_fwd_bwd', rows=rows, ms_median=med, ms_best=best,
                          algorithmic_bytes=byts, gbs=byts / med / 1e6, frac_of_measured_hbm=byts / med / 1e6 / PEAK)))


def bench_loss_targets(which):
    dev = torch.device('cuda')
    n, h, w = 16, 800, 1344
    strides = (8, 16, 32, 64, 128)
    sizes, hh, ww = [], (h + 7) // 8, (w + 7) // 8      # k3 s2 p1 convs: ceil-halving per level
    for _ in strides:
        sizes.append((hh, ww))
        hh, ww = (hh + 1) // 2, (ww + 1) // 2
    pyr = L.Pyramid.make(n, sizes, strides)
    gb, gl = synth.make_gts(n, h, w, 9)
    gb = [b.cuda() for b in gb]
    gl = [b.cuda() for b in gl]
    if 'targets' in which:
        med, best = timeit(lambda: ops.ld_targets(pyr, gb, gl, [(h, w)] * n, device=dev))
        print(json.dumps(dict(kernel='ld_targets(3 launches + H2D meta)', N=n, ms_median=med, ms_best=best)))
    if 'loss' in which:
        tg = ops.ld_targets(pyr, gb, gl, [(h, w)] * n, device=dev)
        R = pyr.rows
        cls = torch.randn(R, 80, device=dev) * 1.5 - 2
        reg = torch.randn(R, 68, device=dev)
        t_cls = torch.randn(R, 80, device=dev) * 1.5 - 2
        t_reg = torch.randn(R, 68, device=dev) * 1.5
        x = torch.randn(R, 256, device=dev).bfloat16()
        t_x = torch.randn(R, 256, device=dev).bfloat16()
        cfg = ops.make_loss_cfg()
        wsum, _ = ops.ld_loss_prepass(cfg, cls, reg, tg)
        norm = ops.loss_normalisers(tg, wsum)
        cnt = tg.counts.cpu()
        for gd in (torch.float32, torch.bfloat16):
            med, best = timeit(lambda: ops.ld_loss_fused(cfg, cls, reg, t_cls, t_reg, x, t_x, tg, norm,
                                                         term_grad=True, grad_dtype=gd, max_fg=int(cnt[1])))
            gb_ = 2 if gd == torch.bfloat16 else 4
            # dense: read logits + labels/weights, write d_cls; zero-fill d_reg; sparse rows are noise
            byts = R * (80 * 4 + 8 + 80 * gb_ + 68 * gb_)
            print(json.dumps(dict(kernel=f'ld_loss_fused grad={gd}', rows=R, n_sparse=int(cnt[0]), n_fg=int(cnt[1]),
                                  ms_median=med, ms_best=best, algorithmic_bytes=byts, gbs=byts / med / 1e6,
                                  frac_of_measured_hbm=byts / med / 1e6 / PEAK)))
        for ph in (1, 2, 4):
            os.environ['LDB200_LOSS_PHASES'] = str(ph)
            med, best = timeit(lambda: ops.ld_loss_fused(cfg, cls, reg, t_cls, t_reg, x, t_x, tg, norm,
                                                         term_grad=True, grad_dtype=torch.bfloat16, max_fg=int(cnt[1])))
            print(json.dumps(dict(kernel=f'ld_loss_fused phase_mask={ph} (bf16 grads)', ms_median=med, ms_best=best)))
        os.environ.pop('LDB200_LOSS_PHASES')
        med, best = timeit(lambda: ops.ld_loss_prepass(cfg, cls, reg, tg))
        print(json.dumps(dict(kernel='ld_loss_prepass', ms_median=med, ms_best=best)))


def bench_conv():
    dev = torch.device('cuda')
    peak = 1680.9
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['bf16_tflops']
    except Exception:
        pass
    shapes = [('head P3 3x3 256->256', 16, 100, 168, 256, 256, 3, 1, 1),
              ('head P4 3x3 256->256', 16, 50, 84, 256, 256, 3, 1, 1),
              ('head P5 3x3 256->256', 16, 25, 42, 256, 256, 3, 1, 1),
              ('l2 1x1 512->128', 16, 100, 168, 512, 128, 1, 1, 0),
              ('l2 1x1 128->512', 16, 100, 168, 128, 512, 1, 1, 0),
              ('l3 3x3 256->256', 16, 50, 84, 256, 256, 3, 1, 1),
              ('l4 1x1 2048->512', 16, 25, 42, 2048, 512, 1, 1, 0),
              ('l1 3x3 64->64', 16, 200, 336, 64, 64, 3, 1, 1)]
    for name, n, h, w, ci, co, k, st, pad in shapes:
        x = torch.randn(n, h, w, ci, device=dev).bfloat16()
        wt = ops.pack_conv_weight(torch.randn(co, ci, k, k, device=dev) * 0.02)
        out = torch.empty(n, h, w, co, device=dev, dtype=torch.bfloat16)
        med, best = timeit(lambda: ops.conv2d_tc(x, wt, co, k, k, st, pad, out=out), iters=10, warm=3)
        fl = 2.0 * n * h * w * ci * co * k * k
        xt = x.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        wtt = torch.randn(co, ci, k, k, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
        medc, bestc = timeit(lambda: torch.nn.functional.conv2d(xt, wtt, None, st, pad), iters=10, warm=3)
        print(json.dumps(dict(kernel='conv2d_tc ' + name, ms_median=med, tflops=fl / med / 1e9,
                              frac_of_measured_bf16=fl / med / 1e9 / peak, cudnn_bf16_ms=medc,
                              cudnn_tflops=fl / medc / 1e9)))


if __name__ == '__main__':
    which = sys.argv[1:] or ['kl', 'loss', 'targets', 'conv']
    L.check(L.load().ldb200_check_device(), 'check_device')
    if 'kl' in which:
        bench_kl()
    if 'loss' in which or 'targets' in which:
        bench_loss_targets(which)
    if 'conv' in which:
        bench_conv()
