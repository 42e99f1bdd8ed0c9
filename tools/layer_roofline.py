"""Per-call roofline of one LD train step: every tensor-core launch (conv fprop / dgrad / wgrad) is timed on its own
with CUDA events (eager, serialised, two passes: the second is reported) and set against max(FLOPs / tensor peak,
algorithmic bytes / HBM peak) for its shape.  Answers "which shapes are furthest from THEIR bound" -- the step-level
fraction alone cannot.

    python tools/layer_roofline.py [batch] [student] > gpurun_out/layers.txt
"""
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import os as _os; _os.environ.setdefault("LDB200_ALLOW_RANDOM_INIT", "1")
import torch

import ld_b200
from ld_b200 import ops, presets, synth
from ld_b200.config import ConfigDict

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
student = sys.argv[2] if len(sys.argv) > 2 else 'r50'
pk = {}
try:
    pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
except Exception:
    pass
PEAK_TF, PEAK_GB = pk.get('bf16_tflops', 1680.9), pk.get('hbm_gbs', 6568.7)

records = []
recording = {'on': False}


def esize(t):
    return t.element_size()


def wrap(name, fn, describe):
    def inner(*a, **k):
        if not recording['on']:
            return fn(*a, **k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        torch.cuda.synchronize()
        key, flop, byts = describe(r, *a, **k)
        records.append((name, key, flop, byts, e0.elapsed_time(e1) * 1e3))
        return r
    return inner


def d_conv(r, x, w_packed, cout, kh=1, kw=1, stride=1, pad=0, out=None, out_dtype=None, scale=None, bias=None,
           add=None, mask=None, relu=False, accumulate=False, flat=None, taps=None, out_hw=None, out_map=None):
    nb, h, w, cin = x.shape
    _, ho, wo, _ = r.shape
    if out_map is not None:
        ho, wo = out_hw
    ntaps = kh * kw if taps is None else len(taps)
    flop = 2.0 * nb * ho * wo * cin * cout * ntaps
    byts = x.numel() * esize(x) / (1 if taps is None or out_map is None else 1) + nb * ho * wo * cout * esize(r) + \
        w_packed.numel() * 2
    if add is not None:
        byts += add.numel() * esize(add) if tuple(add.shape[1:3]) == (ho, wo) else 0
    if mask is not None:
        byts += nb * ho * wo * cout * esize(mask)
    kind = 'dgrad' if taps is not None else 'fprop'
    key = f'{kind} {kh}x{kw}/{stride} {cin:4d}->{cout:4d} @{nb}x{h}x{w}' + (' +add' if add is not None else '') + \
        (' +mask' if mask is not None else '') + (' f32out' if r.dtype == torch.float32 else '') + \
        (' parity-class' if out_map is not None else '')
    return key, flop, byts


def d_multi(r, xs, w_packed, cout, kh=1, kw=1, stride=1, pad=0, outs=None, out_dtype=None, scale=None, bias=None,
            masks=None, relu=False, taps=None, gn_stats=None):
    flop = byts = 0.0
    for x, y in zip(xs, r):
        nb, h, w, cin = x.shape
        flop += 2.0 * nb * h * w * cin * cout * kh * kw
        byts += x.numel() * esize(x) + nb * h * w * cout * esize(y)
    byts += w_packed.numel() * 2
    nb, h, w, cin = xs[0].shape
    kind = 'dgrad' if taps is not None else 'fprop'
    return f'{kind} x{len(xs)}lvl {kh}x{kw} {cin:4d}->{cout:4d} @{nb}x{h}x{w}..', flop, byts


def d_wgrad(r, dy, x, dw, kh=1, kw=1, stride=1, pad=0, scale=None, w_master=None, d_scale=None):
    nb, ho, wo, cout = dy.shape
    _, h, w, cin = x.shape
    flop = 2.0 * nb * ho * wo * cin * cout * kh * kw
    byts = dy.numel() * esize(dy) + x.numel() * esize(x) + dw.numel() * 4
    return f'wgrad {kh}x{kw}/{stride} {cin:4d}->{cout:4d} @{nb}x{h}x{w}', flop, byts


ops.conv2d_tc = wrap('conv', ops.conv2d_tc, d_conv)
ops.conv2d_tc_multi = wrap('conv', ops.conv2d_tc_multi, d_multi)
ops.conv2d_wgrad_tc = wrap('wgrad', ops.conv2d_wgrad_tc, d_wgrad)
for nm in ('gn_relu_fwd', 'gn_relu_bwd', 'gn_apply_multi', 'gn_relu_bwd_multi', 'colsum', 'maxpool3x3s2', 'stem_im2col', 'relu_bwd', 'upsample_add_bwd',
           'ld_loss_fused', 'ld_targets', 'sgd_flat', 'scatter_add_rows', 'ld_loss_prepass'):
    def mk(nm):
        def describe(r, *a, **k):
            t = [v for v in a if isinstance(v, torch.Tensor)]
            for v in a:
                if isinstance(v, (list, tuple)):
                    t += [u for u in v if isinstance(u, torch.Tensor)]
            byts = sum(v.numel() * v.element_size() for v in t)
            if isinstance(r, torch.Tensor):
                byts += r.numel() * r.element_size()
            elif isinstance(r, (tuple, list)):
                byts += sum(v.numel() * v.element_size() for v in r if isinstance(v, torch.Tensor))
            shp = 'x'.join(map(str, t[0].shape)) if t else ''
            return f'{nm} {shp}', 0.0, float(byts)
        return describe
    setattr(ops, nm, wrap('mem', getattr(ops, nm), mk(nm)))

dev = torch.device('cuda')
model = ld_b200.build_detector(ConfigDict(presets.ld_model(student)))
synth.seeded_init(model, 0)
synth.seeded_init(model.teacher_model, 1, prefix='teacher.')
model = model.to(dev).train()
step = ld_b200.LDTrainStep(model, lr=0.01)
b = synth.make_batch(n, 800, 1344, seed=1000)
data = dict(img=b['img'].to(dev), img_metas=b['img_metas'], gt_bboxes=[x.to(dev) for x in b['gt_bboxes']],
            gt_labels=[x.to(dev) for x in b['gt_labels']])
for _ in range(2):
    step.step(data, sync_logs=False)
recording['on'] = True
step.step(data, sync_logs=False)
records.clear()
step.step(data, sync_logs=False)
recording['on'] = False

agg = collections.OrderedDict()
for name, key, flop, byts, us in records:
    a = agg.setdefault((name, key), [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += flop
    a[2] += byts
    a[3] += us
tot_us = sum(v[3] for v in agg.values())
tot_bound = 0.0
rows = []
for (name, key), (cnt, flop, byts, us) in agg.items():
    bound_us = max(flop / (PEAK_TF * 1e12), byts / (PEAK_GB * 1e9)) * 1e6
    tot_bound += bound_us
    rows.append((us - bound_us, name, key, cnt, flop, byts, us, bound_us))
rows.sort(reverse=True)
print(f'# batch {n}, student {student}: {len(records)} timed calls, {tot_us / 1e3:.2f} ms serialised, '
      f'sum of per-call bounds {tot_bound / 1e3:.2f} ms (peaks {PEAK_TF} TF/s, {PEAK_GB} GB/s)')
print(f'# {"gap us":>8} {"us":>8} {"bound":>7} {"frac":>5} {"TF/s":>6} {"GB/s":>6} {"x":>3}  call')
for gap, name, key, cnt, flop, byts, us, bound_us in rows:
    print(f'{gap:10.0f} {us:8.0f} {bound_us:7.0f} {bound_us / us:5.2f} {flop / us / 1e6:6.0f} {byts / us / 1e3:6.0f} {cnt:3d}  {key}')
