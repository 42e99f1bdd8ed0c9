"""Debug (1 GPU): the staged backward of LDTrainStep (the cuts the overlapped gradient exchange uses) against one
plain backward call on the same model and batch: gradients at the cuts and the flat gradient buffer per segment.
    python tests/debug/staged_bwd.py [depth]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault('LDB200_ALLOW_RANDOM_INIT', '1')
import torch

import ld_b200
from ld_b200 import presets, synth
from ld_b200.config import ConfigDict

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 18
dev = torch.device('cuda')
m = ld_b200.build_detector(ConfigDict(presets.ld_model(f'r{depth}', teacher_depth=18)))
synth.seeded_init(m, 0)
synth.seeded_init(m.teacher_model, 1, prefix='teacher.')
m = m.to(dev).train()
b = synth.make_batch(2, 256, 320, seed=100)
d = dict(img=b['img'].to(dev), img_metas=b['img_metas'], gt_bboxes=[x.to(dev) for x in b['gt_bboxes']],
         gt_labels=[x.to(dev) for x in b['gt_labels']])
ts = ld_b200.LDTrainStep(m, lr=0.0, warmup_iters=0)
ts.base_lr = 0.0


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


# plain backward, with the total / lateral gradients at the cuts retained
keep = {}


def hook(module, inputs, outputs):
    views = [t for t in inputs[0] if t.requires_grad][::-1]
    for k, v in enumerate(views):
        v.retain_grad()
        base = v._base
        base.retain_grad()
        keep[k] = (v, base)


h = m.neck.register_forward_hook(hook)
ts.step(d, sync_logs=False)
torch.cuda.synchronize()
h.remove()
g_plain = ts.flat_g.clone()
plain = {k: (v.grad.clone(), base.grad.clone()) for k, (v, base) in keep.items()}
keep.clear()
for k, (gv, gb) in plain.items():
    print(f'plain cut {k}: |view grad| {float(gv.float().norm()):.4e} |base grad| {float(gb.float().norm()):.4e} '
          f'view is base: {gv.data_ptr() == gb.data_ptr()}')
ts.step(d, sync_logs=False)
torch.cuda.synchronize()
print('plain vs plain again:', rel(ts.flat_g, g_plain))

ts._force_staged = ts.overlap = True
ts._hook_handle = m.neck.register_forward_hook(ts._neck_forward_hook)
ts._cut_log = {}
ts.step(d, sync_logs=False)
torch.cuda.synchronize()
log = ts._cut_log
for k in sorted(log):
    print(k, tuple(log[k].shape), f'{float(log[k].float().norm()):.4e}')
for k, (gv, gb) in plain.items():
    lat = log.get(f'lateral{k}')
    thr = log.get(f'through{k}')
    if lat is not None:
        print(f'cut {k}: staged lateral vs plain view grad: {rel(lat, gv):.3e}')
    if thr is not None and lat is not None:
        tot = lat.permute(0, 2, 3, 1).float() + thr.float()
        print(f'cut {k}: staged lateral + through vs plain base grad: {rel(tot, gb):.3e};  through vs (base - view): '
              f'{rel(thr.float(), gb.float() - gv.permute(0, 2, 3, 1).float()):.3e}')
segs = dict(ts._segments)
segs['rest'] = (0, min(a for a, b in ts._segments.values()))
for k, (a, b2) in sorted(segs.items(), key=lambda kv: kv[1]):
    print(f'segment {k:18s}: rel err {rel(ts.flat_g[a:b2], g_plain[a:b2]):.3e}  ratio '
          f'{float(ts.flat_g[a:b2].norm() / (g_plain[a:b2].norm() + 1e-30)):.4f}')
