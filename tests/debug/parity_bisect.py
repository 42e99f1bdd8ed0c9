"""Debug (GPU box): where does the fp32-accurate mode of the CUDA path depart from the CPU oracle?  Prints the relative
error of the forward features (backbone+FPN levels, head outputs) and of EVERY parameter gradient, in network order.
    python tests/debug/parity_bisect.py [student_depth teacher_depth n h w seed]
TEST INFRASTRUCTURE (imports oracle/)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault('LDB200_ALLOW_RANDOM_INIT', '1')
import torch

import ld_b200
from ld_b200 import engine, presets, synth
from ld_b200.config import ConfigDict
from oracle import ld_oracle as lo

a = [int(v) for v in sys.argv[1:7]] if len(sys.argv) >= 7 else [18, 18, 2, 128, 160, 21]
sd, td, n, h, w, seed = a


def rel(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return float((got - want).norm() / (want.norm() + 1e-30)), float((got - want).abs().max() / (want.abs().max() + 1e-30))


m = ld_b200.build_detector(ConfigDict(presets.ld_model(f'r{sd}', teacher_depth=td)))
synth.seeded_init(m, seed)
synth.seeded_init(m.teacher_model, seed + 1, prefix='teacher.')
m = m.cuda().train()
om = lo.OracleLDDetector(sd, td).train()
synth.seeded_init(om, seed)
synth.seeded_init(om.teacher_model, seed + 1, prefix='teacher.')
batch = synth.make_batch(n, h, w, seed=1000 + seed)
img = batch['img']

with engine.precision('fp32'), torch.no_grad():
    feats = m.extract_feat(img.cuda())
    tfeats = m.teacher_model.extract_feat(img.cuda())
    cls, reg = m.bbox_head(feats)
    tcls, treg = m.teacher_model.bbox_head(tfeats)
with torch.no_grad():
    ofeats = om.neck(om.backbone(img))
    otfeats = om.teacher_model.extract_feat(img)
    ocls, oreg = om.bbox_head(ofeats)
    otcls, otreg = om.teacher_model.bbox_head(otfeats)
print('forward (rel-L2, max-abs/max):')
for nm, g, o in (('student fpn', feats, ofeats), ('teacher fpn', tfeats, otfeats), ('student cls', cls, ocls),
                 ('student reg', reg, oreg), ('teacher cls', tcls, otcls), ('teacher reg', treg, otreg)):
    for l, (x, y) in enumerate(zip(g, o)):
        print(f'  {nm} L{l}: %.2e %.2e' % rel(x.float(), y))

with engine.precision('fp32'):
    losses = m(img=img.cuda(), img_metas=batch['img_metas'], gt_bboxes=[b.cuda() for b in batch['gt_bboxes']],
               gt_labels=[b.cuda() for b in batch['gt_labels']])
    loss, log = m._parse_losses(losses)
    loss.backward()
torch.cuda.synchronize()
_, oloss, olog = om.step(batch)
print('losses:', {k: (round(log[k], 6), round(olog[k], 6)) for k in log})
onamed = dict(om.named_parameters())
print('gradients (rel-L2, max-abs/max, |ref|):')
for k, p in m.named_parameters():
    if p.grad is None:
        continue
    r = rel(p.grad, onamed[k].grad)
    flag = ' <<<' if max(r) > 1e-4 else ''
    print(f'  {k:55s} %.2e %.2e  %.3e{flag}' % (r[0], r[1], float(onamed[k].grad.norm())))
