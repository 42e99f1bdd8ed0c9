import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import ld_b200
from ld_b200 import presets, synth, ops
from ld_b200.config import ConfigDict
from ld_b200.backbones import rows_of
from oracle import ld_oracle as lo

sd, td, n, h, w, seed = 50, 101, 2, 256, 320, 22
gold = torch.load('tests/golden/s_r50_r101_2x256x320.pt', weights_only=False)
m = ld_b200.build_detector(ConfigDict(presets.ld_model(f'r{sd}', teacher_depth=td)))
synth.seeded_init(m, seed); synth.seeded_init(m.teacher_model, seed + 1, prefix='teacher.')
m = m.cuda().train()
b = synth.make_batch(n, h, w, seed=1000 + seed)
data = dict(img=b['img'].cuda(), img_metas=b['img_metas'], gt_bboxes=[x.cuda() for x in b['gt_bboxes']], gt_labels=[x.cuda() for x in b['gt_labels']])
losses = m(**data)
for k in losses:
    print(k, [round(float(v), 5) for v in losses[k]], [round(float(v), 5) for v in gold['losses'][k]])
aux = m.bbox_head._last_aux
print('counts', aux['targets'].counts.cpu().tolist()[:20])
# oracle with same weights: features
o = lo.OracleLDDetector(sd, td).train()
synth.seeded_init(o, seed); synth.seeded_init(o.teacher_model, seed + 1, prefix='teacher.')
with torch.no_grad():
    xo = o.neck(o.backbone(b['img'])); to = o.teacher_model.extract_feat(b['img'])
    tcls, treg = o.teacher_model.bbox_head(to)
    scls, sreg = o.bbox_head(xo)
xr = ops.to_rows(xo); tr = ops.to_rows(to)
def rel(a, bb, name):
    a = a.float().cpu(); bb = bb.float()
    print(f'{name:20s} rel_l2={float((a-bb).norm()/bb.norm()):.4f} ratio={float(a.norm()/bb.norm()):.4f}')
rel(aux['t_x_rows'], tr, 't_x_rows')
rel(aux['t_cls_rows'], ops.to_rows(tcls), 't_cls_rows')
rel(aux['t_reg_rows'], ops.to_rows(treg), 't_reg_rows')
rel(aux['cls_rows'], ops.to_rows(scls), 'cls_rows')
rel(aux['reg_rows'], ops.to_rows(sreg), 'reg_rows')
tg = aux['targets']
fg = tg.fg_rows[:int(tg.counts[1])].cpu().long()
d_o = ((xr[fg] - tr[fg]) ** 2).mean()
print('oracle mse on fg rows (all levels pooled)', float(d_o))
