"""Layer-by-layer comparison of the B200 backbone/FPN against the fp32 CPU oracle port."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import ld_b200
from ld_b200 import presets, synth, ops, engine
from ld_b200.config import ConfigDict
from ld_b200.backbones import nchw_view
from oracle import ld_oracle as lo

sd = int(sys.argv[1]) if len(sys.argv) > 1 else 50
seed = 22
m = ld_b200.build_detector(ConfigDict(presets.ld_model(f'r{sd}', teacher_depth=18)))
synth.seeded_init(m, seed)
m = m.cuda().train()
o = lo.OracleLDDetector(sd, 18).train()
synth.seeded_init(o, seed)
b = synth.make_batch(2, 256, 320, seed=1022)
img = b['img']

def rel(a, bb, name):
    a = a.float().cpu(); bb = bb.float()
    print(f'{name:40s} rel_l2={float((a-bb).norm()/bb.norm()):.4f}  mean_ratio={float(a.abs().mean()/bb.abs().mean()):.4f} shape={tuple(bb.shape)}')

with torch.no_grad():
    bo = o.backbone
    x = torch.nn.functional.max_pool2d(torch.relu(bo.bn1(bo.conv1(img))), 3, 2, 1)
    y = m.backbone.stem(img.cuda())
    rel(nchw_view(y), x, 'stem')
    for li in range(1, 5):
        lo_layer = getattr(bo, f'layer{li}')
        my_layer = getattr(m.backbone, f'layer{li}')
        for bi in range(len(lo_layer)):
            x = lo_layer[bi](x)
            # feed the ORACLE's previous activation into my block to isolate per-block error
            y_iso = my_layer[bi](y if False else None) if False else None
            y = my_layer[bi](y)
            rel(nchw_view(y), x, f'layer{li}.{bi}')
    feats_o = o.neck(o.backbone(img))
    feats_m = m.neck(m.backbone(img.cuda()))
    for i, (a, bb) in enumerate(zip(feats_m, feats_o)):
        rel(a, bb, f'fpn{i}')
