"""Stall hunt: `python tools/stress_replay.py stem N` launches the fused stem kernel N times on its own;
`python tools/stress_replay.py replay N [batch]` captures the LD step and replays it N times the way bench.py's resident
loop does (prefetch on the copy stream next to the replay).  Progress lines every 25 iterations."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('LDB200_ALLOW_RANDOM_INIT', '1')
import torch

import ld_b200
from ld_b200 import ops, presets, synth
from ld_b200.config import ConfigDict

mode, count = sys.argv[1], int(sys.argv[2])
dev = torch.device('cuda')
t0 = time.time()


def say(msg):
    print(f'[{mode} +{time.time() - t0:5.1f}s] {msg}', flush=True)


if mode == 'stem':
    n = 16
    g = torch.Generator(device='cuda').manual_seed(0)
    img = torch.randn(n, 3, 800, 1344, device=dev, generator=g)
    w2 = (torch.randn(128, 192, device=dev, generator=g) * 0.05).bfloat16()
    w2[:, 147:] = 0
    sc = [torch.rand(64, device=dev) + 0.5 for _ in range(2)]
    sh = [torch.randn(64, device=dev) for _ in range(2)]
    for i in range(count):
        ops.stem_tc(img, w2, [(sc[0], sh[0]), (sc[1], sh[1])])
        if i % 250 == 249:
            torch.cuda.synchronize()
            say(f'{i + 1} launches')
    torch.cuda.synchronize()
    say('done')
else:
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    model = ld_b200.build_detector(ConfigDict(presets.ld_model('r50')))
    synth.seeded_init(model, 0)
    synth.seeded_init(model.teacher_model, 1, prefix='teacher.')
    model = model.to(dev).train()
    step = ld_b200.LDTrainStep(model, lr=0.01)
    b = synth.make_batch(n, 800, 1344, seed=1000)
    data = dict(img=b['img'].to(dev), img_metas=b['img_metas'], gt_bboxes=[x.to(dev) for x in b['gt_bboxes']],
                gt_labels=[x.to(dev) for x in b['gt_labels']])
    step.capture(data, warmup=2)
    say(f'captured (fused stem: {os.environ.get("LDB200_NO_FUSED_STEM") is None})')
    for i in range(count):
        if getattr(step, '_staged', None) is not data:
            step.prefetch(data)
        step.step_graphed(data)
        step.prefetch(data)
        if i % 25 == 24:
            torch.cuda.synchronize()
            say(f'{i + 1} replays')
    torch.cuda.synchronize()
    say('done')
