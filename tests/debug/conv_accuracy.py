"""Debug (GPU box): accumulation error of the split-bf16 convolution against fp64, one chain vs two chains, next to
torch's own fp32 convolution (cuDNN, TF32 off) on the same operands.
    python tests/debug/conv_accuracy.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F

from ld_b200 import ops

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator(device='cuda').manual_seed(1)
for (n, h, w, cin, cout, k) in ((2, 25, 42, 256, 256, 3), (2, 50, 84, 512, 128, 1), (2, 20, 28, 128, 512, 1), (1, 100, 168, 64, 64, 3)):
    x = torch.randn(n, h, w, cin, device='cuda', generator=g)
    wt = torch.randn(cout, cin, k, k, device='cuda', generator=g) * (2.0 / (cin * k * k)) ** 0.5
    want = F.conv2d(x.double().permute(0, 3, 1, 2), wt.double(), None, 1, k // 2).permute(0, 2, 3, 1)
    w3 = ops.split_weight(wt.permute(0, 2, 3, 1).reshape(cout, -1))
    res = {}
    for two, ch, name in ((False, 0, 'one chain'), (True, 0, 'lead+corr'), (True, 256, 'tap x 256ch'), (True, 64, 'tap x 64ch')):
        ops.PARITY_TWO_PASS, ops.PARITY_CHAIN_CH = two, ch
        res[name] = ops.conv2d_tc(x, w3, cout, k, k, 1, k // 2)
    res['torch fp32 (cuDNN)'] = F.conv2d(x.permute(0, 3, 1, 2), wt, None, 1, k // 2).permute(0, 2, 3, 1)
    res['torch fp32 (CPU)'] = F.conv2d(x.cpu().permute(0, 3, 1, 2), wt.cpu(), None, 1, k // 2).permute(0, 2, 3, 1).cuda()
    line = f'{k}x{k} {cin}->{cout} @{n}x{h}x{w}: '
    for name, y in res.items():
        d = (y.double() - want)
        line += f'{name} max {float(d.abs().max() / want.abs().max()):.2e} rms {float(d.norm() / want.norm()):.2e} | '
    print(line)
