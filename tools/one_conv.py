import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ld_b200 import ops
n, h, w, ci, co, k, st, pad = [int(v) for v in (sys.argv[1:9] if len(sys.argv) > 8 else (16, 100, 168, 128, 512, 1, 1, 0))]
x = torch.randn(n, h, w, ci, device='cuda').bfloat16()
wt = ops.pack_conv_weight(torch.randn(co, ci, k, k, device='cuda') * 0.02)
out = torch.empty(n, (h + 2 * pad - k) // st + 1, (w + 2 * pad - k) // st + 1, co, device='cuda', dtype=torch.bfloat16)
for _ in range(3):
    ops.conv2d_tc(x, wt, co, k, k, st, pad, out=out)
torch.cuda.synchronize()
