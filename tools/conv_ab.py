"""Time a few convolution shapes of the LD step (run once with and once without LDB200_NO_CONV2=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ld_b200 import ops

SHAPES = [  # n, h, w, cin, cout, k, stride, pad
    (16, 100, 168, 256, 256, 3, 1, 1),    # head tower conv at P3
    (16, 50, 84, 1024, 256, 1, 1, 0),     # layer3 bottleneck reduce
    (16, 50, 84, 256, 256, 3, 1, 1),      # layer3 3x3
    (16, 100, 168, 512, 256, 1, 1, 0),    # FPN lateral P3
    (16, 25, 42, 512, 512, 3, 1, 1),      # layer4 3x3
    (16, 25, 42, 2048, 512, 1, 1, 0),     # layer4 reduce
]
tag = 'conv1' if os.environ.get('LDB200_NO_CONV2') else 'conv2'
for n, h, w, ci, co, k, st, pad in SHAPES:
    x = torch.randn(n, h, w, ci, device='cuda').bfloat16()
    wt32 = torch.randn(co, ci, k, k, device='cuda') * (2.0 / (ci * k * k)) ** 0.5
    wt = ops.pack_conv_weight(wt32)
    ho, wo = (h + 2 * pad - k) // st + 1, (w + 2 * pad - k) // st + 1
    out = torch.empty(n, ho, wo, co, device='cuda', dtype=torch.bfloat16)
    for _ in range(3):
        ops.conv2d_tc(x, wt, co, k, k, st, pad, out=out)
    torch.cuda.synchronize()
    want = torch.nn.functional.conv2d(x[:1].permute(0, 3, 1, 2).float(), wt32.bfloat16().float(), stride=st, padding=pad)
    err = float((out[:1].permute(0, 3, 1, 2).float() - want).abs().max() / want.abs().max())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        ops.conv2d_tc(x, wt, co, k, k, st, pad, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = 2.0 * n * ho * wo * co * ci * k * k / ms / 1e9
    print(f'{tag} {n}x{h}x{w} {ci}->{co} k{k}s{st}: {ms * 1e3:8.1f} us  {tf:7.1f} TF/s  relerr {err:.2e}', flush=True)

# weight gradients (LDB200_NO_WGRAD2=1 selects the 1-CTA kernel)
wtag = 'wgrad1' if os.environ.get('LDB200_NO_WGRAD2') else 'wgrad2'
for n, h, w, ci, co, k, st, pad in SHAPES[:1] + SHAPES[2:3] + SHAPES[1:2] + SHAPES[4:5]:
    ho, wo = (h + 2 * pad - k) // st + 1, (w + 2 * pad - k) // st + 1
    x = torch.randn(n, h, w, ci, device='cuda').bfloat16()
    dy = torch.randn(n, ho, wo, co, device='cuda').bfloat16()
    dw = torch.zeros(co, k * k * ci, device='cuda')
    for _ in range(3):
        ops.conv2d_wgrad_tc(dy, x, dw, k, k, st, pad)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        ops.conv2d_wgrad_tc(dy, x, dw, k, k, st, pad)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = 2.0 * n * ho * wo * co * ci * k * k / ms / 1e9
    print(f'{wtag} {n}x{h}x{w} {ci}->{co} k{k}s{st}: {ms * 1e3:8.1f} us  {tf:7.1f} TF/s', flush=True)
