"""Debug (1 GPU): the staged backward of LDTrainStep under CUDA-graph capture (no exchange: world 1, cuts forced on)
against eager plain steps on a twin model.    python tests/debug/staged_graph.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault('LDB200_ALLOW_RANDOM_INIT', '1')
import torch

import ld_b200
from ld_b200 import presets, synth
from ld_b200.config import ConfigDict

dev = torch.device('cuda')


def make():
    m = ld_b200.build_detector(ConfigDict(presets.ld_model('r18', teacher_depth=18)))
    synth.seeded_init(m, 0)
    synth.seeded_init(m.teacher_model, 1, prefix='teacher.')
    return m.to(dev).train()


def batch(seed):
    b = synth.make_batch(2, 256, 320, seed=seed)
    return dict(img=b['img'].to(dev), img_metas=b['img_metas'], gt_bboxes=[x.to(dev) for x in b['gt_bboxes']],
                gt_labels=[x.to(dev) for x in b['gt_labels']])


import ctypes
_rt = ctypes.CDLL('libcudart.so.12')
_rt.cudaStreamIsCapturing.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]


def capture_status():
    """0 none, 1 active, 2 invalidated"""
    st = ctypes.c_int(-1)
    err = _rt.cudaStreamIsCapturing(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(st))
    return st.value if err == 0 else -err


from ld_b200 import lib as _L
_check = _L.check
_bad = [False]


def traced_check(code, what=''):
    _check(code, what)
    if not _bad[0] and capture_status() not in (0, 1):
        _bad[0] = True
        import traceback
        print(f'!! capture status {capture_status()} after {what}', flush=True)
        traceback.print_stack(limit=12)


_L.check = traced_check


_bw = torch.autograd.backward
_n = [0]


def traced_backward(*a, **k):
    r = _bw(*a, **k)
    _n[0] += 1
    if torch.cuda.is_current_stream_capturing() or True:
        print(f'  backward call {_n[0]} (inputs: {len(k.get("inputs") or ())}): capture status {capture_status()}', flush=True)
    return r


torch.autograd.backward = traced_backward

for staged in (False, True):
    ref, m = make(), make()
    ts_ref = ld_b200.LDTrainStep(ref, lr=0.001, warmup_iters=0)
    ts = ld_b200.LDTrainStep(m, lr=0.001, warmup_iters=0)
    if staged:
        os.environ['LDB200_GRAPH_OVERLAP'] = '1'
        ts._force_staged = ts.overlap = True
        ts._hook_handle = m.neck.register_forward_hook(ts._neck_forward_hook)
        ts._cut_log = {}
    if 'eager_first' in sys.argv:     # (needs LDB200_CAPTURE_AFTER_EAGER=1: capture() refuses this order)
        ts.step(batch(5), sync_logs=False)
        ts_ref.step(batch(5), sync_logs=False)
    # same sequence of batches on both: capture() runs `warmup` eager iterations on its batch
    try:
        ts.capture(batch(7), warmup=1)
    except Exception as e:
        print(f'staged={staged}: capture failed: {str(e).splitlines()[0]}', flush=True)
        torch.cuda.synchronize()
        continue
    ts_ref.step(batch(7), sync_logs=False)
    for it in range(3):
        l = ts.step_graphed(batch(20 + it))
        l0, _ = ts_ref.step(batch(20 + it), sync_logs=False)
        torch.cuda.synchronize()
        print(f'staged={staged} graphed step {it}: loss {float(l):.5f} vs eager plain {float(l0):.5f}', flush=True)
    print(f'staged={staged}: max |w - w_ref| = {float((ts.flat_w - ts_ref.flat_w).abs().max()):.3e}; cuts logged: '
          f'{sorted(getattr(ts, "_cut_log", {}) or {})}', flush=True)
