"""One LD train step under the CUDA profiler range (for `ncu --profile-from-start off`) + CPU issue time.
    python tools/profile_step.py [batch] [student]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import os as _os; _os.environ.setdefault("LDB200_ALLOW_RANDOM_INIT", "1")
import torch
import ld_b200
from ld_b200 import presets, synth
from ld_b200.config import ConfigDict

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
student = sys.argv[2] if len(sys.argv) > 2 else 'r50'
dev = torch.device('cuda')
model = ld_b200.build_detector(ConfigDict(presets.ld_model(student)))
synth.seeded_init(model, 0); synth.seeded_init(model.teacher_model, 1, prefix='teacher.')
model = model.to(dev).train()
step = ld_b200.LDTrainStep(model, lr=0.01)
b = synth.make_batch(n, 800, 1344, seed=1000)
data = dict(img=b['img'].to(dev), img_metas=b['img_metas'], gt_bboxes=[x.to(dev) for x in b['gt_bboxes']],
            gt_labels=[x.to(dev) for x in b['gt_labels']])
for _ in range(3):
    step.step(data, sync_logs=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
step.step(data, sync_logs=False)
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_total = time.perf_counter() - t0
print(f'issue {t_issue*1e3:.1f} ms, total {t_total*1e3:.1f} ms, peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB')
# phase split
def timed(fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return r, (time.perf_counter() - t) * 1e3
step.flat_g.zero_()
losses, t_fwd = timed(lambda: model(**data))
loss = sum(sum(v) for k, v in losses.items() if 'loss' in k)
_, t_bwd = timed(lambda: loss.backward())
print(f'forward {t_fwd:.1f} ms, backward {t_bwd:.1f} ms')
torch.cuda.cudart().cudaProfilerStart()
step.step(data, sync_logs=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
