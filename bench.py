#!/usr/bin/env python
"""bench.py -- images/sec of the LD train step (BASELINE.json `metric`):
GFocal-R50 student + frozen GFocal-R101 teacher, LDHead loss (QFL + DFL + GIoU + LD + VLR-LD + KD + IM),
synthetic COCO-shaped batches of 16 x 800x1344 (1333x800 padded to /32), student forward + backward,
teacher forward, target assignment, fused loss, gradient all-reduce, SGD.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

One JSON line on rank 0.  `value` = whole-job images/s with inputs resident in HBM, `e2e` = the same step
through the public API with the batch copied from pinned host memory and the loss read back every step.
Scaling: every GPU runs the metric's bs=16 batch (weak scaling, global batch 16 x N, lr scaled linearly as the
reference's configs prescribe) -- one gradient all-reduce per step is the only exchange.  `--global-batch 16` pins the
reference's own 8 x 2 global batch instead (strong scaling: 16/N images per GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault('LDB200_ALLOW_RANDOM_INIT', '1')   # synthetic benchmark: seeded random weights, no checkpoints

FLOP_PER_IMAGE = 1.8356e12          # 917.79 GMAC (student fwd+bwd 620.26, teacher fwd 297.53) x 2
GLOBAL_BATCH = 16
H, W = 800, 1344


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,' \
            'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,' \
            'clocks_event_reasons.sw_power_cap'
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return None
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


def host_cores():
    """Usable host cores: min(affinity, cgroup CPU quota) -- the GPU box exposes 128 logical CPUs but caps the
    container at 16 (cpu.max = 1600000/100000); oversubscribing 128 threads is 6x slower than 16."""
    n = len(os.sched_getaffinity(0))
    try:
        quota, period = open('/sys/fs/cgroup/cpu.max').read().split()
        if quota != 'max':
            n = min(n, max(1, int(float(quota) / float(period) + 0.5)))
    except Exception:
        pass
    return n


def cpu_reference_step(n_imgs, steps, warmup, seed=1000):
    """The reference's CPU training path (fp32, all host threads): the oracle port is the same ATen-CPU
    arithmetic as the unmodified reference (pinned by tests/golden) -- /root/reference does not exist here."""
    import torch
    from ld_b200 import synth
    from oracle import ld_oracle as lo
    torch.set_num_threads(host_cores())
    m = lo.OracleLDDetector(50, 101).train()
    synth.seeded_init(m, 0)
    synth.seeded_init(m.teacher_model, 1, prefix='teacher.')
    opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=0.01, momentum=0.9, weight_decay=1e-4)
    batch = synth.make_batch(n_imgs, H, W, seed=seed)
    ts = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        _, loss, _ = m.step(batch)
        opt.step()
        if it >= warmup:
            ts.append(time.perf_counter() - t0)
    return n_imgs / (sum(ts) / len(ts)), sum(ts) / len(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ld_b200', choices=['ld_b200', 'reference'])
    ap.add_argument('--student', default='r50')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='eager launches instead of CUDA-graph replay')
    ap.add_argument('--global-batch', type=int, default=0,
                    help='fix the global batch (strong scaling); default: 16 images per GPU (weak scaling)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    config = dict(workload='GFocal-R50 student + GFocal-R101 teacher LD train step, 16 x 800x1344 synthetic '
                           'COCO-shaped (configs/ld/ld_r50_gflv1_r101_fpn_coco_1x.py model section)',
                  global_batch=args.global_batch or GLOBAL_BATCH * args.gpus, per_gpu_batch=(args.global_batch or GLOBAL_BATCH * args.gpus) // args.gpus,
                  image=[H, W], parallelism=f'dp{args.gpus}',
                  launch='CUDA graph replay' if not args.no_graph else 'eager',
                  timing='inputs (51.6 MB images + activations >> 126 MB L2) exceed L2; no explicit flush')

    if args.impl == 'reference':
        if rank != 0:
            return
        config['launch'] = 'reference CPU path (oracle port, torch CPU fp32), bounded sample of 2 of the 16 images per step'
        n_imgs = 2
        v, sec = cpu_reference_step(n_imgs, max(1, min(args.steps, 3)), min(args.warmup, 1))
        cores = host_cores()
        print(json.dumps(dict(
            metric='images/sec GFocal-R50 student + R101 teacher LD train step, 1333x800 bs=16', impl='reference',
            value=v, unit='images/s', n_gpus=args.gpus, steps=max(1, min(args.steps, 3)), warmup=min(args.warmup, 1),
            ms_per_step=sec * 1e3 * (args.global_batch or GLOBAL_BATCH * args.gpus) / n_imgs, higher_is_better=True,
            scaling='strong' if args.global_batch else 'weak', vs_baseline=None,
            dtype='f32', data='synthetic', config=config,
            cpu_baseline=dict(value=v, unit='images/s', cores=cores, kind='port',
                              sample=f'{n_imgs} images of 800x1344 per step (of the 16-image batch), fp32, torch CPU'),
            e2e=dict(value=v, unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))))
        return

    import torch
    import torch.distributed as dist
    import ld_b200
    from ld_b200 import lib as L, presets, synth
    from ld_b200.config import ConfigDict
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world}'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    L.check(L.load().ldb200_check_device(), 'check_device')
    global_batch = args.global_batch or GLOBAL_BATCH * world
    scaling = 'strong' if args.global_batch else 'weak'
    assert global_batch % world == 0
    n = global_batch // world

    model = ld_b200.build_detector(ConfigDict(presets.ld_model(args.student)))
    synth.seeded_init(model, 0)
    synth.seeded_init(model.teacher_model, 1, prefix='teacher.')
    model = model.to(dev).train()
    step = ld_b200.LDTrainStep(model, lr=0.01 * global_batch / 16)   # linear scaling rule of the reference configs
    batch = synth.make_batch(n, H, W, seed=1000 + rank)
    img_pinned = batch['img'].pin_memory()
    gtb_pinned = [b.pin_memory() for b in batch['gt_bboxes']]
    gtl_pinned = [b.pin_memory() for b in batch['gt_labels']]
    resident = dict(img=img_pinned.to(dev), img_metas=batch['img_metas'],
                    gt_bboxes=[b.to(dev) for b in gtb_pinned], gt_labels=[b.to(dev) for b in gtl_pinned])

    host_loss = torch.empty(1, dtype=torch.float32).pin_memory()
    use_graph = not args.no_graph
    if use_graph:
        step.capture(resident, warmup=2)

    def run_resident():
        if use_graph:
            step.step_graphed(None)          # batch + targets already resident in the static buffers
        else:
            step.step(resident, sync_logs=False)

    e2e_data = dict(img=img_pinned, img_metas=batch['img_metas'], gt_bboxes=gtb_pinned, gt_labels=gtl_pinned)

    def run_e2e():
        # public API with HOST inputs: image batch + gt boxes from pinned memory, target assignment, the
        # captured iteration, loss read back
        data = e2e_data
        if use_graph:
            if getattr(step, '_staged', None) is not data:
                step.prefetch(data)          # first call only; afterwards the copy was started during the last step
            loss = step.step_graphed(data)
            step.prefetch(data)              # next batch's H2D overlaps this replay (every step still moves all bytes)
        else:
            data = dict(img=img_pinned.to(dev, non_blocking=True), img_metas=batch['img_metas'],
                        gt_bboxes=[b.to(dev, non_blocking=True) for b in gtb_pinned],
                        gt_labels=[b.to(dev, non_blocking=True) for b in gtl_pinned])
            loss, _ = step.step(data, sync_logs=False)
        host_loss.copy_(loss.reshape(1), non_blocking=True)

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.barrier()
            ms = float(t)
        return ms

    for _ in range(max(args.warmup, 3)):
        run_resident()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = L.launch_count['n']
    ms = timed(run_resident, args.steps)
    launches = (L.launch_count['n'] - l0) // args.steps
    if use_graph:
        launches = step.launches_per_step
    clocks = sampler.stop() if sampler else None
    run_e2e()
    ms_e2e = timed(run_e2e, args.steps)
    assert torch.isfinite(host_loss).all(), 'loss diverged'

    # dominant kernel, timed live: conv2_tc_kernel on the head-tower shape (P3 level, 3x3 256->256), the
    # single largest-FLOP launch of the step (32 forward + 16 dgrad launches of this shape per step)
    from ld_b200 import ops as _ops
    time.sleep(2.0)   # 'timed alone': let the power-capped clocks recover from the step loop before the burst measurement
    kx = torch.randn(n, 100, 168, 256, device=dev).bfloat16()
    kw = _ops.pack_conv_weight(torch.randn(256, 256, 3, 3, device=dev) * 0.02)
    ko = torch.empty(n, 100, 168, 256, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        _ops.conv2d_tc(kx, kw, 256, 3, 3, 1, 1, out=ko)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(20):
        _ops.conv2d_tc(kx, kw, 256, 3, 3, 1, 1, out=ko)
    k1.record()
    torch.cuda.synchronize()
    k_ms = k0.elapsed_time(k1) / 20
    k_flop = 2.0 * n * 100 * 168 * 256 * 256 * 9

    if rank == 0:
        pk = peaks()
        ips = global_batch * args.steps / (ms / 1e3)
        ips_e2e = global_batch * args.steps / (ms_e2e / 1e3)
        peak_tf = pk.get('bf16_tflops_sustained', 1400.0)
        ach = ips * FLOP_PER_IMAGE / 1e12 / world
        out = dict(metric='images/sec GFocal-R50 student + R101 teacher LD train step, 1333x800 bs=16',
                   value=ips, unit='images/s', n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                   ms_per_step=ms / args.steps, higher_is_better=True, scaling=scaling, vs_baseline=None,
                   dtype='bf16', data='synthetic', config=config,
                   e2e=dict(value=ips_e2e, unit='images/s',
                            h2d_bytes_per_step=int(img_pinned.numel() * 4 + sum(b.numel() * 4 for b in gtb_pinned)
                                                   + sum(b.numel() * 8 for b in gtl_pinned)) * world,
                            d2h_bytes_per_step=4 * world),
                   gpu_launches=int(launches),
                   roofline=dict(bound='tensor', kernel='conv2_tc_kernel (2-CTA tcgen05, 256x256 pair tile) head tower P3 3x3 256->256',
                                 achieved=k_flop / k_ms / 1e9, peak=pk.get('bf16_tflops', 1590.0), unit='TFLOP/s',
                                 frac=k_flop / k_ms / 1e9 / pk.get('bf16_tflops', 1590.0),
                                 traffic=230.7e6 * n / 16,
                                 note='algorithmic 2*N*H*W*Cin*Cout*9 FLOP per launch / live CUDA-event duration (20 '
                                      'back-to-back launches), of ' + ('measured' if 'bf16_tflops' in pk else 'fallback')
                                      + ' burst bf16 peak; traffic = dram read+write of the same launch from '
                                        'profiles/r01_conv2_head_p3_summary.txt (ncu --set full, batch 16)'),
                   roofline_step=dict(bound='tensor', achieved=ach, peak=peak_tf, unit='TFLOP/s', frac=ach / peak_tf,
                                      note='conv MACs of the whole step (1.8356 TFLOP/image) / step time per GPU, of '
                                           + ('measured' if 'bf16_tflops_sustained' in pk else 'fallback')
                                           + ' sustained bf16 peak'),
                   clocks=clocks)
        if world == 1 and not args.no_cpu_baseline:
            v, sec = cpu_reference_step(2, 1, 1)
            out['cpu_baseline'] = dict(value=v, unit='images/s', cores=host_cores(), kind='port',
                                       sample='one timed fp32 step (after one warm-up) on 2 images of 800x1344 of the batch; oracle port == reference ATen-CPU path')
        print(json.dumps(out), flush=True)
    if world > 1:
        # tearing down NCCL communicators that a captured CUDA graph still references can block forever:
        # make sure every rank is done, flush, and leave without running the destructors
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == '__main__':
    main()
