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
FLOP_PER_IMAGE_R18 = 1.5435e12      # R18 student variant (BASELINE configs[1]): 771.77 GMAC x 2
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
                                          '--format=csv,noheader,nounits', '-lms', '200'],
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


def cpu_reference_step(n_imgs, steps, warmup, seed=1000, progress=None):
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
        if progress:
            progress(f'cpu step {it + 1}/{warmup + steps}: {time.perf_counter() - t0:.1f} s')
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
    ap.add_argument('--per-gpu-batch', type=int, default=0, help='images per GPU (BASELINE configs[1]: --student r18 --per-gpu-batch 8)')
    ap.add_argument('--main-ld-only', action='store_true', help='BASELINE configs[2]: VLR / KD / IM loss weights 0')
    ap.add_argument('--ldv2', action='store_true',
                    help='BASELINE configs[3]: LDv2Head R50 student + GFocalHead X101-32x4d-DCN teacher (VLR + IMLoss)')
    ap.add_argument('--no-parity', action='store_true', help='skip the fp32-accurate mode side measurement')
    ap.add_argument('--no-micro', action='store_true', help='skip the micro records (LD-KL dense kernel)')
    ap.add_argument('--child', action='store_true', help=argparse.SUPPRESS)   # the measuring process under the supervisor
    ap.add_argument('--clocks', default=None, choices=['during', 'after', 'off'],
                    help='when nvidia-smi samples the clocks: during the timed region (the profiling recipe; default), '
                         'over an identical untimed loop after all measurements (kept from the stall hunt of DESIGN 10: '
                         'the stalls also happen without nvidia-smi next to the loop), or not at all')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    config = dict(workload='GFocal-R50 student + GFocal-R101 teacher LD train step, 16 x 800x1344 synthetic '
                           'COCO-shaped (configs/ld/ld_r50_gflv1_r101_fpn_coco_1x.py model section)',
                  global_batch=args.global_batch or (args.per_gpu_batch or GLOBAL_BATCH) * args.gpus,
                  per_gpu_batch=(args.global_batch or (args.per_gpu_batch or GLOBAL_BATCH) * args.gpus) // args.gpus,
                  student=args.student, losses='main LD only' if args.main_ld_only else 'LD + VLR-LD + KD + IM (ld_r50 config)',
                  head='LDv2Head <- GFocalHead X101-32x4d-DCN teacher (BASELINE configs[3])' if args.ldv2 else 'LDHead <- GFLHead R101 teacher',
                  image=[H, W], parallelism=f'dp{args.gpus}',
                  launch='CUDA graph replay' if not args.no_graph else 'eager',
                  timing='inputs (51.6 MB images + activations >> 126 MB L2) exceed L2; no explicit flush',
                  value_region='per step: image D2D into the static buffer, target assignment on resident gt boxes, '
                               'graph replay (student fwd+bwd, teacher fwd, loss, all-reduce, SGD); e2e adds H2D + loss D2H')

    if args.impl == 'reference':
        if rank != 0:
            return
        config['launch'] = 'reference CPU path (oracle port, torch CPU fp32)'
        n_imgs = 2
        config['sample'] = (f'each timed step = {n_imgs} images of the 16-image 800x1344 batch (bounded sample: a full '
                            f'batch is ~70 s of CPU work); images/s = {n_imgs} / measured step time, ms_per_step is '
                            f'the measured time of that sample step, NOT scaled to 16 images')
        v, sec = cpu_reference_step(n_imgs, max(1, min(args.steps, 3)), min(args.warmup, 1))
        cores = host_cores()
        print(json.dumps(dict(
            metric='images/sec GFocal-R50 student + R101 teacher LD train step, 1333x800 bs=16', impl='reference',
            value=v, unit='images/s', n_gpus=args.gpus, steps=max(1, min(args.steps, 3)), warmup=min(args.warmup, 1),
            ms_per_step=sec * 1e3, higher_is_better=True,
            scaling='strong' if args.global_batch else 'weak', vs_baseline=None,
            dtype='f32', data='synthetic', config=config,
            cpu_baseline=dict(value=v, unit='images/s', cores=cores, kind='port',
                              sample=f'{n_imgs} images of 800x1344 per step (of the 16-image batch), fp32, torch CPU'),
            e2e=dict(value=v, unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))))
        return

    if not args.child and os.environ.get('LDB200_BENCH_NO_SUPERVISOR') is None:
        return supervise(rank)
    if os.environ.get('LDB200_BENCH_SELFTEST'):
        # tests/test_bench_supervisor.py (CPU): a stand-in measuring process -- progress lines, optionally one stall on
        # the first attempt (no `--clocks after` yet), then a result line
        print('[bench rank %d +   0.0s] selftest child' % rank, file=sys.stderr, flush=True)
        att = int(os.environ.get('LDB200_BENCH_ATTEMPT', '0'))
        if os.environ['LDB200_BENCH_SELFTEST'] == 'stall-once' and att == 0:
            time.sleep(3600)
        ph = None
        if os.environ['LDB200_BENCH_SELFTEST'] == 'phases':      # phase b stalls on attempts 0 and 1
            st = PhaseState(os.environ.get('LDB200_BENCH_STATE'), rank, att)
            say = lambda m: print('[bench rank %d] %s' % (rank, m), file=sys.stderr, flush=True)   # noqa: E731
            ph = dict(a=st.run('a', lambda: att, say), b=st.run('b', lambda: time.sleep(3600) if att < 2 else att, say),
                      c=st.run('c', lambda: att, say))
        if rank == 0:
            print(json.dumps(dict(metric='selftest', value=1.0, clocks_mode=args.clocks, phases=ph)), flush=True)
        return

    import torch
    import torch.distributed as dist
    import ld_b200
    from ld_b200 import lib as L, ops as _ops, presets, synth
    from ld_b200.config import ConfigDict
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world}'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    L.check(L.load().ldb200_check_device(), 'check_device')
    per_gpu = args.per_gpu_batch or GLOBAL_BATCH
    global_batch = args.global_batch or per_gpu * world
    scaling = 'strong' if args.global_batch else 'weak'
    assert global_batch % world == 0

    if args.ldv2:
        model = ld_b200.build_detector(ConfigDict(presets.ldv2_model(args.student)))
    else:
        model = ld_b200.build_detector(ConfigDict(presets.ld_model(args.student, main_ld_only=args.main_ld_only)))
    synth.seeded_init(model, 0)
    synth.seeded_init(model.teacher_model, 1, prefix='teacher.')
    model = model.to(dev).train()
    use_graph = not args.no_graph

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

    t_start = time.time()
    activity = {'t': time.time()}
    pending = {'out': None, 'probe': False}     # result line ready while the clock probe runs (see watchdog)
    clocks_mode = args.clocks or 'during'

    def trace(msg):      # progress on stderr (the JSON line on stdout is the contract; the supervisor watches these lines)
        activity['t'] = time.time()
        print(f'[bench rank {rank} +{time.time() - t_start:6.1f}s] {msg}', file=sys.stderr, flush=True)

    def watchdog():
        # a phase that makes no progress for STALL_S seconds: say where every thread is and leave, so that the launcher
        # (the supervisor at N = 1, torchrun at N > 1) sees a failed process instead of a silent hang
        import faulthandler
        while True:
            time.sleep(5.0)
            if time.time() - activity['t'] > STALL_S:
                print(f'[bench rank {rank}] no progress for {STALL_S} s -- stacks follow', file=sys.stderr, flush=True)
                faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
                if pending['probe']:      # every measurement is done: only the clock sampling loop stalled
                    if pending['out'] is not None:
                        pending['out']['clocks'] = dict(error=f'the clock-sampling loop stalled for {STALL_S} s; timings were taken before it')
                        print(json.dumps(pending['out']), flush=True)
                    os._exit(0)
                os._exit(17)

    threading.Thread(target=watchdog, daemon=True).start()

    def measure(n, gbatch, precision='bf16', graph=True, steps=args.steps, warmup=max(args.warmup, 3), clocks=False):
        """One configuration: `n` images per GPU.  -> dict(ms, ms_e2e, launches, clocks, bytes)."""
        trace(f'measure: {n} images/GPU, {precision}, graph={graph}')
        step = ld_b200.LDTrainStep(model, lr=0.01 * gbatch / 16, precision=precision)   # linear scaling rule
        batch = synth.make_batch(n, H, W, seed=1000 + rank)
        img_pinned = batch['img'].pin_memory()
        gtb_pinned = [b.pin_memory() for b in batch['gt_bboxes']]
        gtl_pinned = [b.pin_memory() for b in batch['gt_labels']]
        resident = dict(img=img_pinned.to(dev), img_metas=batch['img_metas'],
                        gt_bboxes=[b.to(dev) for b in gtb_pinned], gt_labels=[b.to(dev) for b in gtl_pinned])
        host_loss = torch.empty(1, dtype=torch.float32).pin_memory()
        if graph:
            step.capture(resident, warmup=2)
            trace('captured')

        def run_resident():
            # inputs resident in HBM; everything else of the step runs every step: image move into the graph's static
            # buffer, target assignment (ldb200_ld_targets) on the resident gt boxes -- both issued for the NEXT batch
            # on the copy stream while the current iteration replays, as in the e2e path -- and the captured iteration
            if graph:
                if getattr(step, '_staged', None) is not resident:
                    step.prefetch(resident)
                step.step_graphed(resident)
                step.prefetch(resident)
            else:
                step.step(resident, sync_logs=False)

        e2e_data = dict(img=img_pinned, img_metas=batch['img_metas'], gt_bboxes=gtb_pinned, gt_labels=gtl_pinned)

        def run_e2e():
            # public API with HOST inputs: image batch + gt boxes from pinned memory, target assignment, the
            # captured iteration, loss read back
            data = e2e_data
            if graph:
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

        for _ in range(warmup):
            run_resident()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank) if (rank == 0 and (clocks == 'probe' or (clocks and clocks_mode == 'during'))) else None
        if sampler:
            sampler.start()
        l0 = L.launch_count['n']
        ms = timed(run_resident, steps)
        trace(f'resident: {ms / steps:.2f} ms/step')
        launches = (L.launch_count['n'] - l0) // steps
        if graph:
            launches = step.launches_per_step + 3      # + the eager target-assignment kernels of every step
        ck = sampler.stop() if sampler else None
        if clocks == 'probe':
            model.bbox_head._static_targets = None
            del step
            torch.cuda.empty_cache()
            return dict(clocks=ck)
        run_e2e()
        ms_e2e = timed(run_e2e, steps)
        trace(f'e2e: {ms_e2e / steps:.2f} ms/step')
        assert torch.isfinite(host_loss).all(), 'loss diverged'
        h2d = int(img_pinned.numel() * 4 + sum(b.numel() * 4 for b in gtb_pinned) + sum(b.numel() * 8 for b in gtl_pinned))
        res = dict(ms=ms / steps, ms_e2e=ms_e2e / steps, launches=int(launches), clocks=ck, h2d=h2d)
        # release the graph, its static buffers and the flat optimizer state before the next configuration
        model.bbox_head._static_targets = None
        del step
        torch.cuda.empty_cache()
        return res

    n = global_batch // world
    phases = PhaseState(os.environ.get('LDB200_BENCH_STATE'), rank, int(os.environ.get('LDB200_BENCH_ATTEMPT', '0')))
    main_res = phases.run('main', lambda: measure(n, global_batch, graph=use_graph, clocks=True), trace)
    ms, ms_e2e = main_res['ms'], main_res['ms_e2e']

    # BASELINE configs[2] as stated (8 x 2 images: GLOBAL batch 16): the same step with 16 / N images per GPU
    strong = None
    if world > 1 and not args.global_batch and not args.per_gpu_batch and GLOBAL_BATCH % world == 0:
        r = phases.run('strong', lambda: measure(GLOBAL_BATCH // world, GLOBAL_BATCH, graph=use_graph), trace)
        strong = dict(value=GLOBAL_BATCH / (r['ms'] / 1e3), unit='images/s', ms_per_step=r['ms'], global_batch=GLOBAL_BATCH,
                      per_gpu_batch=GLOBAL_BATCH // world, scaling='strong',
                      e2e=dict(value=GLOBAL_BATCH / (r['ms_e2e'] / 1e3), unit='images/s'))

    # fp32-accurate ("parity") mode of the same step: fp32 activations, split-bf16 convolutions (tests/test_gpu_parity.py
    # holds it to 1e-4 against the reference).  Eager launches, 3 timed steps.
    parity = None
    if world == 1 and not args.no_parity:
        trace('parity-mode side measurement')

        def parity_phase():
            try:
                r = measure(n, global_batch, precision='fp32', graph=False, steps=3, warmup=2)
                return dict(img_s=global_batch / (r['ms'] / 1e3), ms_per_step=r['ms'], launch='eager',
                            note='LDTrainStep(precision="fp32"): fp32 NHWC activations, every conv / dgrad / wgrad as 6 '
                                 'bf16 plane products in one fp32 TMEM accumulator; loss+grads within 1e-4 of the reference')
            except torch.OutOfMemoryError as e:      # never let the side measurement take the headline down
                torch.cuda.empty_cache()
                return dict(error=str(e)[:120])
        parity = phases.run('parity', parity_phase, trace)

    # dominant kernel, timed live: conv2_tc_kernel on the head-tower shape (P3 level, 3x3 256->256), the
    # single largest-FLOP launch shape of the step (the 5-level head launches are 48 of the step's tcgen05 launches)
    trace('dominant kernel, timed alone')
    time.sleep(2.0)   # 'timed alone': let the power-capped clocks recover from the step loop before the burst measurement
    kx = torch.randn(n, 100, 168, 256, device=dev).bfloat16()
    kw = _ops.pack_conv_weight(torch.randn(256, 256, 3, 3, device=dev) * 0.02)
    ko = torch.empty(n, 100, 168, 256, device=dev, dtype=torch.bfloat16)

    def ktime(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(reps):
            fn()
        k1.record()
        torch.cuda.synchronize()
        return k0.elapsed_time(k1) / reps

    k_ms = phases.run('kernel', lambda: ktime(lambda: _ops.conv2d_tc(kx, kw, 256, 3, 3, 1, 1, out=ko)), trace)
    k_flop = 2.0 * n * 100 * 168 * 256 * 256 * 9

    # micro records (BASELINE configs[4] and the fused loss kernel north_star singles out), HBM-bound
    micro = None
    if rank == 0 and not args.no_micro:
        pk_hbm = peaks().get('hbm_gbs', 6500.0)
        rows = 5 * 200000 * 4
        S = torch.randn(rows, 17, device=dev) * 2
        Tt = torch.randn(rows, 17, device=dev) * 2
        wv = torch.rand(rows, device=dev)
        lossb, dS = torch.zeros(1, device=dev), torch.empty_like(S)
        t_kl = phases.run('micro', lambda: ktime(lambda: _ops.kl_dense_fwd_bwd(S, Tt, wv, 10.0, 0.25, 4.0, loss=lossb, d_student=dS)), trace)
        micro = dict(ld_kl_dense=dict(rows=rows, bins=17, ms=t_kl, bytes=208 * rows, achieved_gbs=208 * rows / t_kl / 1e6,
                                      peak_gbs=pk_hbm, frac=208 * rows / t_kl / 1e6 / pk_hbm,
                                      note='BASELINE configs[4]: 5 levels x 200k anchors x 4 sides x 17 bins, fwd+bwd, '
                                           '208 algorithmic B/row (SURVEY 8d)'))
        del S, Tt, wv, dS

    if rank == 0:
        pk = peaks()
        ips = global_batch / (ms / 1e3)
        ips_e2e = global_batch / (ms_e2e / 1e3)
        peak_tf = pk.get('bf16_tflops_sustained', 1400.0)
        flop_img = FLOP_PER_IMAGE if args.student == 'r50' else FLOP_PER_IMAGE_R18
        ach = ips * flop_img / 1e12 / world
        prof = profile_numbers()
        out = dict(metric='images/sec GFocal-R50 student + R101 teacher LD train step, 1333x800 bs=16',
                   value=ips, unit='images/s', n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                   ms_per_step=ms, higher_is_better=True, scaling=scaling, vs_baseline=None,
                   dtype='bf16', data='synthetic', config=config,
                   e2e=dict(value=ips_e2e, unit='images/s', h2d_bytes_per_step=main_res['h2d'] * world,
                            d2h_bytes_per_step=4 * world),
                   gpu_launches=main_res['launches'],
                   roofline=dict(bound='tensor', kernel='conv2_tc_kernel (2-CTA tcgen05, 256x256 pair tile) head tower P3 3x3 256->256',
                                 achieved=k_flop / k_ms / 1e9, peak=pk.get('bf16_tflops', 1590.0), unit='TFLOP/s',
                                 frac=k_flop / k_ms / 1e9 / pk.get('bf16_tflops', 1590.0),
                                 traffic=prof.get('traffic'), traffic_source=prof.get('traffic_source'),
                                 share_of_step=prof.get('share'), share_source=prof.get('share_source'),
                                 note='algorithmic 2*N*H*W*Cin*Cout*9 FLOP per launch / live CUDA-event duration (20 '
                                      'back-to-back launches), of ' + ('measured' if 'bf16_tflops' in pk else 'fallback')
                                      + ' burst bf16 peak; traffic = dram read+write per launch of the committed ncu '
                                        '--set full capture (batch 16), share_of_step = this kernel\'s share of the '
                                        'kernel time in the committed launch list'),
                   roofline_step=dict(bound='tensor', achieved=ach, peak=peak_tf, unit='TFLOP/s', frac=ach / peak_tf,
                                      note=f'conv MACs of the whole step ({flop_img / 1e12:.4f} TFLOP/image) / step time per GPU, of '
                                           + ('measured' if 'bf16_tflops_sustained' in pk else 'fallback')
                                           + ' sustained bf16 peak'),
                   clocks=main_res['clocks'])
        if strong is not None:
            out['strong'] = strong
        if parity is not None:
            out['parity_mode'] = parity
        if micro is not None:
            out['micro'] = micro
        if world == 1 and not args.no_cpu_baseline:
            trace('cpu baseline (oracle port on the host cores)')
            v, sec = phases.run('cpu_baseline', lambda: list(cpu_reference_step(2, 1, 1, progress=trace)), trace)
            trace('cpu baseline done')
            out['cpu_baseline'] = dict(value=v, unit='images/s', cores=host_cores(), kind='port',
                                       sample='one timed fp32 step (after one warm-up) on 2 images of 800x1344 of the batch; oracle port == reference ATen-CPU path')
        pending['out'] = out
    if clocks_mode == 'after':
        # the same replay loop once more, untimed, with nvidia-smi polling next to it; every number above is final
        pending['probe'] = True
        trace('clock probe (identical untimed loop)')
        ck = measure(n, global_batch, graph=use_graph, clocks='probe')['clocks']
        if rank == 0:
            out['clocks'] = dict(ck or {}, note='sampled over an identical untimed replay loop right after the measurements')
        pending['probe'] = False
    if rank == 0:
        pending['out'] = None
        print(json.dumps(out), flush=True)
    if world > 1:
        # tearing down NCCL communicators that a captured CUDA graph still references can block forever:
        # make sure every rank is done, flush, and leave without running the destructors
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


class PhaseState:
    """Results of the measurement phases that have completed (main, strong, parity, kernel, micro, cpu_baseline), kept in
    a file that survives the supervisor's retries: a repeated attempt takes them from the file instead of measuring
    them again, so a stall costs one phase, not the run.  Every phase is a self-contained measurement (own capture,
    warm-up and timed region).  Rank 0 writes the file; every rank of the node reads it when its attempt starts (all
    children of the previous attempt are gone by then, and a phase with collectives cannot complete on rank 0 alone)."""

    def __init__(self, path, rank, attempt):
        self.path, self.rank, self.done = path, rank, {}
        if path and attempt > 0 and os.path.exists(path):
            try:
                with open(path) as f:
                    self.done = json.load(f)
            except Exception:
                self.done = {}

    def run(self, name, fn, trace=None):
        if name in self.done:
            if trace:
                trace(f'{name}: taken from an earlier attempt of this run')
            return self.done[name]
        r = fn()
        self.done[name] = r
        if self.path and self.rank == 0:
            tmp = self.path + '.tmp'
            with open(tmp, 'w') as f:
                json.dump(self.done, f)
            os.replace(tmp, self.path)
        return r


STALL_S = int(os.environ.get('LDB200_BENCH_STALL_S', 40))           # seconds without a progress line before a measuring process gives up (longest quiet phase: one CPU step, ~10 s)
SUPERVISOR_STALL_S = STALL_S + 20 if 'LDB200_BENCH_STALL_S' not in os.environ else STALL_S
ATTEMPTS = 12
START_S = 0 if 'LDB200_BENCH_STALL_S' in os.environ else 180     # patience before the child's first progress line
BUDGET_S = int(os.environ.get('LDB200_BENCH_BUDGET_S', 780))   # no new attempt after this many seconds (driver: 870 s per N)


def supervise(rank):
    """The measurement runs in a child process (`--child`, same environment: under torchrun every rank's process
    supervises its own child and the children do the rendezvous); this process only watches the child's progress
    lines.  A child that stops making progress is killed by PID and the measurement repeated (late in round 2 about one
    run in three stalled inside the first replays after a graph capture -- with and without the fused stem, with and
    without nvidia-smi polling, cause not found; the stalled kernel dies with the process).  Every rank's child stalls
    together (they meet in the all-reduce), gives up after STALL_S and is restarted here, so the new children
    rendezvous again on the same MASTER_PORT.  Rank 0 prints the child's JSON line."""
    import signal
    try:        # map the CUDA library here too (no context is created): the library is the product under test
        from ld_b200 import lib as L
        L.load()
    except Exception as e:     # the child will fail loudly with the real reason
        print(f'[bench supervisor {rank}] could not load the C-ABI library: {e}', file=sys.stderr, flush=True)
    cmd = [sys.executable, os.path.abspath(__file__), '--child'] + sys.argv[1:]
    last_err = 'no attempt made'
    # one state file per job: under torchrun every rank's supervisor is a child of the same agent process
    job = os.getppid() if 'LOCAL_RANK' in os.environ else os.getpid()
    state_path = os.path.join(os.environ.get('TMPDIR', '/tmp'), f'ldb200_bench_{job}.json')
    if rank == 0 and os.path.exists(state_path):
        os.remove(state_path)
    t_begin = time.time()
    for attempt in range(ATTEMPTS):
        if attempt and time.time() - t_begin > BUDGET_S:
            print(f'[bench supervisor {rank}] time budget of {BUDGET_S} s used up', file=sys.stderr, flush=True)
            break
        env = dict(os.environ, LDB200_BENCH_ATTEMPT=str(attempt), LDB200_BENCH_STATE=state_path)
        if attempt and 'MASTER_PORT' in env:
            # a fresh rendezvous for the restarted children: their own TCPStore (rank 0's child hosts it) on another
            # port, so that no key of the abandoned attempt (NCCL unique id, rank registrations) is read again
            env['MASTER_PORT'] = str(int(env['MASTER_PORT']) + 17 * attempt)
            env['TORCHELASTIC_USE_AGENT_STORE'] = 'False'
        proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
        state = {'t': time.time(), 'out': [], 'seen': False}

        def pump_err():
            for line in proc.stderr:
                state['t'] = time.time()
                state['seen'] = state['seen'] or line.startswith('[bench rank')
                sys.stderr.write(line)
                sys.stderr.flush()

        def pump_out():
            for line in proc.stdout:
                state['t'] = time.time()
                state['out'].append(line)

        th = [threading.Thread(target=pump_err, daemon=True), threading.Thread(target=pump_out, daemon=True)]
        for t in th:
            t.start()
        stalled = False
        while proc.poll() is None:
            time.sleep(1.0)
            # until the child's first progress line (imports, CUDA context, NCCL set-up on a fresh box) be patient
            limit = SUPERVISOR_STALL_S if state['seen'] else max(SUPERVISOR_STALL_S, START_S)
            if time.time() - state['t'] > limit:
                stalled = True
                print(f'[bench supervisor {rank}] attempt {attempt}: no progress for {limit} s, killing pid {proc.pid}',
                      file=sys.stderr, flush=True)
                try:
                    os.kill(proc.pid, signal.SIGKILL)
                except ProcessLookupError:
                    pass
                proc.wait()
        for t in th:
            t.join(timeout=5.0)
        line = next((l for l in reversed(state['out']) if l.lstrip().startswith('{')), None)
        if not stalled and proc.returncode == 0 and (line is not None or rank != 0):
            if line is not None:
                if attempt:
                    d = json.loads(line)
                    d['supervisor'] = dict(retries=attempt, reason=last_err,
                                           note='phases completed before a stall were kept, the stalled phase was measured again')
                    line = json.dumps(d) + '\n'
                sys.stdout.write(line)
                sys.stdout.flush()
            if rank == 0 and os.path.exists(state_path):
                os.remove(state_path)
            return
        last_err = 'stalled' if stalled else f'exit code {proc.returncode}'
        print(f'[bench supervisor {rank}] attempt {attempt} failed: {last_err}', file=sys.stderr, flush=True)
        if not stalled and proc.returncode not in (17, -9, -15):     # a real error: repeating it would not help
            sys.exit(proc.returncode if proc.returncode and proc.returncode > 0 else 1)
        time.sleep(2.0)       # let the other ranks' children go away before the next rendezvous
    sys.exit(1)


def profile_numbers():
    """Numbers that only a profiler can give, read from the COMMITTED ncu summaries (never typed into the source):
    DRAM traffic per launch of the roofline kernel and its share of the step's kernel time."""
    import glob
    import re
    out = {}
    prof = os.path.join(ROOT, 'profiles')
    for f in sorted(glob.glob(os.path.join(prof, 'r*_conv2_head_p3_summary.txt')), reverse=True):
        txt = open(f).read()
        rd = re.search(r'dram__bytes_read\.sum = ([0-9.]+) (\w+)', txt)
        wr = re.search(r'dram__bytes_write\.sum = ([0-9.]+) (\w+)', txt)
        if rd and wr:
            mul = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}
            out['traffic'] = float(rd.group(1)) * mul[rd.group(2)] + float(wr.group(1)) * mul[wr.group(2)]
            out['traffic_source'] = os.path.relpath(f, ROOT)
            break
    for f in sorted(glob.glob(os.path.join(prof, 'r*_launches*share.json')), reverse=True):
        try:
            d = json.load(open(f))
            out['share'] = d['conv2_tc_kernel']['share']
            out['share_source'] = os.path.relpath(f, ROOT)
            break
        except Exception:
            pass
    return out


if __name__ == '__main__':
    main()
