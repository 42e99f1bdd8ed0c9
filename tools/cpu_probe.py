import os, time, torch
print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
for f in ('/sys/fs/cgroup/cpu.max', '/sys/fs/cgroup/cpu/cpu.cfs_quota_us'):
    try: print(f, open(f).read().strip())
    except Exception as e: pass
print('torch default threads', torch.get_num_threads())
x = torch.randn(2, 256, 100, 168); w = torch.randn(256, 256, 3, 3)
for t in (8, 16, 32, 64, 128):
    torch.set_num_threads(t)
    torch.nn.functional.conv2d(x, w, padding=1)
    t0 = time.perf_counter()
    for _ in range(3): torch.nn.functional.conv2d(x, w, padding=1)
    dt = (time.perf_counter() - t0) / 3
    print(f'threads {t:4d}: conv {dt*1e3:8.1f} ms  {2*2*100*168*256*256*9/dt/1e12:.2f} TFLOP/s')
