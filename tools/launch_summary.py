"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel.
    python tools/launch_summary.py gpurun_out/launches.csv [kernel-substring-to-list]"""
import collections, csv, re, sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
tot, cnt, listed = collections.Counter(), collections.Counter(), []
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', row['Kernel Name'])[:72]
    v = float(row['Metric Value'].replace(',', ''))
    v = v / 1e3 if row['Metric Unit'] == 'ns' else (v * 1e3 if row['Metric Unit'] == 'ms' else v)
    tot[name] += v
    cnt[name] += 1
    if len(sys.argv) > 2 and sys.argv[2] in row['Kernel Name']:
        listed.append(round(v))
T = sum(tot.values())
print(f'total {T / 1e3:.2f} ms of kernel time, {sum(cnt.values())} launches')
for k, v in tot.most_common(24):
    print(f'{v:9.0f} us {100 * v / T:5.1f}% x{cnt[k]:4d} {k}')
if listed:
    print(listed)
