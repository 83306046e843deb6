"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total ms, share."""
import collections, csv, re, sys
path = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith('==')) if r]
hdr = rows[0]
ki, vi, mi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
ui = hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= vi or r[mi] != 'gpu__time_duration.sum':
        continue
    v = float(r[vi].replace(',', ''))
    scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[ui], 1e-6)
    name = re.sub(r'\(.*', '', r[ki]).replace('void ', '').replace('<unnamed>::', '')
    name = re.sub(r'at::native::.*?(\w+_kernel\w*).*', r'torch:\1', name)
    agg[name][0] += 1; agg[name][1] += v * scale
tot = sum(v[1] for v in agg.values())
print(f'# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.2f} ms (cold-cache, serialised: compare SHARES)')
print(f'{"kernel":70s} {"launches":>8s} {"ms":>9s} {"share":>7s}')
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{k[:70]:70s} {n:8d} {ms:9.3f} {100 * ms / tot:6.1f}%')
