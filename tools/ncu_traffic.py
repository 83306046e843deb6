"""Extract per-launch DRAM traffic and the headline metrics of captured kernels from an `ncu --set full` report.

    python tools/ncu_traffic.py gpurun_out/x.ncu-rep [--kernel REGEX] [--json OUT --algorithmic-bytes N --label TEXT]

Prints one line per captured launch (duration, DRAM read / write bytes, tensor-pipe and DRAM utilisation, registers, grid) and,
with --json, writes {kernel, dram_bytes_per_launch, algorithmic_bytes, source} of the FIRST matching launch -- the file bench.py
reads for `roofline.traffic` (a profiler cannot run inside the timed region, so the number is tied to the committed capture)."""
import argparse, csv, io, json, re, subprocess, sys

ap = argparse.ArgumentParser()
ap.add_argument('report')
ap.add_argument('--kernel', default='.')
ap.add_argument('--json')
ap.add_argument('--algorithmic-bytes', type=float, default=None)
ap.add_argument('--label', default='')
a = ap.parse_args()
raw = subprocess.run(['ncu', '-i', a.report, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
def f(r, k, default=float('nan')):
    try:
        return float(r[col[k]].replace(',', ''))
    except Exception:
        return default
def scale(k):
    u = units[col[k]] if k in col else ''
    return {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1.0}.get(u, 1.0)
first = None
for r in data:
    name = r[col['Kernel Name']]
    if not re.search(a.kernel, name):
        continue
    rd, wr = f(r, 'dram__bytes_read.sum') * scale('dram__bytes_read.sum'), f(r, 'dram__bytes_write.sum') * scale('dram__bytes_write.sum')
    rec = dict(kernel=name.split('(')[0], us=f(r, 'gpu__time_duration.sum'), dram_read=rd, dram_write=wr,
               tensor_pct=f(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
               dram_pct=f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
               regs=f(r, 'launch__registers_per_thread'), grid=r[col['launch__grid_size']] if 'launch__grid_size' in col else '')
    print(json.dumps(rec))
    if first is None:
        first = rec
if a.json and first:
    json.dump({'kernel': first['kernel'], 'dram_bytes_per_launch': first['dram_read'] + first['dram_write'],
               'dram_read': first['dram_read'], 'dram_write': first['dram_write'], 'us_under_ncu': first['us'],
               'algorithmic_bytes': a.algorithmic_bytes, 'source': f'{a.label or a.report}: ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum'},
              open(a.json, 'w'), indent=1)
