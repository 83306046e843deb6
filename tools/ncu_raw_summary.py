"""Summarise `ncu -i X.ncu-rep --page raw --csv` exports (one row per captured launch): duration, DRAM bytes, tensor-pipe / DRAM / L2
utilisation, achieved occupancy, registers, grid -- the numbers DESIGN.md and bench.py quote.

    python tools/ncu_raw_summary.py gpurun_out/r02_full_*_raw.csv [--json profiles/r02_ncu_full_summary.json]"""
import csv, json, re, sys
args = [a for a in sys.argv[1:] if not a.startswith('--')]
out_json = sys.argv[sys.argv.index('--json') + 1] if '--json' in sys.argv else None
if out_json in args: args.remove(out_json)
KEYS = {
    'us': 'gpu__time_duration.sum', 'dram_read': 'dram__bytes_read.sum', 'dram_write': 'dram__bytes_write.sum',
    'tensor_pct': 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'tensor_subpipe_pct': 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
    'dram_pct': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l2_pct': 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1_pct': 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_busy_pct': 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'issue_active_pct': 'sm__inst_issued.avg.pct_of_peak_sustained_active', 'regs': 'launch__registers_per_thread',
    'grid': 'launch__grid_size', 'block': 'launch__block_size', 'smem_dyn': 'launch__shared_mem_per_block_dynamic',
    'occupancy_pct': 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm_clk_mhz': 'smsp__cycles_elapsed.avg.per_second',
    'stall_long_sb': 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'stall_barrier': 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'stall_short_sb': 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'stall_lg_throttle': 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'stall_mio': 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
}
UNIT = {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1.0, 'us': 1.0, 'ms': 1e3, 'ns': 1e-3, 's': 1e6, 'second': 1e6, 'usecond': 1.0, 'msecond': 1e3, 'nsecond': 1e-3}
allrec = []
for path in args:
    rows = list(csv.reader(l for l in open(path) if not l.startswith('==')))
    if len(rows) < 3:
        print(f'# {path}: empty'); continue
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f'# {path}')
    for r in data:
        rec = {'file': path.split('/')[-1], 'kernel': re.sub(r'\(.*', '', r[col['Kernel Name']])}
        for k, m in KEYS.items():
            if m in col:
                try:
                    v = float(r[col[m]].replace(',', ''))
                except ValueError:
                    continue
                u = units[col[m]]
                if k in ('us', 'dram_read', 'dram_write'):
                    v *= UNIT.get(u, 1.0)
                rec[k] = v
        allrec.append(rec)
        print('  ' + rec['kernel'][:44].ljust(44) + ' '.join(f'{k}={rec[k]:.4g}' for k in KEYS if k in rec))
if out_json:
    json.dump(allrec, open(out_json, 'w'), indent=1)
