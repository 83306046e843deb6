"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA), UBLKCP, plus HMMA (legacy mma.sync) -- from `cuobjdump -sass` of the
shipped library.    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'vqvae_vqgan_pytorch_lightning_b200', 'libvqgan_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
pat = {'UTCHMMA': r'\bUTC\w*MMA', 'UTCHMMA.2CTA': r'UTC\w*MMA\.2CTA', 'LDTM': r'\bLDTM', 'STTM': r'\bSTTM', 'UTMALDG': r'\bUTMALDG', 'UTMASTG': r'\bUTMASTG',
       'UBLKCP': r'\bUBLKCP', 'HMMA': r'\bHMMA', 'REDG.F32x4': r'REDG\.E\.ADD\.F32x4', 'LDGSTS': r'\bLDGSTS'}
cur, counts, arch = None, collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    m = re.search(r'arch = (sm_\w+)', line)
    if m: arch = m.group(1)
    if cur:
        for k, p in pat.items():
            if re.search(p, line): counts[cur][k] += 1
print(f'# cuobjdump -sass {os.path.basename(lib)} ({arch}): kernels with tensor-core / TMA / TMEM instructions')
print(f'{"kernel":70s} ' + ' '.join(f'{k:>12s}' for k in pat))
tot = collections.Counter()
for fn, c in counts.items():
    if any(c[k] for k in ('UTCHMMA', 'LDTM', 'UTMALDG', 'HMMA', 'UBLKCP')):
        name = subprocess.run(['c++filt', fn], capture_output=True, text=True).stdout.strip().replace('(anonymous namespace)::', '').replace('void ', '').split('(')[0][-70:]
        print(f'{name:70s} ' + ' '.join(f'{c[k]:12d}' for k in pat))
    tot.update(c)
print(f'{"TOTAL (all kernels)":70s} ' + ' '.join(f'{tot[k]:12d}' for k in pat))
