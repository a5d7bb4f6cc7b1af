"""usage: python scratch/analyze_prof.py <ncu-rep> <libncb200.so copy> [kernel-substring]"""
import re, csv, collections, subprocess, sys, os, tempfile
rep, so = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else 'tile_kernelILi3ELi0ELi512ELi2'
td = tempfile.mkdtemp()
subprocess.run(f"cd {td} && cuobjdump -xelf all {so} >/dev/null 2>&1 && (for f in *.cubin; do nvdisasm -g -c $f; done) > all.sass 2>/dev/null", shell=True)
subprocess.run(f"ncu -i {rep} --page source --csv --kernel-name regex:$KREGEX > {td}/src.csv 2>/dev/null; ncu -i {rep} --page raw --csv > {td}/raw.csv 2>/dev/null", shell=True)
rows = list(csv.reader(open(f'{td}/raw.csv'))); hdr = rows[0]
for w in ('gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__grid_size','launch__block_size'):
    if w in hdr:
        i = hdr.index(w); print(f"{w:70s}", [r[i][:10] for r in rows[2:]])
for w in hdr:
    if 'issue_stalled' in w and 'per_issue_active' in w:
        i = hdr.index(w); v = [float(r[i]) for r in rows[2:]]
        if max(v) > 0.3: print(f"  stall {w.split('stalled_')[1][:28]:30s}", v)
txt = open(f'{td}/all.sass').read().split('\n')
start = None
for i, l in enumerate(txt):
    if l.startswith('.text.') and kern in l: start = i
ins = []; cur = ('?', 0)
for l in txt[start+1:]:
    if l.startswith('.text.') or l.startswith('.section'):
        if ins: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: ins.append((cur, m.group(2).strip()))
rows = list(csv.reader(open(f'{td}/src.csv')))
hidx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
LI = int(os.environ.get('LAUNCH', '0'))
h = rows[hidx[LI]]; end = hidx[LI+1]-1 if len(hidx) > LI+1 else len(rows)
body = [r for r in rows[hidx[LI]+1:end] if len(r) > 6]
iS = h.index('# Samples'); iE = h.index('Instructions Executed'); iSrc = h.index('Source')
print('sass', len(ins), 'ncu', len(body))
n = min(len(ins), len(body))
byl = collections.Counter(); bye = collections.Counter()
for k in range(n):
    byl[ins[k][0]] += int(body[k][iS]); bye[ins[k][0]] += int(body[k][iE])
ts = sum(byl.values()); te = sum(bye.values())
src = {f: open(os.path.join(os.path.dirname(__file__), '..', 'noisycircuits_b200', 'csrc', f)).read().split('\n') for f in ('ncb_common.h', 'ncb_kernels.cuh')}
for (f, l), c in byl.most_common(30):
    line = src[f][l-1].strip()[:64] if f in src and l-1 < len(src[f]) else ''
    print(f"{f[:12]:12s}:{l:4d} smp {c/ts*100:5.1f}% exe {bye[(f,l)]/te*100:5.1f}%  {line}")
mix = collections.Counter()
for k in range(n):
    t = body[k][iSrc].split(); op = t[1] if t[0].startswith('@') else t[0]
    mix[op.split('.')[0]] += int(body[k][iE])
print([(o, round(c/te*100, 1)) for o, c in mix.most_common(16)])
