"""Aggregate ncu stall samples per CUDA source line (joins the ncu SASS page with nvdisasm line info).
    python tools/ncu_lines.py REP.ncu-rep CUBIN 'kernel-substring' [top-n]
The cubin comes from `cuobjdump -xelf all libsga_b200.so`.  NCU_FILTER='-k regex:name -c 1' selects the kernel in a multi-kernel report."""
import collections, csv, io, re, subprocess, sys
rep, cubin, pat = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(['nvdisasm', '--print-line-info-inline', cubin], capture_output=True, text=True).stdout.splitlines()
# locate the function's text section
start = next(i for i, l in enumerate(dis) if l.strip().startswith('.section') and '.text.' in l and pat in l)
ins = []   # (line-key, sass)
cur = None
inl = []
for l in dis[start + 1:]:
    s = l.strip()
    if s.startswith('.section'):
        break
    m = re.match(r'//## File "([^"]+)", line (\d+)(.*)', s)
    if m:
        f = m.group(1).split('/')[-1]
        if 'inlined at' in m.group(3):
            inl.append(f'{f}:{m.group(2)}')
        else:
            cur = f'{f}:{m.group(2)}'
            inl = []
        continue
    if re.match(r'/\*[0-9a-f]{4,}\*/', s):
        ins.append((cur, tuple(inl), s))
import os
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'] + os.environ.get('NCU_FILTER', '').split(), capture_output=True, text=True).stdout.splitlines()
h = next(i for i, l in enumerate(out) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO('\n'.join(out[h:]))))
print('sass instrs: nvdisasm', len(ins), 'ncu', len(rows))
agg = collections.Counter()
agg_top = collections.Counter()
ex = collections.Counter()
tot = 0
for (key, inl, s), r in zip(ins, rows):
    n = float(r['# Samples'] or 0)
    tot += n
    agg[key] += n
    ex[key] += float(r['Instructions Executed'] or 0)
print('total samples', tot)
src = {}
def line_text(key):
    f, ln = key.split(':')
    for d in ('sgaligner_b200/csrc/',):
        try:
            return open(d + f).read().splitlines()[int(ln) - 1].strip()[:80]
        except Exception:
            pass
    return ''
for key, n in agg.most_common(topn):
    print(f'{n:8.0f} {100 * n / tot:5.1f}%  exec={ex[key]:11.0f}  {key:24s} {line_text(key)}')
