"""Per-instruction hot spots from an ncu report's source page (SASS view).
    python tools/ncu_source.py REP [kernel-index] [top-n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
import os
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'] + os.environ.get('NCU_FILTER', '').split(), capture_output=True, text=True).stdout
lines = out.splitlines()
# the first line is the kernel name record; then header
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO('\n'.join(lines[start:]))))
def num(r, k):
    try:
        return float(r.get(k, '0') or 0)
    except ValueError:
        return 0.0
tot = sum(num(r, '# Samples') for r in rows)
print('instructions', len(rows), 'samples', tot)
print('--- top stall-sample instructions')
for idx, r in sorted(enumerate(rows), key=lambda x: -num(x[1], '# Samples'))[:topn]:
    print(f"{idx:5d} {num(r,'# Samples'):8.0f} {100*num(r,'# Samples')/max(tot,1):5.1f}%  exec={num(r,'Instructions Executed'):10.0f}  {r['Source'].strip()[:90]}")
print('--- shared-memory excessive wavefronts')
for idx, r in sorted(enumerate(rows), key=lambda x: -num(x[1], 'L1 Wavefronts Shared Excessive'))[:15]:
    if num(r, 'L1 Wavefronts Shared Excessive') > 0:
        print(f"{idx:5d} excess={num(r,'L1 Wavefronts Shared Excessive'):10.0f} total={num(r,'L1 Wavefronts Shared'):10.0f} ideal={num(r,'L1 Wavefronts Shared Ideal'):10.0f} {r['Source'].strip()[:80]}")
print('--- local memory')
for idx, r in enumerate(rows):
    s = r['Source']
    if ('LDL' in s or 'STL' in s) and num(r, 'Instructions Executed') > 1e5:
        print(f"{idx:5d} exec={num(r,'Instructions Executed'):10.0f} samples={num(r,'# Samples'):6.0f} {s.strip()[:80]}")
