"""Hot instruction footprint of a kernel from an ncu SASS source page (+ nvdisasm -g -c for function attribution).
usage: hotset.py <src_sass.csv> <nvdisasm output> <kernel mangled-name substring e.g. 7k_sweepE>"""
import csv, re, sys, os
from collections import defaultdict
csvf, disf, kern = sys.argv[1:4]
srcdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'augustus_b200', 'csrc')
fn_of = {}
for f in os.listdir(srcdir):
    cur = '?'
    for i, l in enumerate(open(os.path.join(srcdir, f), errors='ignore'), 1):
        m = re.match(r'\s*(?:template.*>\s*)?(?:AUGB_HDN|AUGB_HD|AUGB_DN|AUGB_EMIT|AUGB_D|__device__|__global__|static|inline)[^;=]*?\b(\w+)\s*\([^;]*\)\s*(?:const)?\s*\{', l)
        if m: cur = m.group(1)
        fn_of[(f, i)] = cur
addr2fn = {}; inside = False; cur = None
for l in open(disf, errors='ignore'):
    if l.startswith('//---') and '.text.' in l:
        inside = kern in l; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m: addr2fn[int(m.group(1), 16)] = fn_of.get(cur, cur[0] if cur else '?')
rows = list(csv.reader(open(csvf)))
hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
h = rows[hi]; ia = h.index('Address'); ie = h.index('Instructions Executed')
ins = []; base = None
for r in rows[hi + 1:]:
    try: a = int(r[ia], 16); n = float(r[ie])
    except Exception: continue
    if base is None: base = a
    ins.append((a - base, n))
tot = sum(n for _, n in ins)
print('static %d instr (%.1f KB), dynamic %.3e' % (len(ins), len(ins) * 16 / 1024, tot))
srt = sorted(ins, key=lambda x: -x[1])
acc = 0; marks = {0.5: None, 0.8: None, 0.9: None, 0.95: None, 0.99: None, 0.999: None}
for i, (a, n) in enumerate(srt):
    acc += n
    for k in marks:
        if marks[k] is None and acc >= k * tot: marks[k] = i + 1
for k in sorted(marks): print('  %.1f%% of dynamic instr from %d static (%.1f KB)' % (100 * k, marks[k], marks[k] * 16 / 1024))
# 128-byte lines touched, weighted
lines = defaultdict(float)
for a, n in ins: lines[a // 128] += n
ls = sorted(lines.values(), reverse=True); acc = 0
for frac in (0.9, 0.95, 0.99):
    acc = 0
    for i, v in enumerate(ls):
        acc += v
        if acc >= frac * tot: print('  %.0f%% from %d lines of 128 B (%.1f KB)' % (100 * frac, i + 1, (i + 1) / 8)); break
# per function: static count of instrs within the 99% hot set, dynamic share
thr = srt[marks[0.99] - 1][1]
fs = defaultdict(lambda: [0, 0, 0.0])
for a, n in ins:
    f = addr2fn.get(a, '?'); fs[f][0] += 1; fs[f][2] += n
    if n >= thr: fs[f][1] += 1
print('function: static / hot-static(99%%) / dynamic%%')
for f, v in sorted(fs.items(), key=lambda kv: -kv[1][1])[:45]:
    print('%5d %5d %6.2f%%  %s' % (v[0], v[1], 100 * v[2] / tot, f))
