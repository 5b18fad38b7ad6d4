"""Aggregate an ncu SASS source-page CSV by CUDA source line using nvdisasm -g line info.
usage: ncu_lines.py <src_sass.csv> <nvdisasm -g -c output> <kernel substring> [top]"""
import csv, re, sys
from collections import defaultdict
csvf, disf, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
# address -> line from nvdisasm
addr2line = {}
cur = None; inside = False
for l in open(disf, errors='ignore'):
    if l.startswith('//---') and '.text.' in l:
        inside = kern in l
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)), 'inlined' in m.group(3)); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m and cur: addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(csvf)))
hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
h = rows[hi]; ia = h.index('Address'); ie = h.index('Instructions Executed'); isamp = h.index('# Samples')
base = None
agg = defaultdict(lambda: [0.0, 0.0])
tot = 0; tots = 0
for r in rows[hi + 1:]:
    try:
        a = int(r[ia], 16); n = float(r[ie]); s = float(r[isamp] or 0)
    except Exception:
        continue
    if base is None: base = a
    key = addr2line.get(a - base, ('?', 0, False))
    agg[key[:2]][0] += n; agg[key[:2]][1] += s; tot += n; tots += s
print('total warp instructions %.3e, samples %d' % (tot, tots))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%6.2f%% inst  %6.2f%% samples  %s:%d' % (100 * v[0] / tot, 100 * v[1] / max(tots, 1), k[0], k[1]))

# ---- per-function aggregation (function = nearest preceding 'AUGB_D|AUGB_HD ... name(' line in the file)
import os
srcdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'augustus_b200', 'csrc')
fn_of = {}
for f in os.listdir(srcdir):
    cur = '?'
    for i, l in enumerate(open(os.path.join(srcdir, f), errors='ignore'), 1):
        m = re.match(r'\s*(?:template.*>\s*)?(?:AUGB_HD|AUGB_D|__device__|__global__|static|inline)[^;=]*?\b(\w+)\s*\([^;]*\)\s*(?:const)?\s*\{', l)
        if m: cur = m.group(1)
        fn_of[(f, i)] = cur
fagg = defaultdict(lambda: [0.0, 0.0])
for k, v in agg.items():
    fn = fn_of.get(k, k[0])
    fagg[fn][0] += v[0]; fagg[fn][1] += v[1]
print('--- by function')
for k, v in sorted(fagg.items(), key=lambda kv: -kv[1][0])[:30]:
    print('%6.2f%% inst  %6.2f%% samples  %s' % (100 * v[0] / tot, 100 * v[1] / max(tots, 1), k))
