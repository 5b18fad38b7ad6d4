"""Static SASS footprint of a kernel by source function (nvdisasm -g -c line table).
usage: sass_static.py <nvdisasm -g -c output> <kernel substring>"""
import re, sys, os
from collections import defaultdict
disf, kern = sys.argv[1:3]
srcdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'augustus_b200', 'csrc')
fn_of = {}
for f in os.listdir(srcdir):
    cur = '?'
    for i, l in enumerate(open(os.path.join(srcdir, f), errors='ignore'), 1):
        m = re.match(r'\s*(?:template.*>\s*)?(?:AUGB_HD|AUGB_DN|AUGB_D|__device__|__global__|static|inline)[^;=]*?\b(\w+)\s*\([^;]*\)\s*(?:const)?\s*\{', l)
        if m: cur = m.group(1)
        fn_of[(f, i)] = cur
inside = False; cur = None
cnt = defaultdict(int); lines = defaultdict(int); tot = 0
for l in open(disf, errors='ignore'):
    if l.startswith('//---') and '.text.' in l:
        inside = kern in l and (kern + 'E') in l or l.strip().endswith(kern)
        inside = re.search(r'\d+' + kern + 'E', l) is not None
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/', l):
        tot += 1
        k = fn_of.get(cur, cur[0] if cur else '?')
        cnt[k] += 1; lines[cur] += 1
print('total static instructions', tot, '= %.1f KB' % (tot * 16 / 1024))
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:40]:
    print('%6d  %5.1f%%  %s' % (v, 100.0 * v / tot, k))
