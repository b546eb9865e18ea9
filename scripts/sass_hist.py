"""Opcode histogram per kernel from `cuobjdump -sass` (static counts; loops count once)."""
import re, subprocess, sys, collections
so = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ''
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
cur = None; hist = collections.defaultdict(collections.Counter)
for line in txt.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m: cur = m.group(1); continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)', line)
    if m and cur: hist[cur][m.group(1)] += 1
for fn, h in hist.items():
    if pat and not re.search(pat, fn): continue
    tot = sum(h.values())
    print(f'== {fn[:110]}  total {tot}')
    print('   ' + '  '.join(f'{k}:{v}' for k, v in h.most_common(22)))
