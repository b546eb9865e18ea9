"""Static opcode counts of one kernel split at barriers (BAR / DEPBAR / LDGDEPBAR): which phase holds the instructions.
usage: python scripts/sass_segments.py <lib.so> <mangled-name-regex>"""
import re, subprocess, sys, collections
so, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
cur = None; seg = collections.Counter(); on = False
def flush(tag):
    global seg
    if sum(seg.values()): print(f'  {tag:28s} {sum(seg.values()):5d}  ' + ' '.join(f'{k}:{v}' for k, v in seg.most_common(14)))
    seg = collections.Counter()
for line in txt.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        if on: flush('end')
        on = re.search(pat, m.group(1)) is not None
        if on: print('==', m.group(1)[:120])
        continue
    if not on: continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)', line)
    if not m: continue
    op = m.group(2); seg[op.split('.')[0]] += 1
    if op.startswith(('BAR', 'DEPBAR', 'LDGDEPBAR')): flush(f'{m.group(1)} {op[:18]}')
if on: flush('end')
