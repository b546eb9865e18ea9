"""Per-phase stall samples of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name regex:<k>`:
instructions are grouped between barriers (BAR / DEPBAR / LDGDEPBAR) = the phases of the pipelined kernels."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'data': []}
        blocks.append(cur)
    elif cur is not None and cur['hdr'] is None and r and r[0] == 'Address':
        cur['hdr'] = r
    elif cur is not None and cur['hdr'] is not None and len(r) >= len(cur['hdr']) - 2:
        cur['data'].append(r)
def op(src):
    s = src.strip()
    if s.startswith('@'):
        s = s.split(None, 1)[1]
    return s.split()[0]
for b in blocks[:1]:
    ix = {h: i for i, h in enumerate(b['hdr'])}
    data = [r for r in b['data'] if r[ix['# Samples']].isdigit()]
    alls = sum(int(r[ix['# Samples']]) for r in data)
    alli = sum(int(r[ix['Instructions Executed']]) for r in data)
    print(b['name'][:100], 'samples', alls, 'warp-inst', alli)
    cur = dict(s=0, i=0, first=None)
    for r in data:
        o = op(r[ix['Source']])
        cur['s'] += int(r[ix['# Samples']]); cur['i'] += int(r[ix['Instructions Executed']])
        if cur['first'] is None: cur['first'] = r[ix['Address']][-5:]
        if o.startswith(('BAR', 'DEPBAR', 'LDGDEPBAR')):
            print(f"  {cur['first']} .. {o:24s} samples {cur['s']:6d} ({100*cur['s']/alls:5.1f}%)  warp-inst {cur['i']:10d} ({100*cur['i']/alli:5.1f}%)")
            cur = dict(s=0, i=0, first=None)
    print(f"  {cur['first']} .. end{'':21s} samples {cur['s']:6d} ({100*cur['s']/alls:5.1f}%)  warp-inst {cur['i']:10d} ({100*cur['i']/alli:5.1f}%)")
    print('  top instructions by samples:')
    for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
        print('   ', r[ix['# Samples']].rjust(6), r[ix['Instructions Executed']].rjust(9), r[ix['Source']].strip()[:100])
