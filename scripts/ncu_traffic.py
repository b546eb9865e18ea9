"""raw ncu csv (ncu -i X.ncu-rep --page raw --csv) -> per-kernel DRAM traffic JSON used by bench.py's roofline.traffic.
usage: python scripts/ncu_traffic.py raw.csv out.json "<note>" [library build id] """
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
unit = {h: rows[1][i] for i, h in enumerate(hdr)}
def val(r, k):
    v = float(r[ix[k]].replace(',', ''))
    u = unit[k].lower()
    return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'msecond': 1e3, 'nsecond': 1e-3}.get(u, 1)
out, total = {}, 0.0
for r in data:
    name = r[ix['Kernel Name']]
    if name in out:
        continue
    rd, wr = val(r, 'dram__bytes_read.sum'), val(r, 'dram__bytes_write.sum')
    out[name] = {'dram_bytes_read': rd, 'dram_bytes_write': wr, 'gpu_time_us': val(r, 'gpu__time_duration.sum')}
    total += rd + wr
out['_whole_path_dram_bytes'] = total
out['_note'] = sys.argv[3] if len(sys.argv) > 3 else ''
if len(sys.argv) > 4:
    out['build_id'] = sys.argv[4].strip()  # bench.py reports roofline.traffic only from a capture of the library it runs
json.dump(out, open(sys.argv[2], 'w'), indent=1)
print(json.dumps(out, indent=1))
