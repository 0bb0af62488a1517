"""Per-level kernel times of the last batch build in an ncu launch list (gpu__time_duration csv)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; out = []
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r)); out.append((d['Kernel Name'].split('(')[0][:30], float(d['Metric Value']) / 1000))
s = [i for i, o in enumerate(out) if 'k_project<1024, 4, 1' in o[0]][-1]
for o in out[s:s + 100]:
    if o[0].startswith('k_top_hist'): print()
    print("%s:%.1f" % (o[0].replace('k_top_', ''), o[1]), end='  ')
print()
