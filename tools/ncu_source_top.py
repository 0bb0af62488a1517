"""Top stall sites of a kernel from an .ncu-rep source page (SASS level), grouped with their dominant stall reasons."""
import csv
import subprocess
import sys


def main(path, top=30, kernel=None):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'] + (['-k', kernel] if kernel else []),
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    iS, iSrc, iEx = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    data = []
    for r in rows[2:]:
        try:
            data.append((int(r[iS]), r))
        except (ValueError, IndexError):
            continue
    tot = sum(s for s, _ in data)
    print('total samples', tot)
    agg = {}
    for s, r in data:
        for c in stall_cols:
            if r[c] not in ('', '0'):
                agg[hdr[c]] = agg.get(hdr[c], 0) + int(r[c])
    print('by reason:', sorted(((v, k) for k, v in agg.items()), reverse=True)[:8])
    idx = sorted(range(len(data)), key=lambda i: -data[i][0])[:top]
    for i in sorted(idx):
        s, r = data[i]
        st = sorted([(int(r[c]), hdr[c]) for c in stall_cols if r[c] not in ('', '0')], reverse=True)[:2]
        print('%5d %7d %5.1f%% %10s  %-60s %s' % (i, s, 100.0 * s / tot, r[iEx], r[iSrc].strip()[:60], st))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30, sys.argv[3] if len(sys.argv) > 3 else None)
