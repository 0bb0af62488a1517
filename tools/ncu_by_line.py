"""Join an `ncu --page source --csv` dump (SASS rows: samples, instructions executed, shared wavefronts) with the line info of
`nvdisasm -g <cubin>` and aggregate per CUDA source line.  usage: ncu_by_line.py <source.csv> <nvdisasm -g output> <mangled kernel> [top]"""
import csv
import re
import sys


def line_map(dis, kernel):
    m, cur, on = [], None, False
    for ln in open(dis):
        if ln.startswith('\t.section'):
            on = ('.text.' + kernel + ',') in ln
            continue
        if not on:
            continue
        g = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if g:
            cur = (g.group(1).split('/')[-1], int(g.group(2)), 'inlined' in g.group(3))
            continue
        g = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if g:
            m.append((int(g.group(1), 16), cur, g.group(2).strip()))
    return m


def main(src, dis, kernel, top=40):
    lm = line_map(dis, kernel)
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    iS, iEx, iSrc = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
    iW = hdr.index('L1 Wavefronts Shared')
    data = [r for r in rows[2:] if len(r) > iW]
    assert len(data) == len(lm), (len(data), len(lm))
    agg = {}
    tot_s = tot_e = 0
    for r, (off, cur, txt) in zip(data, lm):
        s, e, w = int(r[iS] or 0), int(r[iEx] or 0), int(r[iW] or 0)
        a = agg.setdefault(cur[:2] if cur else None, [0, 0, 0, 0])
        a[0] += s; a[1] += e; a[2] += w; a[3] += 1
        tot_s += s; tot_e += e
    print('total samples %d, warp instructions executed %d, SASS rows %d' % (tot_s, tot_e, len(data)))
    print('%-22s %8s %6s %12s %6s %10s %5s' % ('line', 'samples', '%', 'inst_exec', '%', 'smem_wf', 'sass'))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print('%-22s %8d %5.1f%% %12d %5.1f%% %10d %5d' % ('%s:%d' % k if k else '?', a[0], 100.0 * a[0] / tot_s, a[1], 100.0 * a[1] / tot_e, a[2], a[3]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)


def regions(src, dis, kernel, spec):
    """spec: name=lo-hi,... (build.cu line ranges); everything else by file"""
    lm = line_map(dis, kernel)
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    iS, iEx = hdr.index('# Samples'), hdr.index('Instructions Executed')
    iW = hdr.index('L1 Wavefronts Shared')
    data = [r for r in rows[2:] if len(r) > iW]
    rs = []
    for part in spec.split(','):
        nm, rg = part.split('=')
        lo, hi = rg.split('-')
        rs.append((nm, int(lo), int(hi)))
    agg = {}
    ts = te = 0
    for r, (off, cur, txt) in zip(data, lm):
        s, e, w = int(r[iS] or 0), int(r[iEx] or 0), int(r[iW] or 0)
        key = cur[0] if cur else '?'
        if cur and cur[0].endswith('.cu'):
            for nm, lo, hi in rs:
                if lo <= cur[1] <= hi:
                    key = nm
                    break
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += s; a[1] += e; a[2] += w
        ts += s; te += e
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print('%-26s samples %6.1f%%  inst %6.1f%%  smem_wf %10d' % (k, 100.0 * a[0] / ts, 100.0 * a[1] / te, a[2]))
