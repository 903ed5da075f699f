"""Stall samples per CUDA source line from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` output."""
import collections
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr = None
    per, cur, tot = collections.OrderedDict(), None, 0
    for r in rows:
        if r and r[0] == 'Line No':
            hdr = r
            iS, iI = hdr.index('# Samples'), hdr.index('Instructions Executed')
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0].strip().isdigit():
            cur = (int(r[0]), r[1].strip()[:110])
            per.setdefault(cur, [0, 0])
        try:
            s, n = int(r[iS] or 0), int(r[iI] or 0)
        except ValueError:
            s = n = 0
        if cur and r[2].strip():
            per[cur][0] += s
            per[cur][1] += n
            tot += s
    print("total samples", tot)
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%6d %6.2f%% inst=%10d  L%-4d %s" % (v[0], 100 * v[0] / max(tot, 1), v[1], k[0], k[1]))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
