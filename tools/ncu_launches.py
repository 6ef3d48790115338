#!/usr/bin/env python
"""Summarise an ncu --csv launch list (one row per kernel launch: time, DRAM bytes, L2 requests / sectors)."""
import csv
import sys

UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, 'ns': 1e-6, 'us': 1e-3, 'ms': 1, 's': 1e3,
        'usecond': 1e-3, 'msecond': 1, 'nsecond': 1e-6, 'second': 1e3}


def main(path, pattern=""):
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ix = {h: i for i, h in enumerate(hdr)}
    cur = {}
    for row in r:
        k = (int(row[ix['ID']]), row[ix['Kernel Name']])
        v = float(row[ix['Metric Value']].replace(',', ''))
        cur.setdefault(k, {})[row[ix['Metric Name']]] = v * UNIT.get(row[ix['Metric Unit']], 1)
    for (i, name), m in sorted(cur.items()):
        if pattern and pattern not in name:
            continue
        s = "%3d %-46s %9.3f ms" % (i, name[:46], m.get('gpu__time_duration.sum', 0))
        if 'dram__bytes_read.sum' in m:
            s += "  rd %7.2f GB wr %7.2f GB" % (m['dram__bytes_read.sum'] / 1e9, m.get('dram__bytes_write.sum', 0) / 1e9)
        if 'lts__t_requests.sum' in m:
            s += "  L2 req %8.1f M sect %8.1f M" % (m['lts__t_requests.sum'] / 1e6, m.get('lts__t_sectors.sum', 0) / 1e6)
        print(s)


if __name__ == "__main__":
    main(*sys.argv[1:])
