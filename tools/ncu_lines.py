#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals for one kernel of an
.ncu-rep, by joining ncu's SASS page with nvdisasm's line table.

  python tools/ncu_lines.py <report.ncu-rep> <cubin> <kernel-substring> [top]
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def line_table(cubin, kernel):
    """offset -> source line for the first function whose name contains `kernel`."""
    out = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    table, cur, active = {}, None, False
    for ln in out.splitlines():
        m = re.match(r'\s*\.text\.(\S+):', ln)
        if m:
            active = kernel in m.group(1)
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', ln)
        if m and active:
            table[int(m.group(1), 16)] = cur
    return table


def main():
    rep, cubin, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # first kernel block only
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    h = rows[hdr]
    ie, ns, te = h.index('Instructions Executed'), h.index('# Samples'), h.index('Thread Instructions Executed')
    table = line_table(cubin, kernel)
    base = None
    agg = defaultdict(lambda: [0, 0, 0])
    for r in rows[hdr + 1:]:
        if not r or not r[0].startswith('0x'):
            break
        addr = int(r[0], 16)
        base = addr if base is None else base
        key = table.get(addr - base, ('?', 0))
        a = agg[key]
        a[0] += int(r[ie]); a[1] += int(r[ns]); a[2] += int(r[te])
    ti = sum(a[0] for a in agg.values()); tsamp = sum(a[1] for a in agg.values())
    print(f'total warp-instructions {ti}, samples {tsamp}')
    src = {}
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        if f not in src:
            try:
                src[f] = open(f'/root/repo/self_play_racing_b200/csrc/{f}').read().splitlines()
            except OSError:
                src[f] = []
        text = src[f][l - 1].strip() if 0 < l <= len(src[f]) else ''
        lanes = a[2] / a[0] if a[0] else 0
        print(f'{a[0] / ti * 100:5.1f}% inst {a[1] / tsamp * 100:5.1f}% stall  lanes {lanes:4.1f}  {f}:{l}: {text[:100]}')


if __name__ == '__main__':
    main()
