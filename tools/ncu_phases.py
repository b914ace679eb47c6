#!/usr/bin/env python
"""Instruction / stall share per kernel phase: joins ncu's SASS page with the
cubin's line table and buckets rk_step.cu lines by the phase markers found in
the source.   python tools/ncu_phases.py <report> <cubin> <kernel-substring>"""
import csv, io, re, subprocess, sys
sys.path.insert(0, __import__('os').path.dirname(__file__))
import ncu_lines as nl

MARKERS = [(r'__forceinline__ double dmul', 'f64 helpers / shuffles'), (r'void argmin_exact', 'argmin_exact'),
           (r'double ray_segment\(', 'ray_segment f64'), (r'void raycast_walls_exact', 'walls exact (fallback)'),
           (r'double raycast_car_edges', 'car edges f64'), (r'void argmin_culled5_f64', 'argmin f64 slow path'), (r'float sweep_atan2', 'sweep atan2'), (r'bool argmin_lane5', 'argmin lane (grid)'), (r'RayHit grid_ray', 'grid ray DDA'), (r'// ---- candidate search, environment by environment', 'sweep: per-env dirs / keys / winners out'), (r'// ---- float64 distances: every lane finishes', 'lane-per-car ray loop: f64 winners, stores'),
           (r'void raycast_walls_culled', 'ray sweep setup'), (r'// ---- level 1 ----', 'ray level 1'),
           (r'// ---- level 2 ----', 'ray level 2'), (r'int philox_start_slot', 'philox'),
           (r'^__global__', 'prologue / load'), (r'// ---- D:', 'dynamics'), (r'// ---- W:', 'argmin loop glue'),
           (r'// ---- C:', 'wall test'), (r'// ---- X:', 'SAT'), (r'// ---- reward', 'reward / term / place'),
           (r'// ---- episode statistics', 'stats / info / reset'), (r'// ---- write state', 'write state'),
           (r'// ---- observations', 'obs non-ray'), (r'// rays: the warp', 'ray setup (dirs)'),
           (r'// float64 re-evaluation', 'refine f64'), (r'^        } else \{$', 'exact rays')]


def main():
    rep, cubin, kernel = sys.argv[1:4]
    src = open('/root/repo/self_play_racing_b200/csrc/rk_step.cu').read().splitlines()
    starts = []
    for pat, name in MARKERS:
        for i, ln in enumerate(src, 1):
            if re.search(pat, ln):
                starts.append((i, name))
                break
    starts.sort()
    def phase(l):
        name = 'header'
        for i, n in starts:
            if l >= i:
                name = n
        return name
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    h = rows[hdr]
    ie, ns, te = h.index('Instructions Executed'), h.index('# Samples'), h.index('Thread Instructions Executed')
    table = nl.line_table(cubin, kernel)
    agg, base, tot, tots = {}, None, 0, 0
    for r in rows[hdr + 1:]:
        if not r or not r[0].startswith('0x'):
            break
        addr = int(r[0], 16)
        base = addr if base is None else base
        f, l = table.get(addr - base, ('?', 0))
        name = phase(l) if f == 'rk_step.cu' else 'other:' + f
        a = agg.setdefault(name, [0, 0, 0])
        a[0] += int(r[ie]); a[1] += int(r[ns]); a[2] += int(r[te]); tot += int(r[ie]); tots += int(r[ns])
    print(f'total warp-instructions {tot}')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f'{a[0] / tot * 100:5.1f}% inst {a[1] / max(tots, 1) * 100:5.1f}% stall  lanes {a[2] / max(a[0], 1):4.1f}  {k}')


if __name__ == '__main__':
    main()
