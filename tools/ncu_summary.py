#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep (ncu --set full) as the JSON record bench.py attaches to its roofline line.

  python tools/ncu_summary.py <report.ncu-rep> <workload-name> [<out.json>]   # merges into out.json when given
"""
import csv
import io
import json
import subprocess
import sys

FIELDS = {
    'dram_bytes_read': 'dram__bytes_read.sum', 'dram_bytes_write': 'dram__bytes_write.sum',
    'duration_us_under_ncu': 'gpu__time_duration.sum', 'warp_instructions': 'smsp__inst_executed.sum',
    'issue_active_pct': 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'fp64_pipe_pct': 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'fma_pipe_pct': 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'alu_pipe_pct': 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'warps_active_pct': 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'l1_hit_pct': 'l1tex__t_sector_hit_rate.pct', 'l2_hit_pct': 'lts__t_sector_hit_rate.pct',
    'registers': 'launch__registers_per_thread',
    'lanes_per_instruction': 'smsp__thread_inst_executed_per_inst_executed.ratio',
}
SCALE = {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1.0, 'ms': 1e3, 'us': 1.0, 'ns': 1e-3, 's': 1e6}


def main():
    rep, name = sys.argv[1:3]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, row = rows[0], rows[1], rows[2]
    col = {n: i for i, n in enumerate(hdr)}
    rec = {'kernel': row[col['Kernel Name']]}
    for key, metric in FIELDS.items():
        if metric in col:
            v = float(row[col[metric]].replace(',', ''))
            rec[key] = v * SCALE.get(units[col[metric]], 1.0) if key.startswith(('dram', 'duration')) else v
    rec['source'] = f'{rep.split("/")[-1]} (ncu --set full --clock-control none --import-source on, one launch of bench.py)'
    if len(sys.argv) > 3:
        try:
            out = json.load(open(sys.argv[3]))
        except Exception:
            out = {}
        out[name] = rec
        json.dump(out, open(sys.argv[3], 'w'), indent=1)
    print(json.dumps({name: rec}, indent=1))


if __name__ == '__main__':
    main()
