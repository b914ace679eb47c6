#!/bin/bash
# One line per workload: device-resident value, step ms, step-kernel ms, Gymnasium-face value and ms.
#   tools/quick_bench.sh [extra bench.py flags]
for wl in multi2_selfplay_65536 single_65536; do
  python bench.py --steps 400 --warmup 20 --no-cpu-baseline --ppo-updates 0 --no-strong --workload $wl "$@" | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('$wl', round(d['value']/1e6,1), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), 'face', round(e['value']/1e6,1), round(e['ms_per_step'],4))"
done
