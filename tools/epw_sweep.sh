for epw in 2 4 8 16 32; do
  for wl in multi2_selfplay_65536 single_65536; do
    RK_B200_EPW=$epw python bench.py --steps 60 --warmup 10 --no-cpu-baseline --ppo-updates 0 --workload $wl 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('epw $epw', '$wl', round(d['value']/1e6,1), round(d['roofline']['kernel_ms'],4))"
  done
done
