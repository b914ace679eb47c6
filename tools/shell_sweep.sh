for sh in "none" "10" "20" "30" "14,36" "10,25" "8,20,40"; do
  for wl in multi2_selfplay_65536 single_65536; do
    if [ "$sh" = "none" ]; then export RK_B200_SHELLS="1e9"; else export RK_B200_SHELLS="$sh"; fi
    python bench.py --steps 60 --warmup 10 --no-cpu-baseline --workload $wl 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$sh', '$wl', round(d['value']/1e6,1), round(d['roofline']['kernel_ms'],4))"
  done
done
