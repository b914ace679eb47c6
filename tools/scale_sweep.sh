# BASELINE configs[4]: 4 cars/track, 64 rays, varying segment counts S = 2 * n_ctrl * factor
echo "# workload sweep_4car_64ray (4 cars, 64 rays, obs 80-dim), E = 262,144 unless stated; columns: factor, mean S, M agent-steps/s, kernel ms"
for f in 5 10 25 30 35 85 170; do
  python bench.py --steps 12 --warmup 3 --no-cpu-baseline --ppo-updates 0 --workload sweep_4car_64ray --factor $f 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print($f, 24*$f, round(d['value']/1e6,1), round(d['roofline']['kernel_ms'],3))"
done
echo "# E = 1,048,576 (the config's full size), factor 30"
python bench.py --steps 6 --warmup 2 --no-cpu-baseline --ppo-updates 0 --workload sweep_4car_64ray --envs 1048576 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(30, 720, round(d['value']/1e6,1), round(d['roofline']['kernel_ms'],3))"
