for s in 1 2 3 4; do
  echo -n "in_flight=$s: "
  timeout 100 python bench.py --no-cpu-baseline --steps 400 --e2e-steps 4 --in-flight $s 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step']*1e3,1), d['roofline']['kernels_us'], d['clocks'])"
done
