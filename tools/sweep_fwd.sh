timeout 200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for fc in 74 148 222 296 444; do
  echo -n "fill_ctas=$fc: "
  LSS_FILL_CTAS=$fc timeout 100 python bench.py --no-cpu-baseline --steps 200 --e2e-steps 4 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,1), d['roofline']['kernels_us'])"
done
