timeout 200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for fw in 1 2; do for bps in 4 8; do
  echo -n "fill=$fw blocks_per_sm=$bps: "
  LSS_FILL_WARPS=$fw LSS_POOL_BLOCKS_PER_SM=$bps timeout 100 python bench.py --no-cpu-baseline --steps 100 --e2e-steps 4 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,1), d['roofline']['kernels_us'])"
done; done
