# other BASELINE.json configurations (lift-splat stage alone, standalone shapes): device-resident numbers
for c in config4 config5; do
  for s in 1 2; do
    echo -n "$c in_flight=$s: "
    timeout 250 python bench.py --config $c --no-cpu-baseline --steps 40 --warmup 5 --e2e-steps 8 --sets 2 --in-flight $s 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), 'samples/s', round(d['ms_per_step']*1e3,1), 'us/step', r['kernels_us'], 'fwd frac %.3f step frac %.3f' % (r['frac'], r['step_frac_of_hbm_roofline']))"
  done
done
