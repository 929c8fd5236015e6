for cfg in "4 4" "6 6" "8 8" "8 4" "12 6"; do
  set -- $cfg
  echo -n "sets=$1 in_flight=$2: "
  timeout 100 python bench.py --no-cpu-baseline --steps 480 --e2e-steps 4 --sets $1 --in-flight $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step']*1e3,1))"
done
