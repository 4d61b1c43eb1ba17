# same-box A/B of an environment switch:  tools/ab_bench.sh VAR v1 v2 ...   (quick C2 bench per value, twice; "-" = unset)
VAR=$1; shift
for rep in 1 2; do for v in "$@"; do
  if [ "$v" = "-" ]; then unset $VAR; else export $VAR=$v; fi
  timeout 200 python bench.py --steps 1000 --no-c5 --no-matmul --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$VAR=$v', round(d['ms_per_step']*1e3,2), round(d['e2e']['ms_per_step']*1e3,1), json.dumps({k:round(v*1e3,1) for k,v in d['roofline']['kernel_ms_per_call'].items()}))"; done; done
