for s in 2 4 6 8; do python bench.py --steps 10 --warmup 3 --streams $s --e2e-scenes 1 --e2e-steps 1 --no-c3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('streams',$s,round(d['value']/1e9,3),d['ms_per_step'])"; done
