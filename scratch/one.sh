python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --streams 1 --scenes-per-step 4 --e2e-scenes 1 --e2e-steps 1 --no-c3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);r=d['roofline'];print(round(d['value']/1e9,3),{k:round(v/12*1e3,1) for k,v in r['kernel_ms'].items()})"
