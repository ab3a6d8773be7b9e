#!/bin/bash
# A/B of prebuilt libpcacc variants (scratch_so/*.so): long-horizon window stage times and the
# bench workload's per-kernel "alone" times.  Scratch tool for kernel work.
for so in "$@"; do
  cp "$so" pc_accumulation_lib_b200/libpcacc.so
  echo "== $so"
  python tools/bench_c3.py 2>&1 | tail -2
  python bench.py --steps 2 --warmup 3 --e2e-scenes 1 --e2e-steps 1 --no-c3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('bench ms/step', round(d['ms_per_step'],3), d['roofline']['kernel_us_alone'], 'parity', d['parity'].get('status', d['parity']) if isinstance(d.get('parity'),dict) else d.get('parity'))"
done
