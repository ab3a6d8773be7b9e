"""TEST INFRASTRUCTURE — pins the golden fixtures and the oracle against the UNMODIFIED reference.

Runs in the BUILD container only (needs /root/reference):

    python oracle/validate_against_reference.py

1. Re-generates every tests/golden/*.npz in memory by running the literal reference
   (tests/golden/make_golden.py, reference imported through oracle/ref_loader.py) and checks
   that each array is bit-identical to the committed fixture — the fixtures are outputs of the
   reference itself, reproducible from the committed generator.
2. Runs the oracle's own CPU tests (tests/test_oracle_cpu.py), which compare oracle/oracle.py
   with those fixtures `==`.
Exit status 0 = both hold.  Nothing on the product path imports this module.
"""
from __future__ import annotations

import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader                    # noqa: E402


def main() -> int:
    if not ref_loader.available():
        print('validate_against_reference: /root/reference is not present here (build container only)')
        return 2
    from tests.golden import make_golden as mg
    fresh = {}
    mg.save = lambda name, out: fresh.__setitem__(name, out)      # capture instead of writing
    mg.main()
    bad = 0
    for name, out in sorted(fresh.items()):
        path = os.path.join(ROOT, 'tests', 'golden', name)
        with np.load(path, allow_pickle=False) as g:
            keys_ok = set(g.files) == set(out)
            diff = [k for k in out if k in g.files and not (
                np.asarray(out[k]).dtype == g[k].dtype and np.asarray(out[k]).shape == g[k].shape
                and np.array_equal(np.asarray(out[k]), g[k], equal_nan=np.asarray(out[k]).dtype.kind == 'f'))]
        status = 'ok' if keys_ok and not diff else f'MISMATCH {diff[:4]}'
        print(f'{name}: {len(out)} arrays regenerated from the reference: {status}')
        bad += 0 if status == 'ok' else 1
    rc = subprocess.call([sys.executable, '-m', 'pytest', '-q', '-x',
                          os.path.join(ROOT, 'tests', 'test_oracle_cpu.py')], cwd=ROOT)
    print('fixtures reproduce:', bad == 0, '| oracle == fixtures:', rc == 0)
    return 0 if (bad == 0 and rc == 0) else 1


if __name__ == '__main__':
    sys.exit(main())
