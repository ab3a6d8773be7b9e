import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line(
        'markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session', autouse=True)
def _built_libraries():
    """The shared libraries are build artefacts (git-ignored): build them when a fresh
    checkout runs the tests before `__graft_entry__.build()`.  nvcc cross-compiles without a GPU."""
    need = [os.path.join(ROOT, 'pc_accumulation_lib_b200', 'libpcacc.so'),
            os.path.join(ROOT, 'oracle', 'liboracle_fma.so')]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()
    yield


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container (runs under gpurun)')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)


def unpack_bev(g, prefix='bev_'):
    """Inverse of tests/golden/make_golden.py:pack_bev."""
    bev = {}
    for k in g.files:
        if not k.startswith(prefix):
            continue
        name = k[len(prefix):]
        if name.endswith('_n') and (name.startswith('trajs_')
                                    or name.startswith('gt_lanes')):
            base = name[:-2]
            bev[base] = [g[f'{prefix}{base}_{i}'] for i in range(int(g[k]))]
        elif name.startswith('trajs_') or name.startswith('gt_lanes'):
            continue
        else:
            bev[name] = g[k]
    return bev


def unpack_sem_pcs(g, prefix='sem_pcs_'):
    out = []
    for i in range(int(g[f'{prefix}n_frames'])):
        out.append(np.concatenate(
            [g[f'{prefix}xyzi_{i}'], g[f'{prefix}attr_{i}'].astype(np.float64)],
            axis=1))
    return out


def assert_bev_equal(got, want, exact=True):
    """18-key BEV dict comparison.  exact: fp16 planes bit-equal; else allow
    1 fp16 ulp (SURVEY.md §8d parity checks)."""
    assert set(k for k in want) <= set(k for k in got), \
        (sorted(want), sorted(got))
    for k, w in want.items():
        g = got[k]
        if isinstance(w, list):
            assert len(g) == len(w), k
            for a, b in zip(g, w):
                a = np.asarray(a, dtype=np.float64).reshape(-1, 3)
                b = np.asarray(b, dtype=np.float64).reshape(-1, 3)
                assert a.shape == b.shape, (k, a.shape, b.shape)
                np.testing.assert_array_equal(a, b, err_msg=k)
            continue
        assert g.dtype == np.float16 and g.shape == w.shape, (k, g.dtype,
                                                              g.shape)
        if exact:
            bad = np.flatnonzero(g.view(np.uint16) != w.view(np.uint16))
            assert bad.size == 0, (k, bad[:5], g.ravel()[bad[:5]],
                                   w.ravel()[bad[:5]])
        else:
            d = np.abs(g.view(np.int16).astype(np.int32)
                       - w.view(np.int16).astype(np.int32))
            assert d.max() <= 1, (k, int(d.max()))


@pytest.fixture(scope='session')
def golden():
    return load_golden
