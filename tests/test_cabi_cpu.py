"""CPU-side checks of the boundary: libpcacc.so loads, exports every symbol
include/pcacc.h declares, refuses to run without a GPU; the host logic of the
Python mirror (trajectories, heading, eviction bookkeeping) matches the golden
outputs of the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests.conftest import ROOT, load_golden, unpack_bev
from tests.golden import cases


def _header_functions():
    src = open(os.path.join(ROOT, 'include', 'pcacc.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(pcacc_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from pc_accumulation_lib_b200 import _lib
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/pcacc.h but not exported'
        assert n in _lib.SIGNATURES, f'{n} has no ctypes signature'
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.pcacc_abi_version() == 1
    assert lib.pcacc_strerror(-3) == b'ring or frame table capacity exceeded'


def test_bev_params_struct_layout_matches_header():
    from pc_accumulation_lib_b200._lib import BevParams
    # 3 int64 + 20 doubles + 6 int32 = 24 + 160 + 24 (static_assert'ed in raster.cu too)
    assert C.sizeof(BevParams) == 208
    assert BevParams.origin.offset == 24 and BevParams.R.offset == 48
    assert BevParams.road_cls.offset == 184 and BevParams.elevation_max.offset == 204


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from pc_accumulation_lib_b200 import _lib
    from pc_accumulation_lib_b200.device import DeviceCloud
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.pcacc_create(0, 1000, 8, C.byref(h)) == _lib.ERR_CUDA
    assert b'no CPU fallback' in lib.pcacc_last_error(None)
    with pytest.raises(_lib.PcaccError):
        DeviceCloud(1000, 8)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may import,
    link or execute it (bench.py's cpu_baseline legs are the only other user)."""
    pkg = os.path.join(ROOT, 'pc_accumulation_lib_b200')
    bad = re.compile(r'^\s*(from|import)\s+oracle\b|oracle[/.]oracle|liboracle|oracle/', re.M)
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', 'Makefile')):
                txt = open(os.path.join(dp, f)).read()
                assert not bad.search(txt), os.path.join(dp, f)


def _gen():
    from pc_accumulation_lib_b200.bev_generator import SemBEVGenerator
    return SemBEVGenerator


@pytest.mark.parametrize('name,kw', [
    ('bev_direct.npz', {}),
    ('bev_direct_p128.npz', dict(n=20000, seed=78, P=128, view=51.2)),
])
def test_host_trajectory_path_matches_reference(name, kw):
    g = load_golden(name)
    pcs, trajs, aug, gen = cases.bev_direct_inputs(**kw)
    bg = _gen()(gen['sem_idxs'], gen['view_size'], gen['pixel_size'], 0., 0., False,
                gen['int_scaler'], gen['int_sep_scaler'], gen['int_mid_threshold'],
                gen['height_filter'], gen['rgb_fill'])
    for prefix, a in (('bev_', aug), ('bevhead_', None)):
        want = unpack_bev(g, prefix)
        if a is None:
            rot = bg.heading_angle(trajs['ego_traj_present'])
            dx = dy = 0.
            view = gen['view_size']
        else:
            rot, dx, dy = a['rot_ang'], a['trans_dx'], a['trans_dy']
            view = a['zoom_scalar'] * gen['view_size']
        for w in ('present', 'future', 'full'):
            tr = [trajs[f'ego_traj_{w}']] + list(trajs[f'other_trajs_{w}'])
            got = bg.preprocess_trajs([t.copy() for t in tr], rot, dx, dy, view)
            assert len(got) == len(want[f'trajs_{w}'])
            for x, y in zip(got, want[f'trajs_{w}']):
                np.testing.assert_array_equal(np.asarray(x).reshape(-1, 3),
                                              np.asarray(y).reshape(-1, 3))
        # the batched helper (all windows, one call) gives the same rows
        groups = [[trajs[f'ego_traj_{w}']] + list(trajs[f'other_trajs_{w}'])
                  for w in ('present', 'future', 'full')]
        aug_v = dict(rot_ang=rot, trans_dx=dx, trans_dy=dy, zoom_scalar=view / gen['view_size'])
        assert aug_v['zoom_scalar'] * gen['view_size'] == view
        got_b = bg.preprocess_trajs_batch(groups, [aug_v, aug_v])
        for v in range(2):
            for gi, w in enumerate(('present', 'future', 'full')):
                assert len(got_b[v][gi]) == len(want[f'trajs_{w}'])
                for x, y in zip(got_b[v][gi], want[f'trajs_{w}']):
                    np.testing.assert_array_equal(x, np.asarray(y).reshape(-1, 3))


def test_rand_aug_is_injectable_and_ordered_like_the_reference():
    bg = _gen()({'road': 0, 'car': 13, 'truck': 14, 'bus': 15, 'motorcycle': 17}, 80, 64,
                max_trans_radius=5., zoom_thresh=0.1)
    assert bg.do_aug
    bg.rng = np.random.RandomState(7)
    a = bg.rand_aug_params()
    r = np.random.RandomState(7)
    rot = 2 * np.pi * r.random_sample()
    tr = 5. * r.random_sample()
    ta = 2 * np.pi * r.random_sample()
    z = min(max(r.normal(0, 0.1), -0.1), 0.1)
    assert a['rot_ang'] == rot and a['trans_dx'] == tr * np.cos(ta)
    assert a['trans_dy'] == tr * np.sin(ta) and a['zoom_scalar'] == 1 + z


def test_async_writer_files_are_the_reference_format(tmp_path):
    """sem_pc_accum.py:280-308: a .gz holding pickle.dumps(bev); the threaded writer must
    produce what the synchronous one does (read here the way read_compressed_pickle does)."""
    import gzip
    import pickle
    from pc_accumulation_lib_b200.sem_pc_accum import SemanticPointCloudAccumulator as Acc
    rng = np.random.RandomState(5)
    bevs = []
    for k in range(12):
        bev = {f'{n}_{w}': rng.rand(64, 64).astype(np.float16)
               for n in ('road', 'intensity', 'dynamic', 'elevation') for w in ('present', 'future', 'full')}
        bev.update({f'rgb_{w}': rng.rand(3, 64, 64).astype(np.float16) for w in ('present', 'future', 'full')})
        bev.update({f'trajs_{w}': [rng.rand(5, 3)] for w in ('present', 'future', 'full')})
        bevs.append(bev)
    with Acc.async_writer(n_threads=4, max_pending=3) as wr:
        for k, bev in enumerate(bevs):
            wr.submit(bev, f'bev_{k:03d}.pkl', str(tmp_path))
    assert wr.n_written == len(bevs) and wr.bytes_written > 0
    for k, bev in enumerate(bevs):
        Acc.write_compressed_pickle(bev, f'sync_{k:03d}.pkl', str(tmp_path))
        with gzip.open(tmp_path / f'bev_{k:03d}.pkl.gz', 'rb') as f:
            a = f.read()
        with gzip.open(tmp_path / f'sync_{k:03d}.pkl.gz', 'rb') as f:
            b = f.read()
        assert a == b == pickle.dumps(bev)
        got = Acc.read_compressed_pickle(str(tmp_path / f'bev_{k:03d}.pkl.gz'))
        assert sorted(got) == sorted(bev)
        np.testing.assert_array_equal(got['rgb_full'], bev['rgb_full'])
    # an unwritable directory surfaces as IOError at close (the reference prints and goes on)
    wr = Acc.async_writer(2)
    wr.submit(bevs[0], 'x', str(tmp_path / 'missing_dir'))
    with pytest.raises(IOError):
        wr.close()


def test_tracker_helpers_keep_reference_semantics():
    """Host bookkeeping of the nuScenes accumulator that was restructured for speed: the batched
    box transform keeps the single-point product's bits, and the trajectory extraction equals
    the reference's find_nearest_* / parse_* composition (nuscenes_oracle_sem_pc_accum.py:
    272-414) on random sighting patterns."""
    from oracle import oracle as orc
    from pc_accumulation_lib_b200 import NuScenesOracleSemanticPointCloudAccumulator as Acc
    rng = np.random.default_rng(3)
    acc = Acc.__new__(Acc)
    for _ in range(20):
        T = np.linalg.inv(rng.normal(size=(4, 4)) * 10.)
        T[3] = [0., 0., 0., 1.]
        acc.T_global_world = T
        c = rng.normal(size=(30, 3)) * 500.
        want = np.stack([orc.homo_transform(T, c[i:i + 1])[0] for i in range(c.shape[0])])
        np.testing.assert_array_equal(acc._boxes_to_world(list(c)), want)

    def reference_trajs(instances, dyn, ts_start=0, ts_end=None):
        out = []
        for token, seen in instances.items():
            if token not in dyn:
                continue
            poses, tss = zip(*seen)
            try:
                i0 = Acc.find_nearest_ge_idx(tss, ts_start)
                i1 = None if ts_end is None else Acc.find_nearest_le_idx(tss, ts_end) + 1
            except ValueError:
                continue
            poses, tss = poses[i0:i1], tss[i0:i1]
            for run in Acc.parse_seq_into_coherent_seqs(tss):
                if len(run) >= 2:
                    out.append([poses[k].tolist() for k in run])
        return out

    for trial in range(200):
        instances, dyn = {}, []
        for t in range(int(rng.integers(1, 6))):
            ts = np.flatnonzero(rng.random(30) < rng.uniform(0.3, 0.95))
            if ts.size == 0:
                continue
            instances[f'tok{t}'] = [(rng.normal(size=3), int(k)) for k in ts]
            if rng.random() < 0.7:
                dyn.append(f'tok{t}')
        acc.instances, acc.dyn_instances, acc._dyn_set, acc.poses = instances, dyn, set(dyn), [[0., 0., 0.]]
        acc._pose_lists = {}
        split = int(rng.integers(0, 30))
        for kw in (dict(ts_end=split), dict(ts_start=split), dict()):
            try:
                want = reference_trajs(instances, dyn, **kw)
            except IndexError:
                with pytest.raises(IndexError):
                    acc.get_dyn_obj_trajs(**kw)
                continue
            assert acc.get_dyn_obj_trajs(**kw) == want


def test_dataset_mirror_transforms_match_oracle():
    from oracle import oracle as orc
    from pc_accumulation_lib_b200.datasets import nuscenes_utils as nu
    rng = np.random.default_rng(9)
    T = np.linalg.inv(rng.normal(size=(4, 4)))
    T[3] = [0., 0., 0., 1.]
    for n in (1, 2, 257):
        pts = rng.normal(size=(n, 3)) * 100.
        np.testing.assert_array_equal(nu.homo_transform(T, pts), orc.homo_transform(T, pts))
        np.testing.assert_array_equal(nu.apply_tf(T, pts), orc.apply_tf(T, pts))
        q = np.concatenate([pts, rng.normal(size=(n, 2))], axis=1)
        want = q.copy()
        want[:, :3] = orc.apply_tf(T, pts)
        assert nu.apply_tf(T, q, in_place=True) is None
        np.testing.assert_array_equal(q, want)
    with pytest.raises(AssertionError):
        nu.homo_transform(np.eye(3), np.zeros((2, 3)))
    with pytest.raises(AssertionError):
        nu.homo_transform(np.eye(4), np.zeros((2, 4)))


def test_update_poses_keeps_the_reference_bits():
    """sem_pc_accum.py:156-165 evaluates (4,4) @ (4,1) once per stored pose; the mirror's single stacked
    call must give the same float64 bits (the poses decide eviction and the BEV origin)."""
    from pc_accumulation_lib_b200.sem_pc_accum import SemanticPointCloudAccumulator as Acc
    rng = np.random.default_rng(11)
    acc = Acc.__new__(Acc)
    for _ in range(50):
        a, b = rng.normal(0, 0.5), rng.normal(0, 0.02)
        T = np.eye(4)
        T[:3, :3] = (np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
                     @ np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]]))
        T[:3, 3] = rng.normal(0, 3, 3)
        poses = [list(rng.normal(0, 50, 3)) for _ in range(int(rng.integers(1, 40)))]
        want = [list(np.matmul(T, np.array([p + [1]]).T)[:, 0][:-1]) for p in poses]
        acc.poses = [list(p) for p in poses]
        acc.update_poses(T)
        assert acc.poses == want
        assert all(type(p) is list and len(p) == 3 for p in acc.poses)
    acc.poses = []
    acc.update_poses(np.eye(4))
    assert acc.poses == []


def test_binding_constants_match_the_header():
    """Every integer constant the ctypes binding mirrors (status codes, flag bits, dtype codes, staging modes,
    options, kernel classes) has the value include/pcacc.h defines under the PCACC_ name."""
    import re
    from pc_accumulation_lib_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'pcacc.h')).read()
    defs = {}
    for name, val in re.findall(r'^#define\s+PCACC_(\w+)\s+\(?(-?(?:0x[0-9a-fA-F]+|\d+))u?\)?\s*(?:/\*.*)?$', hdr, re.M):
        defs[name] = int(val, 0)
    for name, val in re.findall(r'^\s*PCACC_(\w+)\s*=\s*(-?\d+)\s*,?\s*(?:/\*.*)?$', hdr, re.M):   # enum pcacc_status
        defs[name] = int(val)
    checked = 0
    for attr in dir(_lib):
        if attr.isupper() and isinstance(getattr(_lib, attr), int) and attr in defs:
            assert getattr(_lib, attr) == defs[attr], attr
            checked += 1
    assert checked >= 25, (checked, sorted(defs))
    assert len(_lib.KERNEL_CLASSES) == defs['N_KERNELS']
    for k, want in (('INTEGRATE', 'integrate'), ('SCAN', 'scan'), ('REDUCE_BIG', 'bev_reduce_big'), ('CLASSIFY', 'bev_classify')):
        assert _lib.KERNEL_CLASSES[defs['K_' + k]] == want
