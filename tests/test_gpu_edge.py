"""Edge cases and randomised parity of the CUDA path against the oracle: empty and ragged
frames, ring wrap-around with eviction, capacity errors, > 32 variants in one call, the
1024 x 1024 grid with dynamic points, the probability-map (argmax) integrate, the
north-star elevation-max variant and non-default rgb_fill / class ids.
"""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import oracle as orc                                  # noqa: E402
from pc_accumulation_lib_b200 import synth                        # noqa: E402
from tests.test_gpu_core import (bev_params_from, compare_planes, gen_params,  # noqa: E402
                                 window_cells)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    from pc_accumulation_lib_b200 import device
    return device


def rand_cloud(rng, m, spread=20., dyn_frac=0.05, classes=(0, 0, 0, 1, 2, 8, 13, 14, 15, 17, 5)):
    pc = np.zeros((m, 10))
    pc[:, 0:2] = rng.normal(0., spread, (m, 2))
    pc[:, 2] = rng.normal(0., 2., m)
    pc[:, 3] = rng.uniform(0., 1., m).astype(np.float32)         # float32-representable
    pc[:, 4:7] = rng.integers(0, 256, (m, 3))
    pc[:, 7] = rng.choice(list(classes), m)
    pc[:, 8] = rng.integers(-1, 6, m)
    pc[:, 9] = (rng.random(m) < dyn_frac).astype(float)
    return pc


def oracle_bev(frames, split, gp, origin, rot, dx=0., dy=0., zoom=1.):
    def cat(fs):
        return np.concatenate(fs) if fs else np.zeros((0, 10))
    pcs = {}
    for w, fs in (('present', frames[:split]), ('future', frames[split:]), ('full', frames)):
        pc = cat(fs).copy()
        pc[:, :3] = pc[:, :3] - origin
        pcs[f'pc_{w}'] = pc
    z = np.zeros((2, 3))
    trajs = {f'ego_traj_{w}': z for w in ('present', 'future', 'full')}
    trajs.update({f'other_trajs_{w}': [] for w in ('present', 'future', 'full')})
    return orc.generate(pcs, trajs, gp, return_f64=True, rot_ang=rot, trans_dx=dx, trans_dy=dy,
                        zoom_scalar=zoom, do_warping=True)


def check_against_oracle(dev, cloud, fids, frames, split, gp, origin, rot, dx=0., dy=0., zoom=1.,
                         exact=(0, 2, 3, 4, 5, 6)):
    P = gp['pixel_size']
    bp = bev_params_from(dev, gp, fids[0], fids[split], fids[-1] + 1, origin, rot, dx, dy, zoom)
    o16, o64, cells = cloud.rasterise([bp], P, want_f64=True, want_cells=True)
    assert cloud.sync() & ~2 == 0
    ref = oracle_bev(frames, split, gp, origin, rot, dx, dy, zoom)
    dbg = ref['_debug']
    compare_planes(o16[0].cpu().numpy(), o64[0].cpu().numpy(),
                   {w: dbg[f'planes_f64_{w}'] for w in ('present', 'future', 'full')}, exact=exact)
    np.testing.assert_array_equal(window_cells(cloud, cells, fids[:split], P), dbg['cells_present'])
    np.testing.assert_array_equal(window_cells(cloud, cells, fids[split:], P), dbg['cells_future'])
    return o16


# ---------------------------------------------------------------------------
def test_empty_and_ragged_frames(dev):
    """Frame sizes 0, 1, 3, 1023, 1024, 1025, 130 and a frame whose points are all filtered."""
    rng = np.random.default_rng(1)
    gp = gen_params(synth.kitti_bev_params(pixel_size=32, view_size=60.), height_filter=1.5)
    sizes = [0, 1, 3, 1023, 1024, 1025, 130, 0, 7]
    frames = [rand_cloud(rng, m) for m in sizes]
    cloud = dev.DeviceCloud(capacity_pts=8000, max_frames=32)
    fids = [cloud.integrate_cloud(f) for f in frames]
    assert cloud.sync() == 0
    for fid, f in zip(fids, frames):
        assert cloud.frame_count(fid) == f.shape[0]
        np.testing.assert_array_equal(cloud.export_frame(fid), f)
    check_against_oracle(dev, cloud, fids, frames, 4, gp, np.array([1., -2., 0.3]), 0.4)
    # split at the very beginning / end: an entirely empty present or future window
    for split in (1, len(frames) - 1):
        check_against_oracle(dev, cloud, fids, frames, split, gp, np.zeros(3), -1.1)
    cloud.close()

    # KITTI frustum / gt paths with n = 0 and with every point filtered
    calib = synth.kitti_calib()
    cloud = dev.DeviceCloud(capacity_pts=300_000, max_frames=16)
    seed = synth.seed_for(1, 0)
    rgb, cls = synth.kitti_rgb(seed), synth.kitti_class_map(seed)
    f0 = cloud.integrate_frustum(np.zeros((0, 4), np.float32), calib['p_velo_frame'], rgb, cls,
                                 synth.KITTI_FILTERS)
    pc = synth.kitti_lidar(seed, 8, 300)
    f1 = cloud.integrate_frustum(pc, calib['p_velo_frame'], rgb, np.full_like(cls, 10),
                                 synth.KITTI_FILTERS)          # class 10 is filtered: nothing kept
    f2 = cloud.integrate_gt(pc, np.full((pc.shape[0], 1), 255, np.int16), synth.KITTI_FILTERS)
    f3 = cloud.integrate_gt(np.zeros((0, 4), np.float32), np.zeros((0, 1), np.int16),
                            synth.KITTI_FILTERS)
    f4 = cloud.integrate_gt(pc, synth.kitti_sem_gt(seed, pc.shape[0]), synth.KITTI_FILTERS)
    assert cloud.sync() == 0
    assert [cloud.frame_count(f) for f in (f0, f1, f2, f3, f4)] == [0, 0, 0, 0, pc.shape[0]]
    cloud.close()


@pytest.mark.parametrize('seed', range(6))
def test_randomised_parity(dev, seed):
    """Random grid size, view, rotation, translation, zoom, height filter, class ids, fill
    colour and elevation mode; clouds with crowded cells and dynamic points."""
    rng = np.random.default_rng(100 + seed)
    P = int(rng.choice([32, 64, 128]))
    view = float(rng.choice([20., 51.2, 80.]))
    sem_idxs = {'road': int(rng.integers(0, 4)), 'car': 13, 'truck': int(rng.integers(5, 9)),
                'bus': 15, 'motorcycle': 17}
    gp = dict(sem_idxs=sem_idxs, view_size=view, pixel_size=P, int_scaler=float(rng.choice([1., 20.])),
              int_sep_scaler=float(rng.choice([20., 30.])), int_mid_threshold=float(rng.choice([.12, .5])),
              height_filter=(None if seed % 2 else float(rng.uniform(0.5, 3.))),
              rgb_fill=int(rng.choice([0, 128])),
              elevation_mode=('max' if seed % 3 == 0 else 'min'))
    n_frames = int(rng.integers(2, 7))
    frames = []
    for _ in range(n_frames):
        m = int(rng.integers(1, 9000))
        pc = rand_cloud(rng, m, spread=view / 3, dyn_frac=0.1, classes=range(0, 19))
        k = m // 4                                       # a few very crowded cells
        pc[:k, 0:2] = rng.normal(0., view / 200, (k, 2)) + rng.uniform(-view / 4, view / 4, 2)
        frames.append(pc)
    cloud = dev.DeviceCloud(capacity_pts=sum(f.shape[0] for f in frames) + 64, max_frames=16)
    fids = [cloud.integrate_cloud(f) for f in frames]
    assert cloud.sync() == 0
    split = int(rng.integers(1, n_frames))
    origin = rng.normal(0., 2., 3)
    check_against_oracle(dev, cloud, fids, frames, split, gp, origin, float(rng.uniform(-3.2, 3.2)),
                         float(rng.uniform(-3, 3)), float(rng.uniform(-3, 3)),
                         float(rng.uniform(0.9, 1.1)))
    cloud.close()


def test_more_than_32_variants_and_frame_subranges(dev):
    """40 variants in one call (two launch groups); variants use different frame ranges."""
    rng = np.random.default_rng(3)
    gp = gen_params(synth.nusc_bev_params(pixel_size=32))
    frames = [rand_cloud(rng, 2500, spread=12.) for _ in range(5)]
    cloud = dev.DeviceCloud(capacity_pts=20000, max_frames=16)
    fids = [cloud.integrate_cloud(f) for f in frames]
    cloud.sync()
    variants = []
    for k in range(40):
        b, e = (0, 5) if k % 3 else (1, 4)
        s = b + 1 + k % (e - b - 1)
        variants.append((b, s, e, rng.normal(0, 1, 3), float(rng.uniform(-3, 3)), float(rng.uniform(-2, 2)),
                         float(rng.uniform(-2, 2)), float(rng.uniform(0.9, 1.1))))
    bps = [bev_params_from(dev, gp, fids[b], fids[s], fids[e - 1] + 1, o, r, dx, dy, z)
           for b, s, e, o, r, dx, dy, z in variants]
    o16, o64, _ = cloud.rasterise(bps, 32, want_f64=True)
    cloud.sync()
    o16, o64 = o16.cpu().numpy(), o64.cpu().numpy()
    for k, (b, s, e, o, r, dx, dy, z) in enumerate(variants):
        ref = oracle_bev(frames[b:e], s - b, gp, o, r, dx, dy, z)['_debug']
        compare_planes(o16[k], o64[k], {w: ref[f'planes_f64_{w}'] for w in ('present', 'future', 'full')})
    cloud.close()


def test_ring_wraparound_with_eviction(dev):
    """A ring much smaller than the stream: frames are evicted by the KITTI horizon rule,
    the tail wraps to offset 0, records and BEVs stay equal to the oracle's."""
    frames = []
    for f in range(30):
        seed = synth.seed_for(3, 700 + f)
        pc = synth.kitti_lidar(seed, 8, 250)                      # 2000 points
        frames.append(dict(pc=pc, T=synth.kitti_step_transform(seed),
                           sem_gt=synth.kitti_sem_gt(seed, pc.shape[0])))
    P = 64
    acc = orc.KittiOracle(12.0, synth.kitti_calib()['p_velo_frame'], synth.KITTI_FILTERS,
                          gen_params(synth.kitti_bev_params(pixel_size=P)), use_gt_sem=True)
    cloud = dev.DeviceCloud(capacity_pts=15_000, max_frames=16)     # ~7 frames fit
    offsets = []
    for fr in frames:
        if cloud.live_frames()[1] > 0:
            cloud.rebase(fr['T'], eager=False)
        # the reference evicts after appending; the ring needs the room before: evict what the
        # oracle is about to evict first (same live set at every rasterise / export)
        ev = acc.integrate(fr['pc'], None, None, fr['T'], fr['sem_gt'])
        fid = None
        n_live = cloud.live_frames()[1]
        if ev and ev <= n_live:
            cloud.evict(int(ev))
            fid = cloud.integrate_gt(fr['pc'], fr['sem_gt'], synth.KITTI_FILTERS)
        else:
            fid = cloud.integrate_gt(fr['pc'], fr['sem_gt'], synth.KITTI_FILTERS)
            cloud.evict(int(ev))
        cloud.sync()
        offsets.append(cloud.frame_offset(fid))
    assert min(offsets[1:]) == 0 or any(b < a for a, b in zip(offsets, offsets[1:])), \
        'the ring never wrapped: the test would not exercise wrap-around'
    first, n_live = cloud.live_frames()
    assert n_live == len(acc.sem_pcs)
    for k in range(n_live):
        np.testing.assert_array_equal(cloud.export_frame(first + k), acc.sem_pcs[k])
    p_idx = n_live // 2
    origin = np.array(acc.poses[p_idx])
    rot = orc.heading_rot_ang(np.array(acc.poses[:p_idx]) - origin)
    bp = bev_params_from(dev, acc.gen_params, first, first + p_idx, first + n_live, origin, rot)
    o16, o64, cells = cloud.rasterise([bp], P, want_f64=True, want_cells=True)
    cloud.sync()
    dbg = acc.generate_bev(p_idx, return_f64=True)['_debug']
    compare_planes(o16[0].cpu().numpy(), o64[0].cpu().numpy(),
                   {w: dbg[f'planes_f64_{w}'] for w in ('present', 'future', 'full')},
                   exact=(0, 2, 3, 4, 5))
    fids = list(range(first, first + n_live))
    np.testing.assert_array_equal(window_cells(cloud, cells, fids[:p_idx], P), dbg['cells_present'])
    cloud.close()


def test_capacity_errors(dev):
    from pc_accumulation_lib_b200._lib import ERR_CAPACITY, PcaccError
    rng = np.random.default_rng(5)
    cloud = dev.DeviceCloud(capacity_pts=1000, max_frames=8)
    with pytest.raises(PcaccError) as e:
        cloud.integrate_cloud(rand_cloud(rng, 1200))
    assert e.value.status == ERR_CAPACITY
    cloud.integrate_cloud(rand_cloud(rng, 600))
    with pytest.raises(PcaccError) as e:                      # ring full: would overwrite a live frame
        cloud.integrate_cloud(rand_cloud(rng, 600))
    assert e.value.status == ERR_CAPACITY
    cloud.close()
    cloud = dev.DeviceCloud(capacity_pts=1000, max_frames=4)
    for _ in range(3):
        cloud.integrate_cloud(rand_cloud(rng, 10))
    with pytest.raises(PcaccError) as e:                      # frame table full
        cloud.integrate_cloud(rand_cloud(rng, 10))
    assert e.value.status == ERR_CAPACITY
    cloud.close()


def test_highres_1024_with_dynamic_points(dev):
    """BASELINE config 5 shape at a size the oracle finishes in seconds: P = 1024, V = 80,
    5 % of the vehicle-class points flagged dynamic, elevation in reference mode (min) and
    in the north-star mode (max)."""
    rng = np.random.default_rng(9)
    P = 1024
    frames = []
    for f in range(4):
        pc = rand_cloud(rng, 60_000, spread=25., dyn_frac=0.0, classes=range(0, 19))
        veh = np.isin(pc[:, 7], [13, 14, 15, 17])
        pc[veh & (rng.random(pc.shape[0]) < 0.05), 9] = 1
        frames.append(pc)
    cloud = dev.DeviceCloud(capacity_pts=250_000, max_frames=8)
    fids = [cloud.integrate_cloud(f) for f in frames]
    cloud.sync()
    for mode in ('min', 'max'):
        gp = gen_params(synth.kitti_bev_params(pixel_size=P), elevation_mode=mode)
        check_against_oracle(dev, cloud, fids, frames, 2, gp, np.array([0.5, 0.25, 0.1]), 0.77)
    cloud.close()


def test_probability_map_integrate(dev):
    """North-star variant of the gather: a (H,W,19) float32 probability map; the stored class
    is the first argmax, exactly as if the class-index map had been handed over."""
    seed = synth.seed_for(1, 3)
    pc = synth.kitti_lidar(seed, 16, 900)
    prob = synth.kitti_prob_map(seed)
    cls = np.argmax(prob, axis=2).astype(np.int64)
    rgb = synth.kitti_rgb(seed)
    P = synth.kitti_calib()['p_velo_frame']
    cloud = dev.DeviceCloud(capacity_pts=60_000, max_frames=8)
    fa = cloud.integrate_frustum(pc, P, rgb, prob, synth.KITTI_FILTERS)
    fb = cloud.integrate_frustum(pc, P, rgb, cls, synth.KITTI_FILTERS)
    fc = cloud.integrate_frustum(pc, P, rgb, cls.astype(np.uint8), synth.KITTI_FILTERS, max_depth=30.)
    assert cloud.sync() == 0
    a, b, c = cloud.export_frame(fa), cloud.export_frame(fb), cloud.export_frame(fc)
    want = orc.kitti_obs2sem(pc, rgb, cls, P, synth.KITTI_FILTERS)
    np.testing.assert_array_equal(a, want)
    np.testing.assert_array_equal(b, want)
    assert 0 < c.shape[0] < b.shape[0]            # max_depth drops the far points
    # the same probabilities padded to 20 channels (zero channel: 16-byte rows, vector loads),
    # with ties (first maximum wins, np.argmax), and a 7-channel map (generic loop)
    pad = np.concatenate([prob, np.zeros(prob.shape[:2] + (1,), dtype=np.float32)], axis=2)
    pad[::3, ::5, :] = 0.25                       # all-equal rows: argmax = 0
    pad[1::3, 2::5, 7] = pad[1::3, 2::5, 3] = 2.0  # two equal maxima: the first one (3)
    cls_pad = np.argmax(pad, axis=2).astype(np.int64)
    small = np.ascontiguousarray(prob[:, :, :7])
    cloud.reset()
    fd = cloud.integrate_frustum(pc, P, rgb, np.ascontiguousarray(pad), synth.KITTI_FILTERS)
    fe = cloud.integrate_frustum(pc, P, rgb, small, synth.KITTI_FILTERS)
    assert cloud.sync() == 0
    np.testing.assert_array_equal(cloud.export_frame(fd), orc.kitti_obs2sem(pc, rgb, cls_pad, P, synth.KITTI_FILTERS))
    np.testing.assert_array_equal(cloud.export_frame(fe),
                                  orc.kitti_obs2sem(pc, rgb, np.argmax(small, axis=2).astype(np.int64), P,
                                                    synth.KITTI_FILTERS))
    cloud.close()


def test_cell_population_sweep(dev):
    """Every per-cell population from 1 to 70 points, present / future split at every
    position, colours drawn from a handful of values (ties, even / odd counts): the three
    reduction paths (rank intervals in pass A up to 15 points, rank intervals over shuffles
    up to 32, histograms above) against the oracle, planes compared exactly."""
    from tests.test_gpu_core import bev_params_from, compare_planes, gen_params
    P, view = 32, 32.0                       # 1 m cells
    rng = np.random.default_rng(4242)
    palette = np.array([0, 3, 3, 128, 200, 255, 255, 17])
    pres, fut = [], []
    cell = 0
    for n in range(1, 71):
        for split in sorted({0, 1, n // 2, n - 1, n}):
            cx, cy = cell % (P - 2) + 1, cell // (P - 2) + 1       # keep off the border
            cell += 1
            pts = np.zeros((n, 10))
            pts[:, 0] = (cx + rng.uniform(0.1, 0.9, n)) - P / 2
            pts[:, 1] = (cy + rng.uniform(0.1, 0.9, n)) - P / 2
            pts[:, 2] = rng.normal(0., 1., n)
            pts[:, 3] = rng.integers(0, 256, n) / 255.
            pts[:, 4:7] = palette[rng.integers(0, len(palette), (n, 3))]
            pts[:, 7] = rng.choice([synth.SEM_IDXS['road'], synth.SEM_IDXS['car'], 8, 2], n)
            pts[:, 9] = (rng.random(n) < 0.1).astype(float)
            pres.append(pts[:split])
            fut.append(pts[split:])
    assert cell <= (P - 2) * (P - 2)
    pc_p, pc_f = np.concatenate(pres), np.concatenate(fut)
    pcs = dict(pc_present=pc_p, pc_future=pc_f, pc_full=np.concatenate([pc_p, pc_f]))
    ego = [np.array([[0., -2., 0.], [0., -1., 0.], [0., 0., 0.]]), np.array([[0., 0., 0.], [0., 1., 0.]])]
    trajs = dict(ego_traj_present=ego[0], ego_traj_future=ego[1], ego_traj_full=np.concatenate(ego),
                 other_trajs_present=[], other_trajs_future=[], other_trajs_full=[])
    gp = gen_params(synth.kitti_bev_params(pixel_size=P, view_size=view))
    cloud = dev.DeviceCloud(capacity_pts=pcs['pc_full'].shape[0] + 64, max_frames=8)
    f0 = cloud.integrate_cloud(pc_p)
    f1 = cloud.integrate_cloud(pc_f)
    assert cloud.sync() & ~2 == 0
    bp = bev_params_from(dev, gp, f0, f1, f1 + 1, np.zeros(3), 0.0)
    o16, o64, _ = cloud.rasterise([bp], P, want_f64=True)
    cloud.sync()
    p2 = {k: v.copy() for k, v in pcs.items()}
    t2 = {k: ([x.copy() for x in v] if isinstance(v, list) else v.copy()) for k, v in trajs.items()}
    ref = orc.generate(p2, t2, gp, rot_ang=0.0, do_warping=True, return_f64=True)
    dbg = ref.pop('_debug')
    compare_planes(o16[0].cpu().numpy(), o64[0].cpu().numpy(),
                   {w: dbg[f'planes_f64_{w}'] for w in ('present', 'future', 'full')})
    # the sweep really covers all three paths
    counts = np.bincount(dbg['cells_present'][:, 0] * P + dbg['cells_present'][:, 1], minlength=P * P) + \
        np.bincount(dbg['cells_future'][:, 0] * P + dbg['cells_future'][:, 1], minlength=P * P)
    assert (counts == 15).any() and (counts == 16).any() and (counts == 32).any() and (counts >= 33).any()
    cloud.close()
