"""GPU parity of the C-ABI kernels (through pc_accumulation_lib_b200.device)
against the CPU oracle and the golden vectors produced by the reference.

Bars: pixel indices, masks, kept order, records and BEV cell
indices bit-exact; grid values <= 1e-5 relative before the float16 cast and
<= 1 float16 ulp after it.
"""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import oracle as orc                                  # noqa: E402
from pc_accumulation_lib_b200 import synth                        # noqa: E402
from tests.conftest import load_golden, unpack_bev, unpack_sem_pcs  # noqa: E402
from tests.golden import cases                                    # noqa: E402

pytestmark = pytest.mark.gpu

FLAG_INTENSITY_F32 = 2
F64_RTOL = 1e-5      # north-star tolerance on grid values (fp32-level), we are far inside
F64_ATOL = 1e-7


@pytest.fixture(scope='module')
def dev():
    from pc_accumulation_lib_b200 import device
    return device


def gen_params(bev_params, **kw):
    g = dict(sem_idxs=synth.SEM_IDXS, view_size=bev_params['view_size'],
             pixel_size=bev_params['pixel_size'], int_scaler=bev_params['int_scaler'],
             int_sep_scaler=bev_params['int_sep_scaler'],
             int_mid_threshold=bev_params['int_mid_threshold'],
             height_filter=bev_params['height_filter'], rgb_fill=0)
    g.update(kw)
    return g


def bev_params_from(dev, gp, fb, fs, fe, origin, rot_ang, dx=0., dy=0., zoom=1.):
    return dev.make_bev_params(fb, fs, fe, origin, orc.rotation_matrix_3d(rot_ang), dx, dy,
                               zoom * gp['view_size'], gp['height_filter'], gp['int_scaler'],
                               gp['int_sep_scaler'], gp['int_mid_threshold'],
                               gp.get('rgb_fill', 0), gp['sem_idxs'],
                               gp.get('elevation_mode', 'min') == 'max')


def compare_planes(planes16, planes64, want_bev_f64, exact=(0, 2, 3, 4, 5, 6)):
    """planes16/64: (3,7,P,P) arrays of ours; want_bev_f64: dict w -> (7,P,P)."""
    for wi, w in enumerate(('present', 'future', 'full')):
        want = want_bev_f64[w]
        got = planes64[wi]
        np.testing.assert_allclose(got, want, rtol=F64_RTOL, atol=F64_ATOL, err_msg=w)
        # integer-derived planes are exact
        for p in exact:
            np.testing.assert_array_equal(got[p], want[p], err_msg=f'{w} plane {p}')
        h = want.astype(np.float16)
        d = np.abs(planes16[wi].view(np.int16).astype(np.int32)
                   - h.view(np.int16).astype(np.int32))
        assert d.max() <= 1, (w, int(d.max()))


def window_cells(cloud, cells_t, fids, P):
    """(row, col) of the kept points of frames `fids`, in frame order."""
    out = []
    for f in fids:
        off, n = cloud.frame_offset(f), cloud.frame_count(f)
        c = cells_t[off:off + n].cpu().numpy()
        c = c[c >= 0]
        out.append(np.stack([c // P, c % P], axis=1))
    return np.concatenate(out) if out else np.zeros((0, 2), dtype=np.int64)


# ---------------------------------------------------------------------------
def test_project_and_gather_golden(dev):
    g = load_golden('kitti_project.npz')
    inp = cases.kitti_project_inputs()
    cloud = dev.DeviceCloud(capacity_pts=200_000, max_frames=8)
    u, v, m = cloud.project(inp['pc'], inp['P'], synth.KITTI_IMG_H, synth.KITTI_IMG_W)
    u, v, m = u.cpu().numpy(), v.cpu().numpy(), m.cpu().numpy().astype(bool)
    kept = np.flatnonzero(m)
    np.testing.assert_array_equal(kept, g['kept_idx'])
    np.testing.assert_array_equal(u[kept], g['u'])
    np.testing.assert_array_equal(v[kept], g['v'])
    # all N points, not only the kept ones, against the oracle
    ou, ov, om = orc.project(inp['pc'], inp['P'], synth.KITTI_IMG_H, synth.KITTI_IMG_W)
    np.testing.assert_array_equal(m, om)
    fits = (np.abs(ou) < 2**31 - 1) & (np.abs(ov) < 2**31 - 1)
    np.testing.assert_array_equal(u[fits], ou[fits])
    np.testing.assert_array_equal(v[fits], ov[fits])
    _, _, m30 = cloud.project(inp['pc'], inp['P'], synth.KITTI_IMG_H, synth.KITTI_IMG_W, 30.0)
    np.testing.assert_array_equal(np.flatnonzero(m30.cpu().numpy()), g['kept_idx_maxdepth30'])

    sem = cloud.gen_semantic_pc(inp['pc'], inp['rgb'], inp['P']).cpu().numpy()
    np.testing.assert_array_equal(sem[:, :4], inp['pc'][kept].astype(np.float64))
    np.testing.assert_array_equal(sem[:, 4:], g['gather_rgb'])
    sem = cloud.gen_semantic_pc(inp['pc'], inp['cls'][..., None], inp['P']).cpu().numpy()
    np.testing.assert_array_equal(sem[:, 4], g['gather_cls'])
    sem = cloud.gen_semantic_pc(inp['pc'], inp['prob'], inp['P']).cpu().numpy()
    np.testing.assert_array_equal(sem[:, 4:], g['gather_prob'])
    cloud.close()


def test_projection_edge_cases(dev):
    """depth == 0 -> -1e-6, points behind the camera, exact .5 pixels (half to
    even), NaN / inf coordinates."""
    P = synth.kitti_calib()['p_velo_frame']
    rng = np.random.default_rng(3)
    pc = rng.normal(0, 20, (5000, 4)).astype(np.float32)
    pc[:50, :3] = 0.0                           # projects to (t0,t1,t2): tiny depth
    pc[50:60, 0] = np.nan
    pc[60:70, 1] = np.inf
    Pe = np.array([[2., 0, 0, 0], [0, 2., 0, 0], [0, 0, 1., 0]])   # u = 2x/z: .5 pixels
    pe = np.zeros((400, 4), dtype=np.float32)
    pe[:, 0] = np.arange(400) * 0.25
    pe[:, 1] = np.arange(400) * 0.25
    pe[:, 2] = 1.0
    pe[::7, 2] = 0.0                             # depth exactly 0
    cloud = dev.DeviceCloud(capacity_pts=10_000, max_frames=8)
    for pts, mat, h, w in ((pc, P, 376, 1408), (pe, Pe, 150, 150)):
        u, v, m = cloud.project(pts, mat, h, w)
        with np.errstate(all='ignore'):
            ou, ov, om = orc.project(pts, mat, h, w)
        np.testing.assert_array_equal(m.cpu().numpy().astype(bool), om)
        k = np.flatnonzero(om)
        np.testing.assert_array_equal(u.cpu().numpy()[k], ou[k])
        np.testing.assert_array_equal(v.cpu().numpy()[k], ov[k])
    cloud.close()


def _run_kitti(dev, frames, gt, P, horizon, eager):
    acc = orc.KittiOracle(horizon, synth.kitti_calib()['p_velo_frame'], synth.KITTI_FILTERS,
                          gen_params(synth.kitti_bev_params(pixel_size=P)), use_gt_sem=gt)
    cloud = dev.DeviceCloud(capacity_pts=sum(f['pc'].shape[0] for f in frames) + 64,
                            max_frames=64)
    calib = synth.kitti_calib()
    for fr in frames:
        first, n_live = cloud.live_frames()
        if n_live > 0:
            cloud.rebase(fr['T'], eager=eager)
        if gt:
            cloud.integrate_gt(fr['pc'], fr['sem_gt'], synth.KITTI_FILTERS)
        else:
            cloud.integrate_frustum(fr['pc'], calib['p_velo_frame'], fr['rgb'], fr['cls'],
                                    synth.KITTI_FILTERS)
        ev = acc.integrate(fr['pc'], fr['rgb'], fr['cls'], fr['T'], fr['sem_gt'])
        cloud.evict(int(ev))
    assert cloud.sync() == 0
    return acc, cloud


@pytest.mark.parametrize('eager', [False, True])
@pytest.mark.parametrize('name,kw,gt', [
    ('kitti_seq.npz', dict(n_frames=9), False),
    ('kitti_seq_p256.npz', dict(n_frames=6), False),
    ('kitti_gtsem_seq.npz', dict(n_frames=5, n_beams=8, n_azimuth=300, config=3), True),
])
def test_kitti_sequence(dev, name, kw, gt, eager):
    g = load_golden(name)
    frames = cases.kitti_seq_inputs(use_gt_sem=gt, **kw)
    P = int(g['P'])
    acc, cloud = _run_kitti(dev, frames, gt, P, float(g['horizon']), eager)
    first, n_live = cloud.live_frames()
    want = unpack_sem_pcs(g)
    assert n_live == len(want) == len(acc.sem_pcs)
    for k in range(n_live):
        got = cloud.export_frame(first + k)
        np.testing.assert_array_equal(got, want[k], err_msg=f'frame {k}')   # bit-exact records
    # rasterise at the golden's present index
    p_idx = int(g['present_idx'])
    gp = acc.gen_params
    origin = np.array(acc.poses[p_idx])
    ego_present = np.array(acc.poses[:p_idx]) - origin
    rot = orc.heading_rot_ang(ego_present)
    bp = bev_params_from(dev, gp, first, first + p_idx, first + n_live, origin, rot)
    o16, o64, cells = cloud.rasterise([bp], P, want_f64=True, want_cells=True)
    cloud.sync()
    ref = acc.generate_bev(p_idx, return_f64=True)
    dbg = ref.pop('_debug')
    # lazy re-basing moves z by ~1e-13 m (composed vs sequential chain): the
    # elevation plane is then inside the tolerance instead of bit-equal
    compare_planes(o16[0].cpu().numpy(), o64[0].cpu().numpy(),
                   {w: dbg[f'planes_f64_{w}'] for w in ('present', 'future', 'full')},
                   exact=(0, 2, 3, 4, 5, 6) if eager else (0, 2, 3, 4, 5))
    # cell indices, bit-exact, in reference order
    fids = list(range(first, first + n_live))
    np.testing.assert_array_equal(window_cells(cloud, cells, fids[:p_idx], P), dbg['cells_present'])
    np.testing.assert_array_equal(window_cells(cloud, cells, fids[p_idx:], P), dbg['cells_future'])
    # and the float16 planes against the reference's own output
    want_bev = unpack_bev(g)
    got16 = o16[0].cpu().numpy()
    for wi, w in enumerate(('present', 'future', 'full')):
        for key, sl in (('road', 0), ('intensity', 1), ('dynamic', 5), ('elevation', 6)):
            d = np.abs(got16[wi, sl].view(np.int16).astype(np.int32)
                       - want_bev[f'{key}_{w}'].view(np.int16).astype(np.int32))
            assert d.max() <= 1, (key, w)
        np.testing.assert_array_equal(got16[wi, 2:5], want_bev[f'rgb_{w}'])
    st = cloud.raster_stats()
    assert st['binned'] > 0
    cloud.close()


@pytest.mark.parametrize('eager', [False, True])
def test_long_trajectory_frame_culling(dev, eager):
    """40 frames, 15 m apart: most frames lie outside the 80 m view and are culled by
    their bounding boxes; the result must not change."""
    P, n_frames = 64, 40
    rng = np.random.default_rng(17)
    frames = []
    for f in range(n_frames):
        seed = synth.seed_for(3, 500 + f)
        pc = synth.kitti_lidar(seed, 8, 200)
        yaw = rng.uniform(-0.05, 0.05)
        T = np.eye(4)
        T[:3, :3] = [[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]]
        T[:3, 3] = [15.0 + rng.uniform(-1, 1), rng.uniform(-0.5, 0.5), 0.]
        frames.append(dict(pc=pc, T=np.linalg.inv(T), rgb=None, cls=None,
                           sem_gt=synth.kitti_sem_gt(seed, pc.shape[0])))
    acc, cloud = _run_kitti(dev, frames, True, P, 1e9, eager)
    first, n_live = cloud.live_frames()
    for p_idx in (3, 20, 38):
        origin = np.array(acc.poses[p_idx])
        rot = orc.heading_rot_ang(np.array(acc.poses[:p_idx]) - origin)
        bp = bev_params_from(dev, acc.gen_params, first, first + p_idx, first + n_live, origin, rot)
        o16, o64, cells = cloud.rasterise([bp], P, want_f64=True, want_cells=True)
        cloud.sync()
        ref = acc.generate_bev(p_idx, return_f64=True)
        dbg = ref.pop('_debug')
        compare_planes(o16[0].cpu().numpy(), o64[0].cpu().numpy(),
                       {w: dbg[f'planes_f64_{w}'] for w in ('present', 'future', 'full')},
                       exact=(0, 2, 3, 4, 5, 6) if eager else (0, 2, 3, 4, 5))
        fids = list(range(first, first + n_live))
        np.testing.assert_array_equal(window_cells(cloud, cells, fids[:p_idx], P), dbg['cells_present'])
        np.testing.assert_array_equal(window_cells(cloud, cells, fids[p_idx:], P), dbg['cells_future'])
        st = cloud.raster_stats()
        assert 0 < st['binned'] < 0.5 * cloud.resident_points()
    cloud.close()


def test_nusc_sequence(dev):
    g = load_golden('nusc_seq.npz')
    scene = cases.nusc_seq_inputs()
    P = int(g['P'])
    gp = gen_params(synth.nusc_bev_params(pixel_size=P))
    acc = orc.NuscOracle(synth.NUSC_FILTERS, gp)
    cloud = dev.DeviceCloud(capacity_pts=sum(o['pc'].shape[0] for o in scene) + 64, max_frames=64)
    T_gw = None
    fids = []
    for o in scene:
        if T_gw is None:
            T_gw = np.linalg.inv(o['ego_at_lidar_ts'])
        T_ew = T_gw @ o['ego_at_lidar_ts']
        fids.append(cloud.integrate_records(o['pc'], o['pc_cam_idx'], o['images'], o['_semseg'],
                                            T_ew, synth.NUSC_FILTERS, 255.))
        acc.integrate(o, o['_semseg'])
    # dynamic flags: replay the oracle tracker's verdicts on the device
    pairs_f, pairs_i = [], []
    for ts, s in enumerate(acc.sem_pcs):
        for idx in np.unique(s[s[:, 9] == 1, 8]):
            pairs_f.append(fids[ts])
            pairs_i.append(int(idx))
    assert pairs_f, 'scene must contain dynamic instances'
    cloud.mark_dynamic(pairs_f, pairs_i)
    assert cloud.sync() == 0
    want = unpack_sem_pcs(g)
    for k, fid in enumerate(fids):
        np.testing.assert_array_equal(cloud.export_frame(fid), want[k], err_msg=f'frame {k}')
    p_idx = int(g['present_idx'])
    origin = np.array(acc.poses[p_idx])
    rot = orc.heading_rot_ang(np.array(acc.poses[:p_idx]) - origin)
    bp = bev_params_from(dev, gp, fids[0], fids[p_idx], fids[-1] + 1, origin, rot)
    o16, o64, cells = cloud.rasterise([bp], P, want_f64=True, want_cells=True)
    cloud.sync()
    ref = acc.generate_bev(p_idx, return_f64=True)
    dbg = ref.pop('_debug')
    compare_planes(o16[0].cpu().numpy(), o64[0].cpu().numpy(),
                   {w: dbg[f'planes_f64_{w}'] for w in ('present', 'future', 'full')})
    np.testing.assert_array_equal(window_cells(cloud, cells, fids[:p_idx], P), dbg['cells_present'])
    np.testing.assert_array_equal(window_cells(cloud, cells, fids[p_idx:], P), dbg['cells_future'])
    cloud.close()


def test_nusc_batched_integrate_equals_per_sweep(dev):
    """pcacc_integrate_records_batch (all sweeps in one launch) produces the same
    frames as one pcacc_integrate_records call per sweep."""
    g = load_golden('nusc_seq.npz')
    scene = cases.nusc_seq_inputs()
    cloud = dev.DeviceCloud(capacity_pts=2 * sum(o['pc'].shape[0] for o in scene), max_frames=64)
    T_gw = np.linalg.inv(scene[0]['ego_at_lidar_ts'])
    sweeps = [dict(pc=torch.from_numpy(o['pc']).cuda(), cam=torch.from_numpy(o['pc_cam_idx']).cuda(),
                   rgb=[torch.from_numpy(np.ascontiguousarray(i)).cuda() for i in o['images']],
                   sem=[torch.from_numpy(c).cuda() for c in o['_semseg']],
                   T=T_gw @ o['ego_at_lidar_ts']) for o in scene]
    first = cloud.integrate_records_batch(sweeps, synth.NUSC_FILTERS, 255.)
    assert cloud.sync() == 0
    want = unpack_sem_pcs(g)
    for k in range(len(scene)):
        got = cloud.export_frame(first + k)
        got[:, 9] = want[k][:, 9]            # dyn flags are the tracker's business
        np.testing.assert_array_equal(got, want[k], err_msg=f'frame {k}')
    # a per-sweep integrate after the batch lands densely behind the last frame
    nxt = cloud.integrate_records(scene[0]['pc'], scene[0]['pc_cam_idx'], scene[0]['images'],
                                  scene[0]['_semseg'], T_gw @ scene[0]['ego_at_lidar_ts'],
                                  synth.NUSC_FILTERS, 255.)
    cloud.sync()
    last = first + len(scene) - 1
    assert cloud.frame_offset(nxt) == (cloud.frame_offset(last) + cloud.frame_count(last) + 3) // 4 * 4
    got = cloud.export_frame(nxt)
    got[:, 9] = want[0][:, 9]
    np.testing.assert_array_equal(got, want[0])
    cloud.close()


@pytest.mark.parametrize('name,kw', [
    ('bev_direct.npz', {}),
    ('bev_direct_p128.npz', dict(n=20000, seed=78, P=128, view=51.2)),
])
def test_bev_direct(dev, name, kw):
    """Hand-made clouds with crowded cells (odd / even / large medians), 7 %
    dynamic points, height filter, explicit rot / trans / zoom; 2 variants in
    one batch (explicit angle + heading-aligned)."""
    g = load_golden(name)
    pcs, trajs, aug, gen = cases.bev_direct_inputs(**kw)
    P = gen['pixel_size']
    cloud = dev.DeviceCloud(capacity_pts=pcs['pc_full'].shape[0] + 64, max_frames=8)
    f0 = cloud.integrate_cloud(pcs['pc_present'])
    f1 = cloud.integrate_cloud(pcs['pc_future'])
    # these hand-made float64 intensities are not float32-representable: the ring
    # rounds them (6e-8 relative, inside the 1e-5 bar) and says so
    assert cloud.sync() & ~FLAG_INTENSITY_F32 == 0
    rot_head = orc.heading_rot_ang(trajs['ego_traj_present'])
    zero = np.zeros(3)
    bps = [bev_params_from(dev, gen, f0, f1, f1 + 1, zero, aug['rot_ang'], aug['trans_dx'],
                           aug['trans_dy'], aug['zoom_scalar']),
           bev_params_from(dev, gen, f0, f1, f1 + 1, zero, rot_head)]
    o16, o64, cells = cloud.rasterise(bps, P, want_f64=True, want_cells=True)
    cloud.sync()
    for vi, (a, prefix) in enumerate(((aug, 'bev_'), ({}, 'bevhead_'))):
        ref = orc.generate(*cases.copy_pcs_trajs(pcs, trajs), gen, return_f64=True, **a)
        dbg = ref.pop('_debug')
        compare_planes(o16[vi].cpu().numpy(), o64[vi].cpu().numpy(),
                       {w: dbg[f'planes_f64_{w}'] for w in ('present', 'future', 'full')})
        want_bev = unpack_bev(g, prefix)
        got16 = o16[vi].cpu().numpy()
        for wi, w in enumerate(('present', 'future', 'full')):
            np.testing.assert_array_equal(got16[wi, 2:5], want_bev[f'rgb_{w}'])
            np.testing.assert_array_equal(got16[wi, 0], want_bev[f'road_{w}'])
            np.testing.assert_array_equal(got16[wi, 5], want_bev[f'dynamic_{w}'])
            np.testing.assert_array_equal(got16[wi, 6], want_bev[f'elevation_{w}'])
    np.testing.assert_array_equal(window_cells(cloud, cells, [f0], P),
                                  orc.generate(*cases.copy_pcs_trajs(pcs, trajs), gen,
                                               return_f64=True, **aug)['_debug']['cells_present'])
    cloud.close()


def test_rasterise_is_deterministic_and_empty_ok(dev):
    pcs, trajs, aug, gen = cases.bev_direct_inputs(n=30000, seed=5, P=64, view=40.0)
    P = gen['pixel_size']
    cloud = dev.DeviceCloud(capacity_pts=60_000, max_frames=8)
    f0 = cloud.integrate_cloud(pcs['pc_present'])
    f1 = cloud.integrate_cloud(pcs['pc_future'])
    f2 = cloud.integrate_cloud(np.zeros((0, 10)))           # empty frame
    cloud.sync()
    bp = bev_params_from(dev, gen, f0, f1, f2 + 1, np.zeros(3), 0.3)
    a = cloud.rasterise([bp], P)[0].cpu().numpy().copy()
    for _ in range(3):
        b = cloud.rasterise([bp], P)[0].cpu().numpy()
        assert a.tobytes() == b.tobytes()
    # a view that contains no point at all: every cell takes the empty-cell values
    far = bev_params_from(dev, gen, f0, f1, f2 + 1, np.array([1e6, 1e6, 0.]), 0.0)
    e16, e64, _ = cloud.rasterise([far], P, want_f64=True)
    e64 = e64.cpu().numpy()[0]
    empty = np.zeros((0, 10))
    want = orc.raster_window(empty, P, gen['sem_idxs'], gen['int_scaler'], gen['int_sep_scaler'],
                             gen['int_mid_threshold'], 0, return_f64=True)
    for w in range(3):
        np.testing.assert_allclose(e64[w], want, rtol=1e-12, atol=0)
    cloud.close()
