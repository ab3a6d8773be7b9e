"""Round-2 GPU tests: the host-buffer boundary (pcacc_integrate_records_host: direct /
sparse staging), batch integrate into a nearly full and a wrapped ring, the tracker's batched
box transform, the event-recycled parameter arena, and BASELINE.json configs[0] / configs[4] at
full size against the oracle."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import oracle as orc                                  # noqa: E402
from pc_accumulation_lib_b200 import _lib, synth                  # noqa: E402
from tests.conftest import load_golden, unpack_bev, unpack_sem_pcs, assert_bev_equal  # noqa: E402
from tests.golden import cases                                    # noqa: E402
from tests.test_gpu_core import (bev_params_from, compare_planes, gen_params,  # noqa: E402
                                 window_cells)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    from pc_accumulation_lib_b200 import device
    return device


def test_host_staging_modes_give_identical_records(dev):
    """The same observations through (a) device tensors + the device gather (the tested
    primary), (b) pageable numpy arrays -> PCACC_STAGE_SPARSE, (c) page-locked numpy arrays ->
    PCACC_STAGE_DIRECT: bit-identical frames, equal to the reference golden."""
    g = load_golden('nusc_seq.npz')
    want = unpack_sem_pcs(g)
    scene = cases.nusc_seq_inputs()
    cap = sum(o['pc'].shape[0] for o in scene) + 64
    T_gw = np.linalg.inv(scene[0]['ego_at_lidar_ts'])
    clouds = {k: dev.DeviceCloud(cap, 64) for k in ('device', 'sparse', 'direct', 'auto_pageable')}
    for o in scene:
        T = T_gw @ o['ego_at_lidar_ts']
        clouds['device'].integrate_records(
            torch.from_numpy(o['pc']).cuda(), torch.from_numpy(o['pc_cam_idx']).cuda(),
            [torch.from_numpy(i).cuda() for i in o['images']],
            [torch.from_numpy(c).cuda() for c in o['_semseg']], T, synth.NUSC_FILTERS, 255.)
        clouds['sparse'].integrate_records_host(o['pc'], o['pc_cam_idx'], o['images'], o['_semseg'], T,
                                                synth.NUSC_FILTERS, 255., _lib.STAGE_SPARSE)
        assert clouds['sparse'].last_staging == _lib.STAGE_SPARSE
        clouds['auto_pageable'].integrate_records_host(o['pc'], o['pc_cam_idx'], o['images'],
                                                       o['_semseg'], T, synth.NUSC_FILTERS)
        assert clouds['auto_pageable'].last_staging == _lib.STAGE_SPARSE
        po = dev.pin_observation(o)
        assert _lib.load().pcacc_host_is_pinned(po['pc'].ctypes.data) == 1
        assert _lib.load().pcacc_host_is_pinned(o['pc'].ctypes.data) == 0
        clouds['direct'].integrate_records_host(po['pc'], po['pc_cam_idx'], po['images'], po['_semseg'],
                                                T, synth.NUSC_FILTERS, 255., _lib.STAGE_DIRECT)
        assert clouds['direct'].last_staging == _lib.STAGE_DIRECT
        # pageable arrays cannot be read in place
        with pytest.raises(_lib.PcaccError):
            clouds['direct'].integrate_records_host(o['pc'], o['pc_cam_idx'], o['images'], o['_semseg'],
                                                    T, synth.NUSC_FILTERS, 255., _lib.STAGE_DIRECT)
    for name, c in clouds.items():
        assert c.sync() == 0, name
        first, n_live = c.live_frames()
        assert n_live == len(scene)
        for k in range(n_live):
            got = c.export_frame(first + k)
            got[:, 9] = want[k][:, 9]            # dyn flags are the tracker's business
            np.testing.assert_array_equal(got, want[k], err_msg=f'{name} frame {k}')
    # narrow class maps and an out-of-image pixel through the sparse mode
    o = scene[0]
    T = T_gw @ o['ego_at_lidar_ts']
    c = clouds['sparse']
    c.reset()
    for dt in (np.uint8, np.int16, np.int32):
        fid = c.integrate_records_host(o['pc'], o['pc_cam_idx'], o['images'],
                                       [m.astype(dt) for m in o['_semseg']], T, synth.NUSC_FILTERS,
                                       255., _lib.STAGE_SPARSE)
        assert c.sync() == 0
        got = c.export_frame(fid)
        got[:, 9] = want[0][:, 9]
        np.testing.assert_array_equal(got, want[0], err_msg=str(dt))
    bad = o['pc'].copy()
    k = np.flatnonzero(o['pc_cam_idx'] == 0)[0]
    bad[k, 4] = 0.5
    c.integrate_records_host(bad, o['pc_cam_idx'], o['images'], o['_semseg'], T, synth.NUSC_FILTERS,
                             255., _lib.STAGE_SPARSE)
    assert c.sync() & _lib.FLAG_UV_OUT_OF_IMAGE
    # empty observation, and one where no camera sees anything
    c.reset()
    f0 = c.integrate_records_host(np.zeros((0, 7)), np.zeros(0, dtype=np.int64), o['images'], o['_semseg'],
                                  T, synth.NUSC_FILTERS)
    f1 = c.integrate_records_host(o['pc'], -np.ones_like(o['pc_cam_idx']), o['images'], o['_semseg'], T,
                                  synth.NUSC_FILTERS)
    assert c.sync() == 0
    assert c.frame_count(f0) == 0 and c.frame_count(f1) == 0
    for c in clouds.values():
        c.close()


def test_accumulator_with_pinned_and_pageable_observations():
    """The reference-facing class: pinned observations (kernel reads them in place) and
    pageable ones (sparse staging) give the golden records, tracker verdicts and BEV."""
    from pc_accumulation_lib_b200 import NuScenesOracleSemanticPointCloudAccumulator, pin_observation
    g = load_golden('nusc_seq.npz')
    want = unpack_sem_pcs(g)
    for pinned in (False, True):
        scene = cases.nusc_seq_inputs()
        if pinned:
            scene = [pin_observation(o) for o in scene]
        semseg = synth.SceneSemseg()
        for o in scene:
            semseg.register(o)
        acc = NuScenesOracleSemanticPointCloudAccumulator(
            semseg, synth.NUSC_FILTERS, synth.SEM_IDXS, None,
            synth.nusc_bev_params(pixel_size=int(g['P'])), ring_capacity_pts=400_000, ring_max_frames=64)
        acc.sync_each_integrate = not pinned          # both synchronisation policies
        for o in scene:
            acc.integrate([o])
        assert acc.cloud.last_staging == (_lib.STAGE_DIRECT if pinned else _lib.STAGE_SPARSE)
        np.testing.assert_array_equal(np.array(acc.poses), g['poses'])
        np.testing.assert_array_equal(np.array(acc.seg_dists), g['seg_dists'])
        assert list(g['dyn_instances']) == acc.dyn_instances
        for a, b in zip(acc.sem_pcs, want):
            np.testing.assert_array_equal(a, b)
        bev = acc.generate_bev(int(g['present_idx']), 1, True)[0]
        assert_bev_equal(bev, unpack_bev(g), exact=False)
        # reset(): a second scene on the same accumulator reproduces the first
        acc.reset()
        for o in scene:
            acc.integrate([o])
        for a, b in zip(acc.sem_pcs, want):
            np.testing.assert_array_equal(a, b)
        assert list(g['dyn_instances']) == acc.dyn_instances
        acc.cloud.close()


def test_tracker_box_transform_keeps_reference_bits():
    """The tracker moves all box centres of a sweep with one stacked matrix-vector product; the
    reference calls homo_transform on one centre at a time (nuscenes_oracle_sem_pc_accum.py:
    204-207), which takes numpy's matrix-vector path.  Same bits — on this host's BLAS."""
    from pc_accumulation_lib_b200 import NuScenesOracleSemanticPointCloudAccumulator as Acc
    rng = np.random.default_rng(3)
    acc = Acc.__new__(Acc)
    for _ in range(50):
        T = np.linalg.inv(rng.normal(size=(4, 4)) * 10.)
        T[3] = [0., 0., 0., 1.]
        acc.T_global_world = T
        c = rng.normal(size=(40, 3)) * 500.
        got = acc._boxes_to_world(list(c))
        want = np.stack([orc.homo_transform(T, c[i:i + 1])[0] for i in range(c.shape[0])])
        np.testing.assert_array_equal(got, want)


def test_batch_integrate_into_nearly_full_and_wrapped_ring(dev):
    """pcacc_integrate_records_batch when the capacity / wrap decision needs exact numbers in the
    middle of a batch: frames of the batch that are not enqueued yet must not be
    refreshed from the device table."""
    g = load_golden('nusc_seq.npz')
    want = unpack_sem_pcs(g)
    scene = cases.nusc_seq_inputs()
    n_in = scene[0]['pc'].shape[0]
    T_gw = np.linalg.inv(scene[0]['ego_at_lidar_ts'])
    sweeps = [dict(pc=torch.from_numpy(o['pc']).cuda(), cam=torch.from_numpy(o['pc_cam_idx']).cuda(),
                   rgb=[torch.from_numpy(np.ascontiguousarray(i)).cuda() for i in o['images']],
                   sem=[torch.from_numpy(c).cuda() for c in o['_semseg']],
                   T=T_gw @ o['ego_at_lidar_ts']) for o in scene]

    def check(cloud, first, ks):
        for j, k in enumerate(ks):
            got = cloud.export_frame(first + j)
            got[:, 9] = want[k][:, 9]
            np.testing.assert_array_equal(got, want[k], err_msg=f'frame {k}')

    # (a) the upper bounds of the second batch exceed the capacity, placed behind the first
    #     batch's EXACT end it fits: one sync for the whole batch, in-batch frames keep their
    #     host-side bounds
    cloud = dev.DeviceCloud(capacity_pts=int(5.2 * n_in), max_frames=64)
    f0 = cloud.integrate_records_batch(sweeps[:3], synth.NUSC_FILTERS, 255.)
    f1 = cloud.integrate_records_batch(sweeps[3:6], synth.NUSC_FILTERS, 255.)
    assert cloud.sync() == 0
    assert f1 == f0 + 3 and cloud.live_frames() == (f0, 6)
    check(cloud, f0, range(6))
    offs = [(cloud.frame_offset(f0 + k), cloud.frame_count(f0 + k)) for k in range(6)]
    for a in range(6):                                   # live frames never overlap
        for b in range(a + 1, 6):
            assert offs[a][0] + offs[a][1] <= offs[b][0] or offs[b][0] + offs[b][1] <= offs[a][0]
    # (b) evict the oldest frames; the next batch's second sweep wraps to offset 0
    cloud.evict(4)
    f2 = cloud.integrate_records_batch(sweeps[6:8], synth.NUSC_FILTERS, 255.)
    assert cloud.sync() == 0
    first, n_live = cloud.live_frames()
    assert (first, n_live) == (f0 + 4, 4)
    check(cloud, first, [4, 5, 6, 7])
    assert cloud.frame_offset(f2 + 1) == 0 and cloud.frame_offset(f2) > cloud.frame_offset(f2 - 1)
    # (c) a batch into the wrapped ring that cannot fit fails cleanly and leaves no phantom frames
    with pytest.raises(_lib.PcaccError) as e:
        cloud.integrate_records_batch(sweeps[:4], synth.NUSC_FILTERS, 255.)
    assert e.value.status == _lib.ERR_CAPACITY
    assert cloud.live_frames() == (first, 4)
    cloud.sync()
    check(cloud, first, [4, 5, 6, 7])
    cloud.close()


def test_parameter_arena_recycles_without_device_sync(dev):
    """Thousands of rasterise calls wrap the 4-segment parameter arena many times; results stay
    bit-identical to the first call (a block reused too early would corrupt the parameters)."""
    rng = np.random.default_rng(11)
    pc = np.zeros((20000, 10))
    pc[:, :2] = rng.normal(0., 15., (20000, 2))
    pc[:, 3] = rng.uniform(0., 1., 20000).astype(np.float32)
    pc[:, 4:7] = rng.integers(0, 256, (20000, 3))
    pc[:, 7] = rng.choice([0, 13, 2], 20000)
    cloud = dev.DeviceCloud(20064, 8)
    fid = cloud.integrate_cloud(pc)
    cloud.sync()
    gp = gen_params(synth.kitti_bev_params(pixel_size=64))
    bps = [bev_params_from(dev, gp, fid, fid + 1, fid + 1, np.zeros(3), 0.1 * k, 0.3 * k, -0.2 * k)
           for k in range(32)]
    first, _, _ = cloud.rasterise(bps, 64)
    first = first.clone()
    for _ in range(700):                                  # 700 x 6.9 KB > 4 MB of arena
        out, _, _ = cloud.rasterise(bps, 64)
    cloud.sync()
    assert torch.equal(out.view(torch.int16), first.view(torch.int16))
    cloud.close()


def test_non_multiple_of_32_grid(dev):
    """P*P not a multiple of 32 (P = 100, 50): the reference accepts any pixel_size."""
    rng = np.random.default_rng(21)
    from tests.test_gpu_edge import check_against_oracle, rand_cloud
    for P in (100, 50, 7):
        frames = [rand_cloud(rng, 30000, spread=12.), rand_cloud(rng, 20000, spread=12.)]
        cloud = dev.DeviceCloud(60000, 8)
        fids = [cloud.integrate_cloud(f) for f in frames]
        cloud.sync()
        gp = gen_params(synth.kitti_bev_params(pixel_size=P, view_size=40))
        check_against_oracle(dev, cloud, fids, frames, 1, gp, np.array([0.5, 0.25, 0.1]), 0.4,
                             exact=(0, 2, 3, 4, 5, 6))
        cloud.close()


def test_float64_cloud_must_be_float32_representable(dev):
    cloud = dev.DeviceCloud(1000, 8)
    seed = synth.seed_for(1, 3)
    pc = synth.kitti_lidar(seed, 4, 50)
    calib = synth.kitti_calib()
    u, v, m = cloud.project(pc.astype(np.float64), calib['p_velo_frame'], synth.KITTI_IMG_H, synth.KITTI_IMG_W)
    u2, v2, m2 = cloud.project(pc, calib['p_velo_frame'], synth.KITTI_IMG_H, synth.KITTI_IMG_W)
    assert torch.equal(u, u2) and torch.equal(m, m2)
    with pytest.raises(ValueError):
        cloud.project(pc.astype(np.float64) + 1e-9, calib['p_velo_frame'], synth.KITTI_IMG_H,
                      synth.KITTI_IMG_W)
    cloud.close()


def test_chunk_and_strip_reduce_kernels_agree(dev):
    """k_bev_reduce_chunk (default for float16-only output) against k_bev_reduce (strips) and the
    float64-output instantiation on sparse, crowded and empty grids: identical float16 bits.  Nine
    variants per call, so the candidate selection runs k_bev_classify_mv; the last pass repeats it
    with k_bev_classify (block per variant)."""
    from tests.test_gpu_edge import rand_cloud
    rng = np.random.default_rng(31)
    for P, n, spread in ((256, 40000, 30.), (64, 60000, 6.), (128, 3000, 40.), (1024, 200000, 25.)):
        frames = [rand_cloud(rng, n, spread=spread), rand_cloud(rng, n // 2, spread=spread)]
        cloud = dev.DeviceCloud(2 * n + 64, 8)
        fids = [cloud.integrate_cloud(f) for f in frames]
        cloud.sync()
        gp = gen_params(synth.kitti_bev_params(pixel_size=P, view_size=80))
        bps = [bev_params_from(dev, gp, fids[0], fids[1], fids[1] + 1, np.array([0.5, 0.25, 0.1]),
                               0.3 * k, 0.7 * k, -0.4 * k, 1.0 + 0.02 * k) for k in range(9)]
        bps[4].elevation_max = 1
        bps[3].rgb_fill = 37.0
        a, _, _ = cloud.rasterise(bps, P)
        cloud.set_option(_lib.OPT_REDUCE_STRIPS, 1)
        b, _, _ = cloud.rasterise(bps, P)
        cloud.set_option(_lib.OPT_REDUCE_STRIPS, 0)
        c, _, _ = cloud.rasterise(bps, P, want_f64=True)
        cloud.sync()
        cloud.set_option(_lib.OPT_CLASSIFY_SINGLE, 1)
        d, _, _ = cloud.rasterise(bps, P)
        cloud.set_option(_lib.OPT_CLASSIFY_SINGLE, 0)
        cloud.sync()
        assert torch.equal(a.view(torch.int16), b.view(torch.int16)), P
        assert torch.equal(a.view(torch.int16), c.view(torch.int16)), P
        assert torch.equal(a.view(torch.int16), d.view(torch.int16)), P
        cloud.close()


def test_multi_camera_staging_modes_match_oracle(dev):
    """Three cameras (the golden scene has one): the visible points are spread over cameras 0..2
    with different frames and class maps of a narrow dtype; device gather, sparse staging and
    in-place (page-locked) staging all give the oracle's records."""
    rng = np.random.default_rng(17)
    o = cases.nusc_seq_inputs()[2]
    pc, cam0 = o['pc'], o['pc_cam_idx']
    vis = np.flatnonzero(cam0 >= 0)
    cam = cam0.copy()
    cam[vis] = rng.integers(0, 3, vis.size)
    cam[vis[::11]] = -1                                  # some visible points dropped again
    h, w = o['images'][0].shape[:2]
    rgbs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(3)]
    sems = [rng.integers(0, 19, (h, w)).astype(np.int16) for _ in range(3)]
    T = np.linalg.inv(cases.nusc_seq_inputs()[0]['ego_at_lidar_ts']) @ o['ego_at_lidar_ts']
    want, _ = orc.nusc_obs2sem(pc, cam, rgbs, [s.astype(np.int64) for s in sems], T, synth.NUSC_FILTERS)
    assert want.shape[0] > 200
    cloud = dev.DeviceCloud(pc.shape[0] * 4 + 64, 16)
    f0 = cloud.integrate_records(torch.from_numpy(pc).cuda(), torch.from_numpy(cam).cuda(),
                                 [torch.from_numpy(r).cuda() for r in rgbs],
                                 [torch.from_numpy(s).cuda() for s in sems], T, synth.NUSC_FILTERS, 255.)
    f1 = cloud.integrate_records_host(pc, cam, rgbs, sems, T, synth.NUSC_FILTERS, 255., _lib.STAGE_SPARSE)
    pin = [dev.pinned_like(a) for a in (pc, cam)]
    f2 = cloud.integrate_records_host(pin[0], pin[1], [dev.pinned_like(r) for r in rgbs],
                                      [dev.pinned_like(s) for s in sems], T, synth.NUSC_FILTERS, 255.,
                                      _lib.STAGE_DIRECT)
    # a camera index beyond the list is "seen by no camera" on every path
    cam_bad = cam.copy()
    cam_bad[vis[:5]] = 7
    f3 = cloud.integrate_records_host(pc, cam_bad, rgbs, sems, T, synth.NUSC_FILTERS, 255., _lib.STAGE_SPARSE)
    assert cloud.sync() == 0
    for f in (f0, f1, f2):
        np.testing.assert_array_equal(cloud.export_frame(f), want)
    want_bad, _ = orc.nusc_obs2sem(pc, np.where(cam_bad == 7, -1, cam_bad), rgbs,
                                   [s.astype(np.int64) for s in sems], T, synth.NUSC_FILTERS)
    np.testing.assert_array_equal(cloud.export_frame(f3), want_bad)
    cloud.close()
