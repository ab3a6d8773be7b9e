"""The drop-in Python API (same class / method names as the reference) against
the golden outputs the UNMODIFIED reference produced for the same seeded inputs
(tests/golden/make_golden.py).  These read like a test of the reference itself:
construct the accumulator, integrate(), generate_bev(), compare the dict.
"""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import oracle as orc                                               # noqa: E402
from pc_accumulation_lib_b200 import synth                                     # noqa: E402
from tests.conftest import assert_bev_equal, load_golden, unpack_bev, unpack_sem_pcs  # noqa: E402
from tests.golden import cases                                                 # noqa: E402

pytestmark = pytest.mark.gpu


def _kitti_acc(horizon, P, semseg, use_gt_sem, **kw):
    from pc_accumulation_lib_b200 import Kitti360SemanticPointCloudAccumulator
    return Kitti360SemanticPointCloudAccumulator(
        horizon, synth.kitti_calib(), 1.0, semseg, synth.KITTI_FILTERS, synth.SEM_IDXS,
        use_gt_sem, synth.kitti_bev_params(pixel_size=P), ring_capacity_pts=400_000,
        ring_max_frames=64, **kw)


@pytest.mark.parametrize('eager', [False, True])
@pytest.mark.parametrize('name,kw,gt', [
    ('kitti_seq.npz', dict(n_frames=9), False),
    ('kitti_seq_p256.npz', dict(n_frames=6), False),
    ('kitti_gtsem_seq.npz', dict(n_frames=5, n_beams=8, n_azimuth=300, config=3), True),
])
def test_kitti_accumulator_matches_reference(name, kw, gt, eager):
    g = load_golden(name)
    frames = cases.kitti_seq_inputs(use_gt_sem=gt, **kw)
    sem = None if gt else synth.FakeSemseg([f['cls'] for f in frames])
    acc = _kitti_acc(float(g['horizon']), int(g['P']), sem, gt)
    acc.eager_rebase = eager
    evicted, n_kept = [], []
    for fr in frames:
        evicted.append(acc.integrate([(fr['rgb'], fr['pc'], fr['sem_gt'], fr['T'])]))
        n_kept.append(acc.sem_pcs[-1].shape[0])
    np.testing.assert_array_equal(evicted, g['evicted'])
    np.testing.assert_array_equal(n_kept, g['n_kept'])
    np.testing.assert_array_equal(np.array(acc.poses), g['poses'])
    np.testing.assert_array_equal(np.array(acc.seg_dists), g['seg_dists'])
    want = unpack_sem_pcs(g)
    assert len(acc.sem_pcs) == len(want)
    for a, b in zip(acc.sem_pcs, want):
        np.testing.assert_array_equal(a, b)          # (M,10) float64 records, bit-exact
    bev = acc.generate_bev(int(g['present_idx']), 1, True)[0]
    assert_bev_equal(bev, unpack_bev(g), exact=False)   # fp16 planes <= 1 ulp, trajs exact
    for w in ('present', 'future', 'full'):             # index-derived planes: bit-equal
        for k in ('road', 'rgb', 'dynamic'):
            np.testing.assert_array_equal(bev[f'{k}_{w}'], unpack_bev(g)[f'{k}_{w}'])
    # generate_bev must not change the accumulated state
    for a, b in zip(acc.sem_pcs, want):
        np.testing.assert_array_equal(a, b)


def test_kitti_pose_source_and_errors():
    frames = cases.kitti_seq_inputs(n_frames=3)
    sem = synth.FakeSemseg([f['cls'] for f in frames])
    acc = _kitti_acc(1e9, 32, sem, False)
    with pytest.raises(NotImplementedError):          # no ICP here: the pose is an input
        acc.integrate([(frames[0]['rgb'], frames[0]['pc'], None)])
    acc.pose_source = iter([f['T'] for f in frames])
    for fr in frames:
        acc.integrate([(fr['rgb'], fr['pc'], None)])
    assert len(acc.poses) == 3
    with pytest.raises(UnboundLocalError):            # reference: bev_generator.py:111-123
        acc.generate_bev(1, 1, False)
    with pytest.raises(ValueError):                   # np.concatenate([]) in the reference
        acc.generate_bev(0, 1, True)
    with pytest.raises(NotImplementedError):
        from pc_accumulation_lib_b200 import Kitti360SemanticPointCloudAccumulator
        bp = synth.kitti_bev_params(pixel_size=32)
        bp['type'] = 'rgb'
        Kitti360SemanticPointCloudAccumulator(1e9, synth.kitti_calib(), 1., sem,
                                              synth.KITTI_FILTERS, synth.SEM_IDXS, False, bp)
    # bev_num > 1: a batch of identical heading-aligned BEVs (no augmentation configured)
    bevs = acc.generate_bev(2, 3, True)
    assert len(bevs) == 3
    for b in bevs[1:]:
        np.testing.assert_array_equal(b['rgb_full'], bevs[0]['rgb_full'])


def test_kitti_helpers_match_reference():
    g = load_golden('kitti_project.npz')
    inp = cases.kitti_project_inputs()
    acc = _kitti_acc(1e9, 32, None, True)
    pc5 = np.concatenate([inp['pc'], np.arange(inp['pc'].shape[0], dtype=np.float32)[:, None]], 1)
    img = acc.velo2img(pc5, inp['P'], synth.KITTI_IMG_H, synth.KITTI_IMG_W)
    np.testing.assert_array_equal(img[:, 4].astype(np.int32), g['kept_idx'])
    np.testing.assert_array_equal(img[:, 5].astype(np.int32), g['u'])
    np.testing.assert_array_equal(img[:, 6].astype(np.int32), g['v'])
    sem = acc.gen_semantic_pc(inp['pc'], inp['cls'][..., None], inp['P'])
    np.testing.assert_array_equal(sem[:, 4], g['gather_cls'])
    kept = acc.filter_semseg_pc(np.concatenate([sem, sem[:, -1:]], axis=1))
    assert not np.isin(kept[:, -1], synth.KITTI_FILTERS).any()
    assert kept.shape[0] == (~np.isin(sem[:, -1], synth.KITTI_FILTERS)).sum()
    # whole rows and their order against the oracle's restatements (sem_pc_accum.py:317-321, 367-402)
    np.testing.assert_array_equal(img, orc.velo2img(pc5, inp['P'], synth.KITTI_IMG_H, synth.KITTI_IMG_W))
    near = acc.velo2img(inp['pc'], inp['P'], synth.KITTI_IMG_H, synth.KITTI_IMG_W, max_depth=12.)
    np.testing.assert_array_equal(near, orc.velo2img(inp['pc'], inp['P'], synth.KITTI_IMG_H, synth.KITTI_IMG_W, 12.))
    assert 0 < near.shape[0] < img.shape[0] and near.dtype == np.float64 and near.shape[1] == 6
    sem8 = np.concatenate([sem, sem[:, -1:]], axis=1)
    np.testing.assert_array_equal(kept, orc.filter_semseg_pc(sem8, synth.KITTI_FILTERS))
    assert acc.velo2img(inp['pc'][:0], inp['P'], 10, 10).shape == (0, 6)


def test_nuscenes_oracle_accumulator_matches_reference():
    from pc_accumulation_lib_b200 import NuScenesOracleSemanticPointCloudAccumulator
    g = load_golden('nusc_seq.npz')
    scene = cases.nusc_seq_inputs()
    semseg = synth.SceneSemseg()
    for o in scene:
        semseg.register(o)
    acc = NuScenesOracleSemanticPointCloudAccumulator(
        semseg, synth.NUSC_FILTERS, synth.SEM_IDXS, None,
        synth.nusc_bev_params(pixel_size=int(g['P'])), ring_capacity_pts=400_000,
        ring_max_frames=64)
    for o in scene:
        assert acc.integrate([o]) is None
    np.testing.assert_array_equal(np.array(acc.poses), g['poses'])
    np.testing.assert_array_equal(np.array(acc.seg_dists), g['seg_dists'])
    assert list(g['dyn_instances']) == acc.dyn_instances
    want = unpack_sem_pcs(g)
    for a, b in zip(acc.sem_pcs, want):
        np.testing.assert_array_equal(a, b)          # incl. retroactive dyn flags
    bev = acc.generate_bev(int(g['present_idx']), 1, True)[0]
    assert_bev_equal(bev, unpack_bev(g), exact=False)
    assert len(bev['trajs_full']) > 1                 # ego + dynamic objects
    assert acc.ego_global_xs[0] == scene[0]['ego_global_x']
    # uv outside the image: the reference asserts in pts_feat_from_img
    bad = dict(scene[0])
    bad['pc'] = scene[0]['pc'].copy()
    k = np.flatnonzero(bad['pc_cam_idx'] == 0)[0]
    bad['pc'][k, 4] = 0.5
    with pytest.raises(AssertionError):
        acc.integrate([bad])
    with pytest.raises(NotImplementedError):
        NuScenesOracleSemanticPointCloudAccumulator(
            semseg, synth.NUSC_FILTERS, synth.SEM_IDXS, True, synth.nusc_bev_params())


@pytest.mark.parametrize('name,kw', [
    ('bev_direct.npz', {}),
    ('bev_direct_p128.npz', dict(n=20000, seed=78, P=128, view=51.2)),
])
def test_sem_bev_generator_on_host_clouds(name, kw):
    """BEVGenerator.generate(pcs, trajs, ...) handed numpy clouds, as the
    reference's callers do."""
    from pc_accumulation_lib_b200 import SemBEVGenerator
    g = load_golden(name)
    pcs, trajs, aug, gen = cases.bev_direct_inputs(**kw)
    bg = SemBEVGenerator(gen['sem_idxs'], gen['view_size'], gen['pixel_size'], 0., 0., False,
                         gen['int_scaler'], gen['int_sep_scaler'], gen['int_mid_threshold'],
                         gen['height_filter'], gen['rgb_fill'])
    bev = bg.generate(*cases.copy_pcs_trajs(pcs, trajs), **aug)
    assert_bev_equal(bev, unpack_bev(g, 'bev_'), exact=False)
    bev = bg.generate(*cases.copy_pcs_trajs(pcs, trajs))
    assert_bev_equal(bev, unpack_bev(g, 'bevhead_'), exact=False)
    assert bev['rgb_present'].shape == (3, gen['pixel_size'], gen['pixel_size'])
    assert bev['road_full'].dtype == np.float16
    # generate_bev() on pre-gridded clouds (reference: sem_bev.py:36-262)
    want = unpack_bev(g, 'bev_')
    P = gen['pixel_size']
    grids = {}
    for w in ('present', 'future', 'full'):
        ij, z = g[f'grid_ij_{w}'], g[f'grid_z_{w}']
        # rebuild the pre-processed cloud: grid xy, transformed z, untouched attributes
        from oracle import oracle as orc
        pre = orc.preprocess_pc(pcs[f'pc_{w}'], orc.rotation_matrix_3d(aug['rot_ang']),
                                aug['trans_dx'], aug['trans_dy'],
                                aug['zoom_scalar'] * gen['view_size'], P, gen['height_filter'])
        np.testing.assert_array_equal(pre[:, :2].astype(np.int32), ij)
        grids[w] = pre
    bev2 = bg.generate_bev(grids['present'], grids['future'], grids['full'], want['trajs_present'],
                           want['trajs_future'], want['trajs_full'])
    assert_bev_equal(bev2, want, exact=False)


def test_rgb_bev_generator():
    from pc_accumulation_lib_b200 import RGBBEVGenerator
    from oracle import oracle as orc
    rng = np.random.default_rng(11)
    P = 32
    def cloud(n):
        pc = np.zeros((n, 7))
        pc[:, 0:2] = rng.integers(0, P, (n, 2))
        pc[:, 2] = rng.normal(0, 1, n)
        pc[:, 4:7] = rng.integers(0, 256, (n, 3))
        return pc
    a, b = cloud(3000), cloud(1500)
    bg = RGBBEVGenerator(40., P, rgb_fill=0)
    out = bg.generate_bev(a, b, np.zeros((3, 3)), np.ones((2, 3)))
    assert set(out) == {'rgb_present', 'rgb_future', 'poses_present', 'poses_future'}
    for key, pc in (('rgb_present', a), ('rgb_future', b)):
        cell = (P - 1 - pc[:, 1].astype(int)) * P + pc[:, 0].astype(int)
        for ch in range(3):
            med = orc._segment_median(cell, pc[:, 4 + ch], P * P)
            med[np.isnan(med)] = 0
            want = (med / 255.).reshape(P, P).astype(np.float16)
            np.testing.assert_array_equal(out[key][ch], want)


def test_rgb_bev_generator_matches_reference_golden():
    """RGBBEVGenerator.generate_bev without and with the polynomial warp against the unmodified
    reference (tests/golden/rgb_bev.npz): planes bit-equal (medians / 255 in float16), warped poses equal."""
    import random
    from pc_accumulation_lib_b200 import RGBBEVGenerator
    g = load_golden('rgb_bev.npz')
    c = synth.rgb_bev_inputs()
    bg = RGBBEVGenerator(40., c['P'])

    def call(warp):
        return bg.generate_bev(c['pc_present'].copy(), c['pc_future'].copy(), c['poses_present'].copy(),
                               c['poses_future'].copy(), do_warping=warp)
    out = call(False)
    for k in ('rgb_present', 'rgb_future', 'poses_present', 'poses_future'):
        np.testing.assert_array_equal(np.asarray(out[k]), g[f'plain_{k}'], err_msg=k)
    for s in g['seeds']:
        bg.rng, bg.py_rng = np.random.RandomState(int(s)), random.Random(int(s))
        out = call(True)
        for k in ('rgb_present', 'rgb_future', 'poses_present', 'poses_future'):
            np.testing.assert_array_equal(np.asarray(out[k]), g[f'warp{int(s)}_{k}'], err_msg=f'{k} seed {s}')
    assert not np.array_equal(g['warp5_rgb_present'], g['plain_rgb_present'])     # the warp does something


def test_missing_cuda_or_library_fails_loudly(monkeypatch):
    from pc_accumulation_lib_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libpcacc.so')
    with pytest.raises(ImportError):
        _lib.load()


def test_sem_bev_generator_polynomial_warp():
    """do_warp=True: the warp parameters come from injected RNG streams
    seeded like the reference's global ones; grids are gathered on the device, trajectories
    warped on the host — both equal the reference's output."""
    import random
    from pc_accumulation_lib_b200 import SemBEVGenerator
    g = load_golden('bev_warp.npz')
    pcs, trajs, aug, gen = cases.bev_direct_inputs()
    bg = SemBEVGenerator(gen['sem_idxs'], gen['view_size'], gen['pixel_size'], 0., 0., True,
                         gen['int_scaler'], gen['int_sep_scaler'], gen['int_mid_threshold'],
                         gen['height_filter'], gen['rgb_fill'])
    for s in g['seeds']:
        bg.rng, bg.py_rng = np.random.RandomState(int(s)), random.Random(int(s))
        bev = bg.generate(*cases.copy_pcs_trajs(pcs, trajs), **aug)
        assert_bev_equal(bev, unpack_bev(g, f'bev{int(s)}_'), exact=False)
    # the stand-alone dense warp helper on a float64 stack
    rng = np.random.default_rng(2)
    maps = rng.random((5, 32, 32))
    a_1, a_2 = bg.cal_warp_params(19.5, 16, 31)
    b_1, b_2 = bg.cal_warp_params(12.25, 16, 31)
    from oracle import oracle as orc
    np.testing.assert_array_equal(bg.warp_dense_probmaps(maps, a_1, a_2, b_1, b_2),
                                  orc.warp_dense(maps, a_1, a_2, b_1, b_2))


def test_input_side_kernels_match_reference_golden():
    """pcacc_assign_boxes / pcacc_project_cameras through the mirror of
    datasets/nuscenes_utils.py, against the reference's own outputs (bit-exact)."""
    from pc_accumulation_lib_b200.datasets import nuscenes_utils as nu
    g = load_golden('input_side.npz')
    c = cases.input_side_inputs()
    for tag, pts in (('f64', c['pc']), ('f32', c['pc_f32'])):
        box, cnt = nu.assign_points_to_boxes(pts, c['boxes'], c['sizes'], c['tolerance'])
        np.testing.assert_array_equal(box, g[f'box_{tag}'])
        np.testing.assert_array_equal(cnt, g[f'cnt_{tag}'])
    b = 4
    m = nu.find_points_in_box(c['pc'], c['boxes'][b], c['sizes'][b], c['tolerance'])
    assert m.dtype == bool and int(m.sum()) == int(g['cnt_f64'][b])
    uv, idx = nu.project_to_cameras(c['pc'], c['glob_from_ego'], c['cams'])
    np.testing.assert_array_equal(idx, g['pc_cam_idx'])
    np.testing.assert_array_equal(uv, g['pc_uv'])
    # one camera, points already in its frame (NuScenesCamera.project_pts3d)
    j = 3
    pc_cam = nu.homo_transform(np.linalg.inv(c['cams'][j]['glob_from_self']),
                               nu.homo_transform(c['glob_from_ego'], c['pc']))
    uv1, mask1 = nu.project_pts3d(pc_cam, c['cams'][j]['cam_K'], c['cams'][j]['img_wh'])
    np.testing.assert_array_equal(mask1, g[f'mask_cam{j}'])
    np.testing.assert_array_equal(uv1[mask1], g[f'uv_cam{j}'][mask1])
    # empty inputs and more boxes than one launch holds
    box, cnt = nu.assign_points_to_boxes(np.zeros((0, 3)), c['boxes'], c['sizes'], 0.01)
    assert box.shape == (0,) and not cnt.any()
    many = c['boxes'] * 12
    box, cnt = nu.assign_points_to_boxes(c['pc'], many, c['sizes'] * 12, c['tolerance'])
    want = np.where(g['box_f64'] >= 0, g['box_f64'] + 11 * len(c['boxes']), -1)
    np.testing.assert_array_equal(box, want)
    np.testing.assert_array_equal(cnt, np.tile(g['cnt_f64'], 12))


def test_standalone_helpers_match_reference_golden():
    """pts_feat_from_img (both branches) and static_obj_partitioning_by_elev through the drop-in
    functions against outputs of the unmodified reference (tests/golden/helpers.npz)."""
    from pc_accumulation_lib_b200.bev_generator import SemBEVGenerator
    from pc_accumulation_lib_b200.datasets import nuscenes_utils as nu
    g = load_golden('helpers.npz')
    c = synth.helper_inputs()
    near = nu.pts_feat_from_img(c['uv'], c['img4'], 'nearest')
    assert near.dtype == g['nearest4'].dtype
    np.testing.assert_array_equal(near, g['nearest4'])
    b = nu.pts_feat_from_img(c['uv'], c['img2d'], 'bilinear')
    assert b.shape == g['bilinear2d'].shape
    assert np.array_equal(b, g['bilinear2d'], equal_nan=True)
    # per-channel blend of a multi-channel image == the 2-D result of each channel
    b4 = nu.pts_feat_from_img(c['uv'], c['img4'].astype(np.float64), 'bilinear')
    for ch in range(4):
        want = orc.pts_feat_from_img(c['uv'], np.ascontiguousarray(c['img4'][:, :, ch].astype(np.float64)), 'bilinear')
        assert np.array_equal(b4[:, ch], want, equal_nan=True), ch
    bad = c['uv'].copy()
    bad[3, 0] = 0.5
    with pytest.raises(AssertionError):
        nu.pts_feat_from_img(bad, c['img2d'], 'nearest')
    with pytest.raises(AssertionError):
        nu.pts_feat_from_img(c['uv'], c['img2d'], 'cubic')

    gen = SemBEVGenerator(synth.SEM_IDXS, 40., c['P'])
    pc = c['pc'].copy()
    st, dy, el, ob = gen.static_obj_partitioning_by_elev(pc, c['elev_thresh'])
    for got, key in ((pc, 'pc_after'), (st, 'pc_static'), (dy, 'pc_dynamic'), (el, 'elevmap'), (ob, 'obs')):
        np.testing.assert_array_equal(got, g[key], err_msg=key)
    # negative grid coordinates wrap once like numpy's indexing; beyond that the reference raises
    pc2 = c['pc'].copy()
    pc2[0, 0] = -3.
    want = pc2.copy()
    orc.static_obj_partitioning_by_elev(want, c['P'], c['elev_thresh'])
    gen.static_obj_partitioning_by_elev(pc2, c['elev_thresh'])
    np.testing.assert_array_equal(pc2, want)
    pc2[0, 0] = float(c['P'])
    with pytest.raises(IndexError):
        gen.static_obj_partitioning_by_elev(pc2, c['elev_thresh'])


def test_per_step_methods_match_reference_golden():
    """Every per-step public method of the generators / accumulator that pcacc_rasterise fuses, called
    stand-alone with the reference's arguments, against outputs of the unmodified reference
    (tests/golden/standalone.npz): bit-exact except the weighted histogram and the intensity map (float64
    atomics: summation order) and exp (device libm vs numpy, <= 2 ulp)."""
    import torch
    from pc_accumulation_lib_b200 import SemanticPointCloudAccumulator
    from pc_accumulation_lib_b200.bev_generator import SemBEVGenerator, RGBBEVGenerator
    g = load_golden('standalone.npz')
    c = synth.standalone_inputs()
    P, view = c['P'], c['view']
    S = synth.SEM_IDXS
    gen = SemBEVGenerator(S, 40., P, height_filter=c['height_filter'], rgb_fill=7)
    eq = np.testing.assert_array_equal

    eq(gen.crop_view(c['pc'].copy(), view), g['crop_view'])
    eq(SemBEVGenerator.crop_view(c['pc'].copy(), view), g['crop_view'])         # static in the reference
    eq(gen.geometric_transform(c['pc'].copy(), c['rot_ang'], c['dx'], c['dy'], view), g['geometric_transform'])
    pc_g, trajs = gen.preprocess_pc_and_trajs(c['pc'].copy(), [], c['rot_ang'], c['dx'], c['dy'], view)
    eq(pc_g, g['preprocess_pc'])
    assert trajs == []
    # rows of another width (the reference takes (N, 3 + M)) and an empty cloud
    eq(gen.crop_view(c['pc'][:, :4].copy(), view), g['crop_view'][:, :4])
    assert gen.crop_view(np.zeros((0, 10)), view).shape == (0, 10)

    eq(gen.gen_gridmap_count_map(c['edges']), g['count_map'])
    np.testing.assert_allclose(gen.gen_gridmap_count_map(c['edges'], weights=c['weights']),
                               g['count_map_weighted'], rtol=0, atol=1e-12)
    eq(gen.gen_sem_probmap(c['edges'], ['road']), g['sem_probmap_road'])
    eq(gen.gen_sem_probmap(c['edges'], ['car', 'truck', 'bus', 'motorcycle']), g['sem_probmap_veh'])
    np.testing.assert_allclose(gen.gen_intensity_map(c['edges'], 'road'), g['intensity_map'], rtol=0, atol=1e-13)
    a, b = gen.partition_semantic_pc(c['edges'], [S['car'], S['bus']], 7)
    assert np.array_equal(a, g['partition_sel'], equal_nan=True)
    assert np.array_equal(b, g['partition_rest'], equal_nan=True)
    a0, b0 = gen.partition_semantic_pc(c['edges'], [], 7)
    assert a0.shape == (0, 10) and np.array_equal(b0, c['edges'], equal_nan=True)
    post = gen.dirichlet_dist_expectation([m.copy() for m in c['maps']], obs_weight=2)
    assert isinstance(post, list) and len(post) == 3
    eq(np.stack(post), g['dirichlet'])
    for key, args in (('road_marking', (1., 30., 0.12)), ('road_marking_kitti', (20., 20., 0.5))):
        got = gen.road_marking_transform(c['vals'].copy(), *args)
        np.testing.assert_allclose(got, g[key], rtol=1e-15, atol=0)
        assert got.max() <= 1.
    np.testing.assert_allclose(gen.sigmoid(c['vals']), g['sigmoid'], rtol=1e-15, atol=0)
    assert gen.extract_aug_dict({'max_translation_radius': 3., 'zoom_threshold': .2}) == (3., .2)

    elev, obs = gen.get_elevation_map(c['grid'])
    eq(elev, g['elevmap'])
    eq(obs, g['elev_obs'])
    assert obs.dtype == np.bool_
    bad = c['grid'].copy()
    bad[7, 0] = P
    with pytest.raises(IndexError):
        gen.get_elevation_map(bad)
    eq(np.stack(gen.get_rgb_maps(c['grid'])), g['rgb_maps'])
    assert gen.rgb_fill == 7
    rgen = RGBBEVGenerator(40., P, rgb_fill=7)
    eq(np.stack(rgen.get_rgb_maps(c['grid'][:, :7])), g['rgb_maps'])

    vf = SemanticPointCloudAccumulator.velo2frame
    np.testing.assert_allclose(vf(c['pts32'][:, :3], c['P34']), g['velo2frame32'], rtol=1e-13, atol=1e-12)
    np.testing.assert_allclose(vf(c['pc'][4:, :3], c['P34']), g['velo2frame64'], rtol=1e-13, atol=1e-12)
    eq(vf(c['pts32'][:, :3], c['P34']), orc.velo2frame(c['pts32'][:, :3], c['P34']))

    # CUDA tensors in -> CUDA tensors out, no host round trip
    t = torch.from_numpy(c['edges']).cuda()
    m = gen.gen_sem_probmap(t, ['road'])
    assert isinstance(m, torch.Tensor) and m.is_cuda
    eq(m.cpu().numpy(), g['sem_probmap_road'])


def test_per_step_operators_edge_cases():
    """Empty clouds, every step switched off, rows of odd width, all-selected / none-selected partitions and
    the oracle on random inputs of other sizes for the stand-alone operators (pcacc_preprocess_pc,
    pcacc_cell_stats, pcacc_partition_semantic_pc, pcacc_dirichlet_expectation, pcacc_road_marking,
    pcacc_elevation_map, pcacc_velo2frame)."""
    import torch
    from pc_accumulation_lib_b200.device import DeviceCloud
    ops = DeviceCloud(1024, 8)
    rng = np.random.default_rng(3)
    eq = np.testing.assert_array_equal
    # empty inputs
    assert ops.preprocess_pc(np.zeros((0, 10)), np.eye(3), 0., 0., 10., 1., 10., 16).shape == (0, 10)
    m = ops.cell_stats(np.zeros((0, 10)), 8, [1], 7, weight_col=3)
    assert all(float(v.abs().sum()) == 0. for v in m.values())
    m = ops.cell_stats(np.zeros((0, 10)), 8, [1], 7, finish=1, want=('sel', 'rest'))
    eq(m['sel'].cpu().numpy(), np.full((8, 8), 0.5))                   # uniform prior alone
    a, b = ops.partition_semantic_pc(np.zeros((0, 10)), [1], 7)
    assert a.shape == (0, 10) and b.shape == (0, 10)
    assert ops.velo2frame(np.zeros((0, 3), dtype=np.float32), np.eye(3, 4)).shape == (0, 3)
    assert ops.road_marking(np.zeros((0,))).shape == (0,)
    elev, obs = ops.elevation_map(np.zeros((0, 3)), 8)
    assert not obs.any() and float(elev.abs().sum()) == 0.
    # every step off: the identity, rows of width 3 and 13
    for cols in (3, 13):
        pc = rng.normal(0., 5., (1000, cols))
        eq(ops.preprocess_pc(pc).cpu().numpy(), pc)
    # several sizes against the oracle (tile boundaries of the look-back compaction: 255, 256, 257, 70001)
    for n in (1, 255, 256, 257, 70001):
        pc = np.zeros((n, 10))
        pc[:, :2] = rng.uniform(-30., 30., (n, 2))
        pc[:, 2] = rng.normal(0., 1.5, n)
        pc[:, 3] = rng.uniform(0., 1., n)
        pc[:, 7] = rng.integers(0, 19, n)
        R = orc.rotation_matrix_3d(1.1)
        want = orc.preprocess_pc(pc, R, 0.7, -1.3, 44., 64, 1.0) if n > 1 else None
        got = ops.preprocess_pc(pc, R, 0.7, -1.3, 44., 1.0, 44., 64).cpu().numpy()
        if want is not None:           # (a single point takes another BLAS path in the reference: SURVEY 8c)
            eq(got, want)
        grid = orc.preprocess_pc(pc, np.eye(3), 0., 0., 60., 64, None) if n > 1 else pc * 0
        sel, rest = ops.partition_semantic_pc(grid, [0, 5, 7], 7)
        ws, wr = orc.partition_semantic_pc(grid, [0, 5, 7], 7)
        eq(sel.cpu().numpy(), ws)
        eq(rest.cpu().numpy(), wr)
        alls, none = ops.partition_semantic_pc(grid, list(range(19)), 7)
        assert alls.shape[0] == grid.shape[0] and none.shape[0] == 0
        m = ops.cell_stats(grid, 64, [0], 7, weight_col=3, finish=2, want=('sel', 'wsum'))
        np.testing.assert_allclose(m['wsum'].cpu().numpy(), orc.intensity_map(grid, 64, 0), rtol=0, atol=1e-13)
        m = ops.cell_stats(grid, 64, [13, 14, 15, 17], 7, finish=1, want=('sel', 'rest'))
        eq(m['sel'].cpu().numpy(), orc.sem_probmap(grid, 64, [13, 14, 15, 17]))
    maps = rng.integers(0, 5, (4, 9, 7)).astype(np.float64)
    eq(ops.dirichlet_expectation(maps, 3.).cpu().numpy(), np.stack(orc.dirichlet_expectation(list(maps), 3.)))
    t = torch.from_numpy(maps).cuda()
    ops.dirichlet_expectation(t, 3.)
    eq(t.cpu().numpy(), maps)                                           # the caller's tensor is not modified
    assert ops.sync() == 0
    ops.close()


def test_per_step_operators_reject_bad_arguments():
    """The stand-alone entry points fail loudly (PcaccError with the library's message) instead of reading
    out of bounds: too many classes, a class column outside the row, a weight column outside the row, a
    non-positive grid."""
    from pc_accumulation_lib_b200.device import DeviceCloud
    from pc_accumulation_lib_b200._lib import PcaccError
    ops = DeviceCloud(1024, 8)
    pc = np.zeros((10, 10))
    with pytest.raises(PcaccError, match='cell_stats'):
        ops.cell_stats(pc, 8, list(range(40)), 7)
    with pytest.raises(PcaccError, match='cell_stats'):
        ops.cell_stats(pc, 8, [1], 12)
    with pytest.raises(PcaccError, match='cell_stats'):
        ops.cell_stats(pc, 8, [1], 7, weight_col=10)
    with pytest.raises(PcaccError, match='cell_stats'):
        ops.cell_stats(pc, 0, [1], 7)
    with pytest.raises(PcaccError, match='partition_semantic_pc'):
        ops.partition_semantic_pc(pc, [1], 10)
    with pytest.raises(PcaccError, match='preprocess_pc'):
        ops.preprocess_pc(pc, np.eye(3), 0., 0., 10., None, 10., 0)        # pos2grid needs P > 0
    with pytest.raises(AssertionError):
        ops.preprocess_pc(np.zeros((4, 2)))                                  # rows need x, y, z
    with pytest.raises(AssertionError):
        ops.cell_stats(pc, 8, weights=np.zeros(3))                           # one weight per point
    # the handle stays usable after an error return
    m = ops.cell_stats(pc, 8, want=('sel',))
    assert float(m['sel'].sum()) == 10.
    ops.close()
