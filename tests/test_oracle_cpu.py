"""The CPU oracle (oracle/oracle.py) against the golden vectors produced by
the literal reference (tests/golden/make_golden.py).  Everything here is
bit-exact: same host arithmetic, restated."""
import numpy as np
import pytest

from oracle import oracle as orc
from pc_accumulation_lib_b200 import synth
from tests.conftest import (assert_bev_equal, load_golden, unpack_bev,
                            unpack_sem_pcs)
from tests.golden import cases


def test_matmul_equals_fma_chain_on_this_host():
    """The premise of SURVEY.md §0: numpy's f64 matmul (N>=2) for the shapes
    on the path is a sequential FMA chain over k."""
    rng = np.random.default_rng(5)
    for n in (2, 3, 17, 4097):
        T = np.eye(4)
        T[:3, :3] += rng.normal(0, 0.3, (3, 3))
        T[:3, 3] = rng.normal(0, 5, 3)
        p = rng.normal(0, 30, (n, 3))
        ref = np.matmul(T, np.concatenate([p, np.ones((n, 1))], 1).T).T[:, :3]
        np.testing.assert_array_equal(orc.affine(T, p), ref)
        P = rng.normal(0, 100, (3, 4))
        p32 = p.astype(np.float32)
        ref = np.matmul(P, np.concatenate([p32, np.ones((n, 1))], 1).T).T
        np.testing.assert_array_equal(orc.affine(P, p32), ref)
        R = orc.rotation_matrix_3d(0.7)
        big = rng.normal(0, 30, (n, 10))
        ref = np.matmul(R, big[:, :3].T).T
        np.testing.assert_array_equal(orc.rot33(R, big[:, :3]), ref)


def test_kitti_project_golden():
    g = load_golden('kitti_project.npz')
    inp = cases.kitti_project_inputs()
    assert str(g['input_sha256']) == cases.digest(inp['pc'], inp['P'],
                                                  inp['rgb'], inp['prob'])
    u, v, mask = orc.project(inp['pc'], inp['P'], synth.KITTI_IMG_H,
                             synth.KITTI_IMG_W)
    kept = np.flatnonzero(mask)
    np.testing.assert_array_equal(kept, g['kept_idx'])
    np.testing.assert_array_equal(u[kept], g['u'])
    np.testing.assert_array_equal(v[kept], g['v'])
    sem = orc.gen_semantic_pc(inp['pc'], inp['rgb'], inp['P'])
    np.testing.assert_array_equal(sem[:, 4:], g['gather_rgb'])
    sem = orc.gen_semantic_pc(inp['pc'], inp['cls'][..., None], inp['P'])
    np.testing.assert_array_equal(sem[:, 4], g['gather_cls'])
    sem = orc.gen_semantic_pc(inp['pc'], inp['prob'], inp['P'])
    np.testing.assert_array_equal(sem[:, 4:], g['gather_prob'])
    _, _, m30 = orc.project(inp['pc'], inp['P'], synth.KITTI_IMG_H,
                            synth.KITTI_IMG_W, max_depth=30.0)
    np.testing.assert_array_equal(np.flatnonzero(m30),
                                  g['kept_idx_maxdepth30'])


def _gen_params(bev_params):
    return dict(sem_idxs=synth.SEM_IDXS, view_size=bev_params['view_size'],
                pixel_size=bev_params['pixel_size'],
                int_scaler=bev_params['int_scaler'],
                int_sep_scaler=bev_params['int_sep_scaler'],
                int_mid_threshold=bev_params['int_mid_threshold'],
                height_filter=bev_params['height_filter'], rgb_fill=0)


@pytest.mark.parametrize('name,kw,gt', [
    ('kitti_seq.npz', dict(n_frames=9), False),
    ('kitti_seq_p256.npz', dict(n_frames=6), False),
    ('kitti_gtsem_seq.npz',
     dict(n_frames=5, n_beams=8, n_azimuth=300, config=3), True),
])
def test_kitti_sequence_golden(name, kw, gt):
    g = load_golden(name)
    frames = cases.kitti_seq_inputs(use_gt_sem=gt, **kw)
    assert str(g['input_sha256']) == cases.kitti_seq_digest(frames)
    P = int(g['P'])
    acc = orc.KittiOracle(float(g['horizon']),
                          synth.kitti_calib()['p_velo_frame'],
                          synth.KITTI_FILTERS,
                          _gen_params(synth.kitti_bev_params(pixel_size=P)),
                          use_gt_sem=gt)
    evicted, n_kept = [], []
    for fr in frames:
        evicted.append(acc.integrate(fr['pc'], fr['rgb'], fr['cls'], fr['T'],
                                     fr['sem_gt']))
        n_kept.append(acc.sem_pcs[-1].shape[0])
    np.testing.assert_array_equal(evicted, g['evicted'])
    np.testing.assert_array_equal(n_kept, g['n_kept'])
    np.testing.assert_array_equal(np.array(acc.poses), g['poses'])
    np.testing.assert_array_equal(np.array(acc.seg_dists), g['seg_dists'])
    want = unpack_sem_pcs(g)
    assert len(want) == len(acc.sem_pcs)
    for a, b in zip(acc.sem_pcs, want):
        np.testing.assert_array_equal(a, b)
    bev = acc.generate_bev(int(g['present_idx']))
    assert_bev_equal(bev, unpack_bev(g))


def test_nusc_sequence_golden():
    g = load_golden('nusc_seq.npz')
    scene = cases.nusc_seq_inputs()
    assert str(g['input_sha256']) == cases.nusc_seq_digest(scene)
    P = int(g['P'])
    acc = orc.NuscOracle(synth.NUSC_FILTERS,
                         _gen_params(synth.nusc_bev_params(pixel_size=P)))
    for o in scene:
        acc.integrate(o, o['_semseg'])
    np.testing.assert_array_equal(np.array(acc.poses), g['poses'])
    np.testing.assert_array_equal(np.array(acc.seg_dists), g['seg_dists'])
    assert list(g['dyn_instances']) == acc.dyn_instances
    for a, b in zip(acc.sem_pcs, unpack_sem_pcs(g)):
        np.testing.assert_array_equal(a, b)
    bev = acc.generate_bev(int(g['present_idx']))
    assert_bev_equal(bev, unpack_bev(g))


@pytest.mark.parametrize('name,kw', [
    ('bev_direct.npz', {}),
    ('bev_direct_p128.npz', dict(n=20000, seed=78, P=128, view=51.2)),
])
def test_bev_direct_golden(name, kw):
    g = load_golden(name)
    pcs, trajs, aug, gen = cases.bev_direct_inputs(**kw)
    assert str(g['input_sha256']) == cases.digest(
        pcs['pc_present'], pcs['pc_future'], trajs['ego_traj_full'])
    bev = orc.generate(pcs, trajs, gen, return_f64=True, **aug)
    dbg = bev.pop('_debug')
    assert_bev_equal(bev, unpack_bev(g, 'bev_'))
    for w in ('present', 'future', 'full'):
        ij = g[f'grid_ij_{w}']
        st = pcs[f'pc_{w}']
        # golden holds all cropped points; oracle debug holds static ones
        pcg = orc.preprocess_pc(st, orc.rotation_matrix_3d(aug['rot_ang']),
                                aug['trans_dx'], aug['trans_dy'],
                                aug['zoom_scalar'] * gen['view_size'],
                                gen['pixel_size'], gen['height_filter'])
        np.testing.assert_array_equal(pcg[:, :2].astype(np.int32), ij)
        np.testing.assert_array_equal(pcg[:, 2], g[f'grid_z_{w}'])
    bev = orc.generate(pcs, trajs, gen)
    assert_bev_equal(bev, unpack_bev(g, 'bevhead_'))


def test_bev_warp_golden():
    """Polynomial warp (do_warp=True): the oracle, fed the same RNG streams the reference
    drew from (global numpy + `random`, seeded per case), reproduces the reference's warped
    grids and trajectories."""
    import random
    g = load_golden('bev_warp.npz')
    pcs, trajs, aug, gen = cases.bev_direct_inputs()
    assert str(g['input_sha256']) == cases.digest(
        pcs['pc_present'], pcs['pc_future'], trajs['ego_traj_full'])
    gen = dict(gen, do_warp=True)
    for s in g['seeds']:
        rngs = (np.random.RandomState(int(s)), random.Random(int(s)))
        bev = orc.generate(*cases.copy_pcs_trajs(pcs, trajs), gen, warp_rngs=rngs, **aug)
        assert_bev_equal(bev, unpack_bev(g, f'bev{int(s)}_'))


def test_input_side_matches_reference_golden():
    """box -> point assignment and multi-camera projection, pinned to
    datasets/nuscenes_utils.py run unmodified (tests/golden/make_golden.py gen_input_side)."""
    g = load_golden('input_side.npz')
    c = cases.input_side_inputs()
    for tag, pts in (('f64', c['pc']), ('f32', c['pc_f32'])):
        box, cnt = orc.assign_boxes(pts, c['boxes'], c['sizes'], c['tolerance'])
        np.testing.assert_array_equal(box, g[f'box_{tag}'])
        np.testing.assert_array_equal(cnt, g[f'cnt_{tag}'])
    assert (g['cnt_f64'] == 0).any() and (g['box_f64'] >= 0).sum() > 100
    pc_glob = orc.homo_transform(c['glob_from_ego'], c['pc'])
    for j, cam in enumerate(c['cams']):
        pc_cam = orc.homo_transform(np.linalg.inv(cam['glob_from_self']), pc_glob)
        uv, mask = orc.project_pts3d(pc_cam, cam['cam_K'], cam['img_wh'])
        np.testing.assert_array_equal(uv, g[f'uv_cam{j}'])
        np.testing.assert_array_equal(mask, g[f'mask_cam{j}'])
    uv, idx = orc.project_to_cameras(c['pc'], c['glob_from_ego'], c['cams'])
    np.testing.assert_array_equal(uv, g['pc_uv'])
    np.testing.assert_array_equal(idx, g['pc_cam_idx'])
    # the seams are seen by two cameras: the later one must have won
    seen = np.stack([g[f'mask_cam{j}'] for j in range(len(c['cams']))])
    assert (seen.sum(axis=0) > 1).sum() > 100


def test_standalone_helpers_match_reference_golden():
    """SURVEY.md 8f rank 4 remainder: pts_feat_from_img (nearest + bilinear) and
    static_obj_partitioning_by_elev against outputs of the unmodified reference."""
    from pc_accumulation_lib_b200 import synth
    g = load_golden('helpers.npz')
    c = synth.helper_inputs()
    np.testing.assert_array_equal(orc.pts_feat_from_img(c['uv'], c['img4'], 'nearest'), g['nearest4'])
    b = orc.pts_feat_from_img(c['uv'], c['img2d'], 'bilinear')
    assert np.array_equal(b, g['bilinear2d'], equal_nan=True)
    assert (~np.isfinite(b)).sum() > 0          # integral coordinates: the reference's 0/0
    pc = c['pc'].copy()
    st, dy, el, ob = orc.static_obj_partitioning_by_elev(pc, c['P'], c['elev_thresh'])
    for got, key in ((pc, 'pc_after'), (st, 'pc_static'), (dy, 'pc_dynamic'), (el, 'elevmap'), (ob, 'obs')):
        np.testing.assert_array_equal(got, g[key], err_msg=key)


def test_per_step_methods_match_reference_golden():
    """The per-step public methods the fused rasteriser replaces (crop_view ... velo2frame): the
    oracle's restatements against outputs of the unmodified reference (tests/golden/standalone.npz)."""
    from pc_accumulation_lib_b200 import synth
    g = load_golden('standalone.npz')
    c = synth.standalone_inputs()
    P, view = c['P'], c['view']
    R = orc.rotation_matrix_3d(c['rot_ang'])
    S = synth.SEM_IDXS
    eq = np.testing.assert_array_equal
    eq(orc.crop_view(c['pc'], view), g['crop_view'])
    assert g['crop_view'].shape[0] < c['pc'].shape[0] - 100
    eq(orc.geometric_transform_pc(c['pc'], R, c['dx'], c['dy'], view), g['geometric_transform'])
    eq(orc.preprocess_pc(c['pc'], R, c['dx'], c['dy'], view, P, c['height_filter']), g['preprocess_pc'])
    eq(orc.gridmap_count_map(c['edges'], P), g['count_map'])
    assert g['count_map'].sum() == c['edges'].shape[0] - 21       # 10 + 10 outside, one NaN
    np.testing.assert_allclose(orc.gridmap_count_map(c['edges'], P, c['weights']), g['count_map_weighted'],
                               rtol=0, atol=1e-12)
    eq(orc.sem_probmap(c['edges'], P, [S['road']]), g['sem_probmap_road'])
    eq(orc.sem_probmap(c['edges'], P, [S[k] for k in ('car', 'truck', 'bus', 'motorcycle')]),
       g['sem_probmap_veh'])
    np.testing.assert_allclose(orc.intensity_map(c['edges'], P, S['road']), g['intensity_map'], rtol=0,
                               atol=1e-13)
    a, b = orc.partition_semantic_pc(c['edges'], [S['car'], S['bus']], 7)
    assert np.array_equal(a, g['partition_sel'], equal_nan=True)
    assert np.array_equal(b, g['partition_rest'], equal_nan=True)
    eq(np.stack(orc.dirichlet_expectation(c['maps'], 2)), g['dirichlet'])
    eq(orc.road_marking_transform(c['vals'], 1., 30., 0.12), g['road_marking'])
    eq(orc.road_marking_transform(c['vals'], 20., 20., 0.5), g['road_marking_kitti'])
    eq(orc.sigmoid(c['vals']), g['sigmoid'])
    elev, obs = orc.elevation_map(c['grid'], P)
    eq(elev, g['elevmap'])
    eq(obs, g['elev_obs'])
    eq(np.stack(orc.rgb_maps(c['grid'], P, rgb_fill=7)), g['rgb_maps'])
    assert (g['rgb_maps'] % 1 == 0.5).sum() > 10                  # even counts: half-integer medians
    np.testing.assert_allclose(orc.velo2frame(c['pts32'][:, :3], c['P34']), g['velo2frame32'], rtol=1e-13,
                               atol=1e-12)
    np.testing.assert_allclose(orc.velo2frame(c['pc'][4:, :3], c['P34']), g['velo2frame64'], rtol=1e-13,
                               atol=1e-12)
