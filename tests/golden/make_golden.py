"""Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_loader.py) on the seeded inputs
of tests/golden/cases.py.  Run in the build container only:

    python tests/golden/make_golden.py

The reference ships no golden vectors of its own, so these
outputs of the reference itself are what pins the oracle and the CUDA path.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader                      # noqa: E402
from pc_accumulation_lib_b200 import synth         # noqa: E402
from tests.golden import cases                     # noqa: E402

WINDOWS = ('present', 'future', 'full')


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def pack_bev(bev: dict, prefix: str, out: dict):
    for k, v in bev.items():
        if k.startswith('trajs_') or k == 'gt_lanes':
            out[f'{prefix}{k}_n'] = np.int64(len(v))
            for i, t in enumerate(v):
                out[f'{prefix}{k}_{i}'] = np.asarray(t, dtype=np.float64)
        else:
            out[f'{prefix}{k}'] = v


def pack_sem_pcs(sem_pcs, prefix, out):
    out[f'{prefix}n_frames'] = np.int64(len(sem_pcs))
    for i, s in enumerate(sem_pcs):
        out[f'{prefix}xyzi_{i}'] = s[:, :4].copy()
        out[f'{prefix}attr_{i}'] = s[:, 4:].astype(np.int32)
        assert np.array_equal(out[f'{prefix}attr_{i}'], s[:, 4:])


def save(name, out):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **out)
    print(f'{name}: {os.path.getsize(path) / 1024:.0f} KiB, {len(out)} arrays')


def gen_kitti_project(ref):
    inp = cases.kitti_project_inputs()
    acc = ref_loader.make_kitti_accum(
        ref, 1e9, synth.kitti_calib(), synth.KITTI_FILTERS, synth.SEM_IDXS,
        synth.kitti_bev_params(pixel_size=32))
    pc, P = inp['pc'], inp['P']
    n = pc.shape[0]
    pc5 = np.concatenate([pc, np.arange(n, dtype=np.float32)[:, None]], axis=1)
    img = acc.velo2img(pc5, P, synth.KITTI_IMG_H, synth.KITTI_IMG_W)
    out = {
        'input_sha256': np.array(cases.digest(pc, P, inp['rgb'], inp['prob'])),
        'kept_idx': img[:, 4].astype(np.int32),
        'u': img[:, 5].astype(np.int32), 'v': img[:, 6].astype(np.int32),
    }
    assert np.array_equal(img[:, :4], pc[out['kept_idx']].astype(np.float64))
    sem_rgb = acc.gen_semantic_pc(pc, inp['rgb'], P)
    sem_cls = acc.gen_semantic_pc(pc, inp['cls'][..., None], P)
    sem_prob = acc.gen_semantic_pc(pc, inp['prob'], P)
    assert np.array_equal(sem_rgb[:, :4], img[:, :4])
    out['gather_rgb'] = sem_rgb[:, 4:].astype(np.uint8)
    out['gather_cls'] = sem_cls[:, 4].astype(np.int32)
    out['gather_prob'] = sem_prob[:, 4:].astype(np.float32)
    assert np.array_equal(out['gather_prob'].astype(np.float64),
                          sem_prob[:, 4:])
    # max_depth variant
    img_d = acc.velo2img(pc5, P, synth.KITTI_IMG_H, synth.KITTI_IMG_W,
                         max_depth=30.0)
    out['kept_idx_maxdepth30'] = img_d[:, 4].astype(np.int32)
    save('kitti_project.npz', out)


def gen_kitti_seq(ref, name, use_gt_sem, P, horizon, present_idx, **kw):
    frames = cases.kitti_seq_inputs(use_gt_sem=use_gt_sem, **kw)
    bev_params = synth.kitti_bev_params(pixel_size=P)
    sem = None if use_gt_sem else synth.FakeSemseg(
        [f['cls'] for f in frames])
    acc = ref_loader.make_kitti_accum(
        ref, horizon, synth.kitti_calib(), synth.KITTI_FILTERS,
        synth.SEM_IDXS, bev_params, sem, use_gt_sem=use_gt_sem)
    evicted = []
    n_kept = []
    for fr in frames:
        ref_loader.ICP_QUEUE.append(fr['T'])
        with quiet():
            evicted.append(acc.integrate(
                [(fr['rgb'], fr['pc'], fr['sem_gt'])]))
        n_kept.append(acc.sem_pcs[-1].shape[0])
    out = {'input_sha256': np.array(cases.kitti_seq_digest(frames)),
           'evicted': np.array(evicted, dtype=np.int64),
           'n_kept': np.array(n_kept, dtype=np.int64),
           'poses': np.array(acc.poses, dtype=np.float64),
           'seg_dists': np.array(acc.seg_dists, dtype=np.float64),
           'present_idx': np.int64(present_idx), 'P': np.int64(P),
           'horizon': np.float64(horizon)}
    pack_sem_pcs(acc.sem_pcs, 'sem_pcs_', out)
    t = time.time()
    with quiet():
        bev = acc.generate_bev(present_idx, 1, True)[0]
    print(f'  reference generate_bev: {time.time() - t:.2f} s')
    pack_bev(bev, 'bev_', out)
    save(name, out)


def gen_nusc_seq(ref, name, P, present_idx, **kw):
    scene = cases.nusc_seq_inputs(**kw)
    semseg = synth.SceneSemseg()
    for o in scene:
        semseg.register(o)
    acc = ref_loader.make_nusc_accum(ref, synth.NUSC_FILTERS, synth.SEM_IDXS,
                                     synth.nusc_bev_params(pixel_size=P),
                                     semseg)
    for o in scene:
        with quiet():
            acc.integrate([o])
    out = {'input_sha256': np.array(cases.nusc_seq_digest(scene)),
           'poses': np.array(acc.poses, dtype=np.float64),
           'seg_dists': np.array(acc.seg_dists, dtype=np.float64),
           'dyn_instances': np.array(acc.dyn_instances),
           'present_idx': np.int64(present_idx), 'P': np.int64(P)}
    pack_sem_pcs(acc.sem_pcs, 'sem_pcs_', out)
    n_dyn = sum(int((s[:, 9] == 1).sum()) for s in acc.sem_pcs)
    print(f'  dynamic tokens {acc.dyn_instances}, dyn points {n_dyn}')
    with quiet():
        bev = acc.generate_bev(present_idx, 1, True)[0]
    pack_bev(bev, 'bev_', out)
    save(name, out)


def gen_bev_direct(ref, name, **kw):
    pcs, trajs, aug, gen = cases.bev_direct_inputs(**kw)
    g = ref.SemBEVGenerator(gen['sem_idxs'], gen['view_size'],
                            gen['pixel_size'], 0., 0., False,
                            gen['int_scaler'], gen['int_sep_scaler'],
                            gen['int_mid_threshold'], gen['height_filter'],
                            gen['rgb_fill'])
    out = {'input_sha256': np.array(cases.digest(
        pcs['pc_present'], pcs['pc_future'], trajs['ego_traj_full']))}
    p, t = cases.copy_pcs_trajs(pcs, trajs)
    with quiet():
        bev = g.generate(p, t, **aug)
    pack_bev(bev, 'bev_', out)
    # heading-aligned variant (do_warping False -> angle from the ego track)
    p, t = cases.copy_pcs_trajs(pcs, trajs)
    with quiet():
        bev = g.generate(p, t)
    pack_bev(bev, 'bevhead_', out)
    # preprocessed clouds (cell indices) of the explicit-angle variant
    p, t = cases.copy_pcs_trajs(pcs, trajs)
    view = aug['zoom_scalar'] * gen['view_size']
    for w in WINDOWS:
        with quiet():
            pcg, _ = g.preprocess_pc_and_trajs(
                p[f'pc_{w}'], [], aug['rot_ang'], aug['trans_dx'],
                aug['trans_dy'], view)
        out[f'grid_ij_{w}'] = pcg[:, :2].astype(np.int32)
        out[f'grid_z_{w}'] = pcg[:, 2].copy()
    save(name, out)


def gen_bev_warp(ref, name, seeds=(11, 12, 13), **kw):
    """do_warp=True: the reference draws the warp from the GLOBAL numpy and `random` RNGs
    (bev_generator.py:621-661); both are seeded right before each call."""
    import random
    pcs, trajs, aug, gen = cases.bev_direct_inputs(**kw)
    g = ref.SemBEVGenerator(gen['sem_idxs'], gen['view_size'], gen['pixel_size'], 0., 0., True,
                            gen['int_scaler'], gen['int_sep_scaler'], gen['int_mid_threshold'],
                            gen['height_filter'], gen['rgb_fill'])
    out = {'input_sha256': np.array(cases.digest(
        pcs['pc_present'], pcs['pc_future'], trajs['ego_traj_full'])),
        'seeds': np.array(seeds, dtype=np.int64)}
    for s in seeds:
        p, t = cases.copy_pcs_trajs(pcs, trajs)
        np.random.seed(s)
        random.seed(s)
        with quiet():
            bev = g.generate(p, t, **aug)
        pack_bev(bev, f'bev{s}_', out)
    save(name, out)


def gen_input_side(ref, name):
    """datasets/nuscenes_utils.py run unmodified: find_points_in_box / apply_tf (box loop of
    inst_centric_get_sweeps restated around them), homo_transform and
    NuScenesCamera.project_pts3d (called unbound on a stand-in `self`; nuscenes-devkit's
    view_points is absent here and is restated from its published source)."""
    import types
    nu = ref.nusc_utils
    c = cases.input_side_inputs()

    def view_points(points, view, normalize):          # nuscenes-devkit geometry_utils
        assert view.shape[0] <= 4 and view.shape[1] <= 4 and points.shape[0] == 3
        viewpad = np.eye(4)
        viewpad[:view.shape[0], :view.shape[1]] = view
        nbr_points = points.shape[1]
        points = np.concatenate((points, np.ones((1, nbr_points))))
        points = np.dot(viewpad, points)
        points = points[:3, :]
        if normalize:
            points = points / points[2:3, :].repeat(3, 0).reshape(3, nbr_points)
        return points

    nu.view_points = view_points
    out = {}
    # box loop (:412-470): in order, later boxes overwrite
    for tag, pts in (('f64', c['pc']), ('f32', c['pc_f32'])):
        box = -np.ones(pts.shape[0], dtype=np.int32)
        cnt = np.zeros(len(c['boxes']), dtype=np.int32)
        for b, (T, size) in enumerate(zip(c['boxes'], c['sizes'])):
            m = nu.find_points_in_box(pts, T, size, c['tolerance'])
            cnt[b] = m.sum()
            box[m] = b
        out[f'box_{tag}'], out[f'cnt_{tag}'] = box, cnt
    # multi-camera projection (nuscenes_obs_dataloader.py:176-198)
    pc_in_glob = nu.homo_transform(c['glob_from_ego'], c['pc'])
    pc_uv = np.zeros((c['pc'].shape[0], 2), dtype=float)
    pc_cam_idx = -np.ones(c['pc'].shape[0], dtype=int)
    for j, cam in enumerate(c['cams']):
        pc_in_cam = nu.homo_transform(np.linalg.inv(cam['glob_from_self']), pc_in_glob)
        fake = types.SimpleNamespace(cam_K=cam['cam_K'], img_wh=cam['img_wh'])
        uv, mask = nu.NuScenesCamera.project_pts3d(fake, pc_in_cam)
        pc_uv[mask] = uv[mask]
        pc_cam_idx[mask] = j
        out[f'uv_cam{j}'], out[f'mask_cam{j}'] = uv, mask
    out['pc_uv'], out['pc_cam_idx'] = pc_uv, pc_cam_idx.astype(np.int64)
    out['input_digest'] = np.frombuffer(cases.digest(c['pc'], c['glob_from_ego'],
                                                     *[k['glob_from_self'] for k in c['cams']],
                                                     *c['boxes'], *c['sizes']).encode(), dtype=np.uint8)
    save(name, out)


def gen_helpers(ref, name):
    """The two stand-alone helpers of SURVEY.md 8f rank 4, run unmodified: pts_feat_from_img (nearest on
    a 4-channel image; bilinear on a 2-D image, the only shape the reference's broadcasting handles) and
    SemBEVGenerator.static_obj_partitioning_by_elev."""
    c = synth.helper_inputs()
    nu = ref.nusc_utils
    out = {'nearest4': nu.pts_feat_from_img(c['uv'], c['img4'], 'nearest')}
    with np.errstate(divide='ignore', invalid='ignore'):
        out['bilinear2d'] = nu.pts_feat_from_img(c['uv'], c['img2d'], 'bilinear')
    g = ref.SemBEVGenerator(synth.SEM_IDXS, 40., c['P'])
    pc = c['pc'].copy()
    with quiet():
        st, dy, elev, obs = g.static_obj_partitioning_by_elev(pc, c['elev_thresh'])
    out.update(pc_after=pc, pc_static=st, pc_dynamic=dy, elevmap=elev, obs=obs)
    out['input_digest'] = np.frombuffer(cases.digest(c['uv'], c['img2d'], c['img4'], c['pc']).encode(),
                                        dtype=np.uint8)
    save(name, out)


def gen_rgb_bev(ref, name, seeds=(5, 6)):
    """bev_generator/rgb_bev.py run unmodified: generate_bev without and with the polynomial warp
    (global numpy / `random` RNGs seeded right before each call)."""
    import random
    from bev_generator import rgb_bev as ref_rgb        # importable after ref_loader.load()
    c = synth.rgb_bev_inputs()
    g = ref_rgb.RGBBEVGenerator(40., c['P'])
    out = {'seeds': np.array(seeds, dtype=np.int64)}
    with quiet():
        bev = g.generate_bev(c['pc_present'].copy(), c['pc_future'].copy(), c['poses_present'].copy(),
                             c['poses_future'].copy(), do_warping=False)
    for k, v in bev.items():
        out[f'plain_{k}'] = np.asarray(v)
    for s in seeds:
        np.random.seed(s)
        random.seed(s)
        with quiet():
            bev = g.generate_bev(c['pc_present'].copy(), c['pc_future'].copy(), c['poses_present'].copy(),
                                 c['poses_future'].copy(), do_warping=True)
        for k, v in bev.items():
            out[f'warp{s}_{k}'] = np.asarray(v)
    save(name, out)


def gen_standalone(ref, name):
    """The per-step public methods of BEVGenerator / SemBEVGenerator / SemanticPointCloudAccumulator that
    the fused rasteriser replaces, each run unmodified on synth.standalone_inputs()."""
    c = synth.standalone_inputs()
    P, view = c['P'], c['view']
    g = ref.SemBEVGenerator(synth.SEM_IDXS, 40., P, height_filter=c['height_filter'], rgb_fill=7)
    out = {}
    with quiet():
        out['crop_view'] = g.crop_view(c['pc'].copy(), view)
        out['geometric_transform'] = g.geometric_transform(c['pc'].copy(), c['rot_ang'], c['dx'], c['dy'], view)
        pc_g, _ = g.preprocess_pc_and_trajs(c['pc'].copy(), [], c['rot_ang'], c['dx'], c['dy'], view)
        out['preprocess_pc'] = pc_g
        out['count_map'] = g.gen_gridmap_count_map(c['edges'])
        out['count_map_weighted'] = g.gen_gridmap_count_map(c['edges'], weights=c['weights'])
        out['sem_probmap_road'] = g.gen_sem_probmap(c['edges'], ['road'])
        out['sem_probmap_veh'] = g.gen_sem_probmap(c['edges'], ['car', 'truck', 'bus', 'motorcycle'])
        out['intensity_map'] = g.gen_intensity_map(c['edges'], 'road')
        a, b = g.partition_semantic_pc(c['edges'], [synth.SEM_IDXS['car'], synth.SEM_IDXS['bus']], 7)
        out['partition_sel'], out['partition_rest'] = a, b
        post = g.dirichlet_dist_expectation([m.copy() for m in c['maps']], obs_weight=2)
        out['dirichlet'] = np.stack(post)
        out['road_marking'] = g.road_marking_transform(c['vals'].copy(), 1., 30., 0.12)
        out['road_marking_kitti'] = g.road_marking_transform(c['vals'].copy(), 20., 20., 0.5)
        out['sigmoid'] = g.sigmoid(c['vals'])
        elev, obs = g.get_elevation_map(c['grid'])
        out['elevmap'], out['elev_obs'] = elev, obs
        r, gg, b_ = g.get_rgb_maps(c['grid'])
        out['rgb_maps'] = np.stack([r, gg, b_])
        vf = ref.KittiAccum.velo2frame
        out['velo2frame32'] = vf(c['pts32'][:, :3], c['P34'])
        out['velo2frame64'] = vf(c['pc'][4:, :3], c['P34'])
    out['input_digest'] = np.frombuffer(cases.digest(c['pc'], c['grid'], c['edges'], c['vals'], c['P34'],
                                                     c['pts32'], c['weights'], *c['maps']).encode(),
                                        dtype=np.uint8)
    save(name, out)


def main():
    ref = ref_loader.load()
    if len(sys.argv) > 1 and sys.argv[1] == 'standalone':
        gen_standalone(ref, 'standalone.npz')
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'rgb_bev':
        gen_rgb_bev(ref, 'rgb_bev.npz')
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'helpers':
        gen_helpers(ref, 'helpers.npz')
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'input_side':
        gen_input_side(ref, 'input_side.npz')
        return
    gen_kitti_project(ref)
    gen_kitti_seq(ref, 'kitti_seq.npz', use_gt_sem=False, P=64, horizon=14.0,
                  present_idx=3, n_frames=9)
    gen_kitti_seq(ref, 'kitti_seq_p256.npz', use_gt_sem=False, P=256,
                  horizon=1e9, present_idx=3, n_frames=6)
    gen_kitti_seq(ref, 'kitti_gtsem_seq.npz', use_gt_sem=True, P=64,
                  horizon=1e9, present_idx=2, n_frames=5, n_beams=8,
                  n_azimuth=300, config=3)
    gen_nusc_seq(ref, 'nusc_seq.npz', P=64, present_idx=4)
    gen_bev_direct(ref, 'bev_direct.npz')
    gen_bev_direct(ref, 'bev_direct_p128.npz', n=20000, seed=78, P=128,
                   view=51.2)
    gen_bev_warp(ref, 'bev_warp.npz')
    gen_input_side(ref, 'input_side.npz')
    gen_helpers(ref, 'helpers.npz')
    gen_rgb_bev(ref, 'rgb_bev.npz')


if __name__ == '__main__':
    main()
