"""Seeded synthetic inputs of the shapes named in BASELINE.json / SURVEY.md §8(d).

The reference reads KITTI-360 / nuScenes from disk (obs_dataloaders/*, out of
scope); these generators stand in for that I/O so that the reference (when
importable), the oracle and the CUDA path all see identical bytes.  Nothing in
here computes anything on the hot path: it only fabricates observations in the
input contract of `integrate()`:

  KITTI-360   observations[0] = (rgb, pc (N,4) f32, sem_gt (N,1) int16 | None)
              (reference: kitti360_sem_pc_accum.py:61-67,
               obs_dataloaders/kitti360_obs_dataloader.py:87-106)
  nuScenes    observations[0] = dict(images, pc (N,7) f64, pc_cam_idx, ...)
              (reference: obs_dataloaders/nuscenes_obs_dataloader.py:103-220)

Seeds follow SURVEY.md §8(d): seed = 20260 + 1000*config + frame id.
"""
from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------------------
#  Constants of the two dataset shapes
# ---------------------------------------------------------------------------
KITTI_IMG_W, KITTI_IMG_H = 1408, 376
KITTI_FX = KITTI_FY = 552.554261
KITTI_CX, KITTI_CY = 682.049453, 238.769549
KITTI_N_CLASSES = 19
# run_kitti360_bev_gen.py:98-99 / run_nuscenes_bev_gen.py:125
KITTI_FILTERS = [10, 11, 12, 16, 18, 255]
NUSC_FILTERS = [10, 11, 12, 16, 18]
SEM_IDXS = {'road': 0, 'car': 13, 'truck': 14, 'bus': 15, 'motorcycle': 17}

NUSC_IMG_W, NUSC_IMG_H = 1600, 900
NUSC_K = np.array([[1266.417203046554, 0.0, 816.2670197447984],
                   [0.0, 1266.417203046554, 491.50706579294757],
                   [0.0, 0.0, 1.0]])


def seed_for(config: int, idx: int) -> int:
    return 20260 + 1000 * config + idx


def kitti_bev_params(pixel_size: int = 256, view_size=80, height_filter=None):
    # defaults of run_kitti360_bev_gen.py:25-73,128-139
    return {
        'type': 'sem', 'view_size': view_size, 'pixel_size': pixel_size,
        'max_trans_radius': 0., 'zoom_thresh': 0., 'do_warp': False,
        'int_scaler': 20., 'int_sep_scaler': 20., 'int_mid_threshold': 0.5,
        'height_filter': height_filter,
    }


def nusc_bev_params(pixel_size: int = 256):
    # defaults of run_nuscenes_bev_gen.py:71-85
    return {
        'type': 'sem', 'view_size': 51.2, 'pixel_size': pixel_size,
        'max_trans_radius': 0., 'zoom_thresh': 0., 'do_warp': False,
        'int_scaler': 1., 'int_sep_scaler': 30., 'int_mid_threshold': 0.12,
        'height_filter': 3.,
    }


# ---------------------------------------------------------------------------
#  KITTI-360-shaped
# ---------------------------------------------------------------------------
def kitti_calib() -> dict:
    """Calibration dict in the layout of run_kitti360_bev_gen.py:104-119."""
    p_cam_frame = np.array([[KITTI_FX, 0., KITTI_CX, 0.],
                            [0., KITTI_FY, KITTI_CY, 0.],
                            [0., 0., 1., 0.]])
    # velodyne (x fwd, y left, z up) -> camera (x right, y down, z fwd)
    h_velo_cam = np.array([[0.0371, -0.9993, 0.0026, 0.2628],
                           [-0.0089, -0.0029, -0.9999, -0.1121],
                           [0.9993, 0.0371, -0.0090, -0.8296],
                           [0., 0., 0., 1.]])
    p_velo_frame = np.matmul(p_cam_frame, h_velo_cam)
    return {
        'h_velo_cam': h_velo_cam, 'p_cam_frame': p_cam_frame,
        'p_velo_frame': p_velo_frame, 'c_x': KITTI_CX, 'c_y': KITTI_CY,
        'f_x': KITTI_FX, 'f_y': KITTI_FY,
    }


def kitti_lidar(seed: int, n_beams: int = 64, n_azimuth: int = 1875,
                dtype=np.float32) -> np.ndarray:
    """64-beam sweep hitting a ground plane 1.73 m below the sensor or an
    obstacle at U(8,70) m, whichever is nearer.  Returns (N,4) [x,y,z,i]."""
    rng = np.random.default_rng(seed)
    elev = np.deg2rad(np.linspace(-24.8, 2.0, n_beams))
    azim = np.linspace(-np.pi, np.pi, n_azimuth, endpoint=False)
    el, az = np.meshgrid(elev, azim, indexing='ij')
    el = el.ravel()
    az = az.ravel() + rng.normal(0., 1e-4, el.size)
    r_obst = rng.uniform(8., 70., el.size)
    with np.errstate(divide='ignore'):
        r_ground = np.where(el < -1e-3, 1.73 / np.sin(-el), np.inf)
    r = np.minimum(r_ground, r_obst)
    pc = np.empty((el.size, 4), dtype=np.float64)
    pc[:, 0] = r * np.cos(el) * np.cos(az)
    pc[:, 1] = r * np.cos(el) * np.sin(az)
    pc[:, 2] = r * np.sin(el)
    pc[:, 3] = rng.uniform(0., 1., el.size)
    return pc.astype(dtype)


def kitti_rgb(seed: int, h: int = KITTI_IMG_H, w: int = KITTI_IMG_W):
    rng = np.random.default_rng(seed + 500_000)
    return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)


def kitti_prob_map(seed: int, h: int = KITTI_IMG_H, w: int = KITTI_IMG_W,
                   k: int = KITTI_N_CLASSES) -> np.ndarray:
    """(h,w,k) f32 softmax of seeded logits with a row-dependent road prior
    (stand-in for the ONNX network's per-pixel class probabilities)."""
    rng = np.random.default_rng(seed + 700_000)
    logits = rng.normal(0., 1., (h, w, k)).astype(np.float32)
    rows = np.linspace(0., 1., h, dtype=np.float32)[:, None]
    logits[:, :, 0] += 4.0 * rows - 1.0          # road towards the bottom
    logits[:, :, 13] += 1.5 * (1.0 - rows)       # some cars higher up
    logits -= logits.max(axis=2, keepdims=True)
    e = np.exp(logits)
    return (e / e.sum(axis=2, keepdims=True)).astype(np.float32)


def kitti_class_map(seed: int, h: int = KITTI_IMG_H, w: int = KITTI_IMG_W,
                    k: int = KITTI_N_CLASSES) -> np.ndarray:
    """(h,w) int64 class-index map: argmax of `kitti_prob_map`."""
    return np.argmax(kitti_prob_map(seed, h, w, k), axis=2).astype(np.int64)


def kitti_class_map_fast(seed: int, h: int = KITTI_IMG_H,
                         w: int = KITTI_IMG_W) -> np.ndarray:
    """Cheap (h,w) uint8 class map for large benches (no softmax)."""
    rng = np.random.default_rng(seed + 700_000)
    cls = rng.integers(0, KITTI_N_CLASSES, (h, w), dtype=np.uint8)
    road = rng.random((h, w)) < np.linspace(0.05, 0.9, h)[:, None]
    cls[road] = 0
    return cls


def kitti_sem_gt(seed: int, n: int, unfiltered_only: bool = True):
    """(n,1) int16 per-point GT classes for the `use_gt_sem` path
    (kitti360_sem_pc_accum.py:138-144)."""
    rng = np.random.default_rng(seed + 900_000)
    if unfiltered_only:
        allowed = np.array([c for c in range(KITTI_N_CLASSES)
                            if c not in KITTI_FILTERS], dtype=np.int16)
    else:
        allowed = np.array(list(range(KITTI_N_CLASSES)) + [255],
                           dtype=np.int16)
    cls = allowed[rng.integers(0, allowed.size, n)]
    cls[rng.random(n) < 0.45] = 0
    return cls.reshape(n, 1).astype(np.int16)


def kitti_step_transform(seed: int) -> np.ndarray:
    """T_new_prev (4,4): maps points of the previous ego frame into the new
    one after driving forward 1-4.5 m with |yaw| <= 0.01 rad.  Stands in for
    the ICP result at kitti360_sem_pc_accum.py:123-126."""
    rng = np.random.default_rng(seed + 300_000)
    fwd = rng.uniform(1.0, 4.5)
    lat = rng.normal(0., 0.02)
    yaw = rng.uniform(-0.01, 0.01)
    pitch = rng.normal(0., 0.001)
    cy, sy = np.cos(yaw), np.sin(yaw)
    cp, sp = np.cos(pitch), np.sin(pitch)
    rz = np.array([[cy, -sy, 0.], [sy, cy, 0.], [0., 0., 1.]])
    ry = np.array([[cp, 0., sp], [0., 1., 0.], [-sp, 0., cp]])
    t_prev_new = np.eye(4)            # pose of the new frame in the old one
    t_prev_new[:3, :3] = rz @ ry
    t_prev_new[:3, 3] = [fwd, lat, rng.normal(0., 0.01)]
    return np.linalg.inv(t_prev_new)


class FakeSemseg:
    """Object with the `.pred(rgb) -> (1,1,H,W)` contract of
    utils/onnx_utils.py:32-44; returns pre-generated class maps in turn."""

    def __init__(self, class_maps):
        self._maps = list(class_maps)
        self._i = 0

    def pred(self, rgb):
        m = self._maps[self._i % len(self._maps)]
        self._i += 1
        return m[None, None]


# ---------------------------------------------------------------------------
#  nuScenes-shaped (oracle pose)
# ---------------------------------------------------------------------------
def _rot_z(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0.], [s, c, 0.], [0., 0., 1.]])


def nusc_ego_poses(seed: int, n: int) -> list:
    """n SE(3) `ego_at_lidar_ts` matrices (ego -> global), ~5 m per step."""
    rng = np.random.default_rng(seed + 100_000)
    yaw = rng.uniform(-np.pi, np.pi)
    pos = np.array([rng.uniform(300., 2000.), rng.uniform(300., 2000.), 0.])
    out = []
    for _ in range(n):
        T = np.eye(4)
        T[:3, :3] = _rot_z(yaw)
        T[:3, 3] = pos
        out.append(T)
        step = rng.uniform(4.0, 6.0)
        pos = pos + np.array([step * np.cos(yaw), step * np.sin(yaw),
                              rng.normal(0., 0.02)])
        yaw += rng.uniform(-0.03, 0.03)
    return out


def nusc_project_pts3d(pc_cam, cam_k, img_wh, depth_thres=1e-3):
    """Restatement of NuScenesCamera.project_pts3d
    (datasets/nuscenes_utils.py:112-136; nuscenes-devkit view_points with
    normalize=True is uv = (K p)[:2] / (K p)[2])."""
    valid = pc_cam[:, 2] > depth_thres
    out = np.zeros((pc_cam.shape[0], 2)) - 10.
    kp = (cam_k @ pc_cam[valid].T)
    out[valid] = (kp[:2] / kp[2:3]).T
    inside = (out > 1) & (out < np.asarray(img_wh, dtype=float) - 1)
    return out, np.all(inside, axis=1) & valid


def nusc_boxes(seed: int, n_boxes: int = 10, n_moving: int = 3):
    """Static description of the scene's annotated boxes (global frame)."""
    rng = np.random.default_rng(seed + 200_000)
    boxes = []
    clss = [0, 1, 3, 5, 0, 7, 0, 4, 0, 2, 6, 0]
    for b in range(n_boxes):
        boxes.append({
            'token': f'inst_{seed}_{b:03d}',
            'cls': clss[b % len(clss)],
            'offset': np.array([rng.uniform(8., 25.), rng.uniform(-6., 6.), 0.8]),
            'vel': (np.array([rng.uniform(1.5, 3.0), rng.uniform(-.3, .3), 0.])
                    if b < n_moving else np.zeros(3)),
            'size': np.array([4.2, 1.9, 1.6]),
        })
    return boxes


def nusc_obs(seed: int, T_ego_global: np.ndarray, ts: int, boxes=None,
             anchor_global=None, n_beams: int = 32, n_azimuth: int = 1084,
             img_h: int = NUSC_IMG_H, img_w: int = NUSC_IMG_W,
             with_images: bool = True) -> dict:
    """One observation dict in the nuScenes dataloader's output contract.

    CAM_FRONT only (the reference's commented single-camera option,
    obs_dataloaders/nuscenes_obs_dataloader.py:31).  `images` holds uint8
    (H,W,3) arrays (np.array(PIL) yields the same)."""
    rng = np.random.default_rng(seed)
    elev = np.deg2rad(np.linspace(-30.0, 10.0, n_beams))
    azim = np.linspace(-np.pi, np.pi, n_azimuth, endpoint=False)
    el, az = np.meshgrid(elev, azim, indexing='ij')
    el = el.ravel()
    az = az.ravel() + rng.normal(0., 1e-4, el.size)
    n = el.size
    r_obst = rng.uniform(4., 60., n)
    with np.errstate(divide='ignore'):
        r_ground = np.where(el < -1e-3, 1.84 / np.sin(-el), np.inf)
    r = np.minimum(r_ground, r_obst)
    xyz = np.empty((n, 3))
    xyz[:, 0] = r * np.cos(el) * np.cos(az)
    xyz[:, 1] = r * np.cos(el) * np.sin(az)
    xyz[:, 2] = r * np.sin(el) + 1.84
    intensity = rng.integers(0, 256, n).astype(np.float64)

    # camera: ego (x fwd, y left, z up) -> cam (x right, y down, z fwd)
    r_cam_ego = np.array([[0., -1., 0.], [0., 0., -1.], [1., 0., 0.]])
    t_cam = np.array([1.70, 0.016, 1.51])
    pc_cam = (r_cam_ego @ (xyz - t_cam).T).T
    uv, in_img = nusc_project_pts3d(pc_cam, NUSC_K, (img_w, img_h))
    pc_uv = np.zeros((n, 2))
    pc_uv[in_img] = uv[in_img]
    pc_cam_idx = -np.ones(n, dtype=int)
    pc_cam_idx[in_img] = 0

    # boxes: assign instance index to points inside axis-aligned boxes
    inst = -np.ones(n)
    inst_tokens, inst_cls, inst_center = [], [], []
    if boxes:
        T_global_ego = np.linalg.inv(T_ego_global)
        for b in boxes:
            center_g = anchor_global + b['offset'] + b['vel'] * ts
            c_e = (T_global_ego @ np.append(center_g, 1.))[:3]
            half = 0.5 * b['size']
            inside = np.all(np.abs(xyz - c_e) < half, axis=1)
            idx = len(inst_tokens)
            inst[inside] = idx
            inst_tokens.append(b['token'])
            inst_cls.append(int(b['cls']))
            inst_center.append(center_g.copy())

    obs = {
        'pc': np.concatenate([xyz, intensity[:, None], pc_uv, inst[:, None]],
                             axis=1),
        'pc_cam_idx': pc_cam_idx,
        'ego_at_lidar_ts': T_ego_global,
        'ego_global_x': float(T_ego_global[0, 3]),
        'ego_global_y': float(T_ego_global[1, 3]),
        'inst_tokens': inst_tokens,
        'inst_cls': inst_cls,
        'inst_center': inst_center,
        'meta': {'cam_channels': ['CAM_FRONT']},
    }
    if with_images:
        obs['images'] = [rng.integers(0, 256, (img_h, img_w, 3),
                                      dtype=np.uint8)]
        cls = rng.integers(0, KITTI_N_CLASSES, (img_h, img_w)).astype(np.int64)
        road = rng.random((img_h, img_w)) < np.linspace(0.05, 0.9, img_h)[:, None]
        cls[road] = 0
        obs['_semseg'] = [cls]     # consumed by the stand-in semseg model
    return obs


def nusc_scene(seed: int, n_samples: int = 40, **kw) -> list:
    """List of observation dicts for one scene."""
    poses = nusc_ego_poses(seed, n_samples)
    boxes = nusc_boxes(seed)
    anchor = poses[0][:3, 3].copy()
    return [nusc_obs(seed + 1 + ts, poses[ts], ts, boxes, anchor, **kw)
            for ts in range(n_samples)]


class SceneSemseg:
    """Stand-in semseg for nuScenes scenes: `.pred(rgb)` returns the class map
    that `nusc_obs` generated for the image with the same id()."""

    def __init__(self):
        self._by_id = {}

    def register(self, obs: dict):
        for img, cls in zip(obs['images'], obs['_semseg']):
            self._by_id[id(img)] = cls

    def pred(self, rgb):
        return self._by_id[id(rgb)][None, None]


# ---------------------------------------------------------------------------
# the dataloader's input side: multi-camera projection, boxes
# ---------------------------------------------------------------------------
def _rigid(rng, yaw_range, t_scale, tilt=0.02):
    yaw, pitch, roll = rng.uniform(-yaw_range, yaw_range), rng.normal(0, tilt), rng.normal(0, tilt)
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    R = (np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]]) @ np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
         @ np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]]))
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = rng.normal(0, t_scale, 3)
    return T


def input_side_inputs(n=12000, n_boxes=24, n_cams=6, seed=91):
    """Lidar points in the ego frame, six cameras looking around (CAM_* of nuScenes: 1600x900,
    fx = fy ~ 1266, 60 degrees apart, overlapping at the seams so that 'a later camera
    overwrites an earlier one' is exercised), and boxes in the target frame, some of them
    overlapping and some empty."""
    rng = np.random.default_rng(seed)
    r = rng.uniform(2., 60., n)
    az = rng.uniform(-np.pi, np.pi, n)
    pc = np.stack([r * np.cos(az), r * np.sin(az), rng.normal(-0.5, 1.2, n)], axis=1)
    pc[:50] = 0.0                                  # degenerate points at the sensor origin
    glob_from_ego = _rigid(rng, np.pi, 300., 0.01)
    cams = []
    for j in range(n_cams):
        yaw = 2 * np.pi * j / n_cams + rng.normal(0, 0.01)
        # camera axes: z forward, x right, y down
        fwd = np.array([np.cos(yaw), np.sin(yaw), 0.])
        right = np.array([np.sin(yaw), -np.cos(yaw), 0.])
        down = np.array([0., 0., -1.])
        ego_from_cam = np.eye(4)
        ego_from_cam[:3, :3] = np.stack([right, down, fwd], axis=1)
        ego_from_cam[:3, 3] = [1.5 * np.cos(yaw), 0.5 * np.sin(yaw), 1.5]
        fx = 1266.4 * (0.55 if j % 2 else 1.0)     # every other camera wide: neighbours overlap
        cams.append(dict(glob_from_self=glob_from_ego @ ego_from_cam,
                         cam_K=np.array([[fx, 0., 816.3 + j], [0., fx, 491.5 - j], [0., 0., 1.]]),
                         img_wh=np.array([1600., 900.])))
    boxes, sizes = [], []
    for b in range(n_boxes):
        T = _rigid(rng, np.pi, 0.0, 0.0)
        T[:3, 3] = [rng.uniform(-40, 40), rng.uniform(-40, 40), rng.normal(-0.3, 0.3)]
        if b % 5 == 1:                             # overlaps the previous box
            T[:3, 3] = boxes[-1][:3, 3] + [0.8, 0.3, 0.0]
        if b % 7 == 3:                             # far away: no points
            T[:3, 3] = [500. + 30. * b, 500., 0.]
        boxes.append(T)
        sizes.append(np.array([rng.uniform(3.5, 12.), rng.uniform(1.6, 3.), rng.uniform(1.4, 3.5)]))
    # points exactly on box faces / at box centres
    pc[50:50 + n_boxes] = [T[:3, 3] if b % 7 != 3 else [1., 1., 1.] for b, T in enumerate(boxes)]
    return dict(pc=pc, pc_f32=pc.astype(np.float32), glob_from_ego=glob_from_ego, cams=cams,
                boxes=boxes, sizes=sizes, tolerance=1e-2)


def helper_inputs(seed=123, n=4000, P=48):
    """Inputs of the two stand-alone helpers of SURVEY.md 8f rank 4: pixel coordinates strictly
    inside a 97 x 61 image (some with integral u or v: the bilinear 0/0), a float64 single-channel
    image, an int64 4-channel image, and a cloud in grid coordinates with stacked points per cell."""
    rng = np.random.default_rng(seed)
    H, W = 61, 97
    uv = np.stack([rng.uniform(1.001, W - 1.001, n), rng.uniform(1.001, H - 1.001, n)], axis=1)
    uv[:40, 0] = np.floor(uv[:40, 0]) + 0.5            # exact halves: np.round to even
    uv[40:60, 1] = np.floor(uv[40:60, 1])              # integral v: total == 0 in the bilinear branch
    uv[60:70] = np.floor(uv[60:70]) + [0.0, 0.25]      # integral u
    uv = np.clip(uv, 1.001, None)
    img2d = rng.normal(0., 50., (H, W))
    img4 = rng.integers(0, 256, (H, W, 4)).astype(np.int64)
    pc = np.zeros((n, 10))
    pc[:, 0:2] = np.floor(rng.uniform(0, P, (n, 2)))
    pc[:200, 0:2] = [5., 7.]                            # a crowded cell
    pc[200:230, 0] += 0.75                              # fractional grid coordinates truncate
    pc[:, 2] = rng.normal(0., 1.5, n)
    pc[:, 3] = rng.uniform(0., 1., n)
    pc[:, 7] = rng.integers(0, 19, n)
    pc[::17, 8] = 1.                                    # already flagged
    pc[5::29, 8] = 3.                                   # neither 0 nor 1: dropped from both outputs
    return dict(uv=uv, img2d=img2d, img4=img4, pc=pc, P=P, elev_thresh=0.8)


def rgb_bev_inputs(seed=31, P=32, n=3000):
    """Pre-processed clouds (grid coordinates, r g b in columns 4..6) and pose lists for
    RGBBEVGenerator.generate_bev."""
    rng = np.random.default_rng(seed)

    def cloud(m):
        pc = np.zeros((m, 7))
        pc[:, 0:2] = rng.integers(0, P, (m, 2))
        pc[:, 2] = rng.normal(0, 1, m)
        pc[:, 4:7] = rng.integers(0, 256, (m, 3))
        return pc
    poses_p = np.stack([np.linspace(3, P - 4, 9), np.linspace(5, P - 6, 9), np.zeros(9)], axis=1).round()
    poses_f = np.stack([np.linspace(P - 4, 2, 7), np.linspace(4, P - 3, 7), np.zeros(7)], axis=1).round()
    return dict(P=P, pc_present=cloud(n), pc_future=cloud(n // 2), poses_present=poses_p, poses_future=poses_f)


def standalone_inputs(seed=77, n=6000, P=48, view=37.2):
    """Inputs of the stand-alone per-step methods (crop_view, geometric_transform,
    preprocess_pc_and_trajs, gen_gridmap_count_map, gen_sem_probmap, gen_intensity_map,
    partition_semantic_pc, dirichlet_dist_expectation, road_marking_transform, get_elevation_map,
    get_rgb_maps, velo2frame): a metric (n,10) cloud with points on the crop boundary and a NaN row,
    and a cloud in grid coordinates with the histogram's edge cases (a coordinate equal to P,
    fractional and negative coordinates, NaN)."""
    rng = np.random.default_rng(seed)
    pc = np.zeros((n, 10))
    pc[:, 0:2] = rng.uniform(-30., 30., (n, 2))
    pc[:, 2] = rng.normal(0., 1.2, n)
    pc[:, 3] = rng.uniform(0., 1., n)
    pc[:, 4:7] = rng.integers(0, 256, (n, 3))
    pc[:, 7] = rng.integers(0, 19, n)
    pc[:, 8] = rng.integers(-1, 40, n)
    pc[:, 9] = rng.integers(0, 2, n)
    pc[0, 0:2] = [0.5 * view, 0.]            # on the boundary: strict comparisons drop it
    pc[1, 0:2] = [0., -0.5 * view]
    pc[2, 0] = np.nan
    pc[3, 2] = np.inf
    grid = np.zeros((n, 10))
    grid[:, 0:2] = np.floor(rng.uniform(0, P, (n, 2)))
    grid[:300, 0:2] = [11., 30.]             # a crowded cell
    grid[:, 2] = rng.normal(0., 1.5, n)
    grid[:, 3] = rng.uniform(0., 1., n)
    grid[:, 4:7] = rng.integers(0, 256, (n, 3))
    grid[:, 7] = rng.integers(0, 19, n)
    edges = grid.copy()
    edges[300:320, 0] = P                    # right edge: histogram2d's last bin
    edges[320:330, 1] = P
    edges[330:340, 0] = -1.                  # outside
    edges[340:350, 1] = P + 1.
    edges[350:380, 0:2] += 0.6               # fractional
    edges[380, 0] = np.nan
    vals = rng.normal(0.3, 0.4, (P, P))
    maps = [rng.integers(0, 9, (P, P)).astype(np.float64) for _ in range(3)]
    P34 = np.concatenate([rng.normal(0., 1., (3, 3)), rng.normal(0., 5., (3, 1))], axis=1)
    pts32 = rng.normal(0., 20., (n, 4)).astype(np.float32)
    return dict(pc=pc, grid=grid, edges=edges, P=P, view=view, rot_ang=0.7, dx=1.5, dy=-2.25,
                height_filter=1.1, vals=vals, maps=maps, P34=P34, pts32=pts32,
                weights=rng.uniform(-1., 2., n))
