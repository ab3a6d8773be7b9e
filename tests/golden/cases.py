"""Input builders for the golden cases — shared by make_golden.py (which feeds
them to the literal reference), the oracle tests and the GPU parity tests, so
all three see identical bytes.  Inputs are regenerated from seeds
(pc_accumulation_lib_b200.synth); each golden file stores a sha256 of its
inputs so generator drift is reported as such and not as a parity failure."""
from __future__ import annotations

import hashlib

import numpy as np

from pc_accumulation_lib_b200 import synth


def digest(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


# --- case: one KITTI-360-shaped frame through velo2img / gen_semantic_pc ----
def kitti_project_inputs():
    seed = synth.seed_for(1, 0)
    pc = synth.kitti_lidar(seed)                         # (120000,4) f32
    # a few hand-made edge points: depth == 0, behind camera, exact .5 pixels
    calib = synth.kitti_calib()
    rgb = synth.kitti_rgb(seed)
    prob = synth.kitti_prob_map(seed)
    cls = np.argmax(prob, axis=2).astype(np.int64)
    return dict(pc=pc, P=calib['p_velo_frame'], rgb=rgb, prob=prob, cls=cls)


# --- case: KITTI-360-shaped sequence through integrate / generate_bev -------
def kitti_seq_inputs(n_frames=7, n_beams=16, n_azimuth=900, config=1,
                     use_gt_sem=False):
    frames = []
    for f in range(n_frames):
        seed = synth.seed_for(config, f)
        pc = synth.kitti_lidar(seed, n_beams, n_azimuth)
        fr = dict(pc=pc, T=synth.kitti_step_transform(seed))
        if use_gt_sem:
            fr['sem_gt'] = synth.kitti_sem_gt(seed, pc.shape[0],
                                              unfiltered_only=False)
            fr['rgb'] = np.zeros((4, 4, 3), dtype=np.uint8)   # unused
            fr['cls'] = None
        else:
            fr['rgb'] = synth.kitti_rgb(seed)
            fr['cls'] = synth.kitti_class_map(seed)
            fr['sem_gt'] = None
        frames.append(fr)
    return frames


def kitti_seq_digest(frames):
    parts = []
    for fr in frames:
        parts += [fr['pc'], fr['T'], fr['rgb']]
        parts.append(fr['cls'] if fr['cls'] is not None else fr['sem_gt'])
    return digest(*parts)


# --- case: nuScenes-shaped scene ---------------------------------------------
def nusc_seq_inputs(n_samples=8, n_azimuth=542, config=2):
    return synth.nusc_scene(synth.seed_for(config, 0), n_samples,
                            n_azimuth=n_azimuth)


def nusc_seq_digest(scene):
    parts = []
    for o in scene:
        parts += [o['pc'], o['pc_cam_idx'], o['ego_at_lidar_ts'],
                  o['images'][0], o['_semseg'][0]]
    return digest(*parts)


# --- case: BEVGenerator.generate on a hand-made cloud ------------------------
def bev_direct_inputs(n=6000, seed=77, P=32, view=40.0):
    rng = np.random.default_rng(seed)
    def cloud(m):
        pc = np.zeros((m, 10))
        pc[:, 0:2] = rng.normal(0., 14., (m, 2))
        pc[:, 2] = rng.normal(0., 2., m)
        pc[:, 3] = rng.uniform(0., 1., m)
        pc[:, 4:7] = rng.integers(0, 256, (m, 3))
        pc[:, 7] = rng.choice([0, 0, 0, 1, 2, 8, 13, 14, 15, 17, 5], m)
        pc[:, 8] = rng.integers(-1, 4, m)
        pc[:, 9] = (rng.random(m) < 0.07).astype(float)
        # concentrate some points in a few cells: even/odd/large medians
        k = m // 5
        pc[:k, 0:2] = rng.normal(0., 0.6, (k, 2)) + [3.0, -2.0]
        return pc
    present, future = cloud(n), cloud(n // 2)
    full = np.concatenate([present, future])
    ego = np.cumsum(rng.normal([1.5, 0.2, 0.], 0.1, (14, 3)), axis=0) - 10.
    other = [np.cumsum(rng.normal([0.5, 1.2, 0.], 0.1, (9, 3)), axis=0) - 4.,
             np.cumsum(rng.normal([2.5, -2.2, 0.], 0.1, (30, 3)), axis=0) - 30.]
    pcs = {'pc_present': present, 'pc_future': future, 'pc_full': full}
    trajs = {'ego_traj_present': ego[:8], 'ego_traj_future': ego[8:],
             'ego_traj_full': ego,
             'other_trajs_present': [o[:5] for o in other],
             'other_trajs_future': [o[5:] for o in other],
             'other_trajs_full': other}
    aug = dict(rot_ang=0.83, trans_dx=1.25, trans_dy=-0.75, zoom_scalar=1.1,
               do_warping=True)
    gen = dict(sem_idxs=synth.SEM_IDXS, view_size=view, pixel_size=P,
               int_scaler=20., int_sep_scaler=20., int_mid_threshold=0.5,
               height_filter=2.5, rgb_fill=0)
    return pcs, trajs, aug, gen


def copy_pcs_trajs(pcs, trajs):
    """The generator mutates its inputs in place
    (bev_generator/bev_generator.py:224-231): hand it copies."""
    p = {k: (None if v is None else v.copy()) for k, v in pcs.items()}
    t = {}
    for k, v in trajs.items():
        t[k] = [x.copy() for x in v] if isinstance(v, list) else v.copy()
    return p, t


# --- case: the dataloader's input side ----------------
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
