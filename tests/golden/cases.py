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
# the generator lives in synth.py: bench.py times the same inputs and must not import tests/
input_side_inputs = synth.input_side_inputs
