"""NuScenesOracleSemanticPointCloudAccumulator — drop-in for the reference's
`nuscenes_oracle_sem_pc_accum.py` (GT ego pose, world-frame accumulation,
nearest-pixel gather from N cameras, fake detector / tracker that flags dynamic
points retroactively, dynamic-object trajectories).

Per-point work (gather, filter, ego->world transform, dyn flags) runs in
libpcacc; the tracker's per-box bookkeeping (a handful of scalars per sweep,
nuscenes_oracle_sem_pc_accum.py:191-250,272-414) stays on the host and drives
`pcacc_mark_dynamic`.
"""
from __future__ import annotations

import numpy as np

from .sem_pc_accum import SemanticPointCloudAccumulator


def homo_transform(tf, points):
    """datasets/nuscenes_utils.py:46-60 for the handful of box centres the
    tracker moves (host side; clouds are transformed on the device)."""
    assert tf.shape == (4, 4), f'{tf.shape} is not (4, 4)'
    assert points.shape == (points.shape[0], 3), f'{points.shape} is not (N, 3)'
    p = np.concatenate([points, np.ones((points.shape[0], 1))], axis=1)
    return (tf @ p.T)[:3, :].T


class NuScenesOracleSemanticPointCloudAccumulator(SemanticPointCloudAccumulator):
    def __init__(self, semseg_onnx_path=None, semseg_filters: list = None, sem_idxs: dict = None,
                 use_gt_sem: bool = None, bev_params: dict = None, loc: str = None,
                 get_gt_lanes: bool = False, dataroot: str = None, **ring_kwargs):
        if use_gt_sem:
            raise NotImplementedError()
        super().__init__(None, None, semseg_onnx_path, semseg_filters, sem_idxs, False,
                         bev_params, **ring_kwargs)
        self.use_gt_sem = use_gt_sem
        self.ts = 0
        self.xyz_idx, self.int_idx, self.rgb_idx = 0, 3, 4
        self.sem_idx, self.inst_idx, self.dyn_idx = 7, 8, 9
        self.T_global_world = None
        self.ego_pose_z = 1.
        self.instances = {}
        self.dyn_instances = []
        self.dyn_obj_trans_thresh = 1.0
        self.token2idx = []
        self.track_inst_clss = [0, 1, 2, 3, 5]
        self.map = loc
        self.ego_global_xs = []
        self.ego_global_ys = []
        self.get_gt_lanes = get_gt_lanes
        self.gt_lane_poses = None
        if self.get_gt_lanes:
            raise NotImplementedError(
                'GT lane centrelines need the nuscenes-devkit map API (out of scope); assign '
                'accumulator.gt_lane_poses (list of (n,3) arrays, global frame) yourself')

    def integrate(self, observations: list):
        obs = observations[0]
        rgbs, pc, pc_cam_idx = obs['images'], obs['pc'], obs['pc_cam_idx']
        T_ego_global = obs['ego_at_lidar_ts']
        if self.T_global_world is None:
            self.T_global_world = np.linalg.inv(T_ego_global)
            if self.gt_lane_poses is not None:
                self.gt_lane_poses = [homo_transform(self.T_global_world, lane)
                                      for lane in self.gt_lane_poses]

        fid, pose, semsegs = self.obs2sem_vec_space(rgbs, pc, pc_cam_idx, T_ego_global,
                                                    self.ego_pose_z)
        self._fids.append(fid)
        self.poses.append(pose)
        self.rgbs.append(rgbs)
        self.semsegs.append(semsegs)
        self.ego_global_xs.append(obs['ego_global_x'])
        self.ego_global_ys.append(obs['ego_global_y'])

        # fake detector / tracker: host bookkeeping, device flag updates
        self.token2idx.append({'ts': self.ts})
        mark_f, mark_i = [], []
        for idx, token in enumerate(obs['inst_tokens']):
            if obs['inst_cls'][idx] not in self.track_inst_clss:
                continue
            centre = homo_transform(self.T_global_world,
                                    np.expand_dims(obs['inst_center'][idx], 0))[0]
            self.instances.setdefault(token, []).append((centre, self.ts))
            self.token2idx[-1][token] = idx
            if token in self.dyn_instances:          # known mover: flag the new sweep
                mark_f.append(fid)
                mark_i.append(idx)
                continue
            seen = self.instances[token]
            if len(seen) < 2:
                continue
            moved = self.cal_pose_change(seen[0][0][:2], seen[-1][0][:2])
            if moved > self.dyn_obj_trans_thresh:    # newly dynamic: flag every sighting
                self.dyn_instances.append(token)
                for pc_ts, f in enumerate(self._fids):
                    if token in self.token2idx[pc_ts]:
                        mark_f.append(f)
                        mark_i.append(self.token2idx[pc_ts][token])
        self.cloud.mark_dynamic(mark_f, mark_i)

        if len(self.poses) > 1:
            self.seg_dists.append(self.dist(np.array(self.poses[-1]), np.array(self.poses[-2])))
        self.ts += 1
        if self.sync_each_integrate:
            self._sync()

    def obs2sem_vec_space(self, rgbs, pc, pc_cam_idx, T_ego_global, ego_pose_z: float = 0):
        T_ego_world = self.T_global_world @ T_ego_global
        pose = T_ego_world[:3, -1].tolist()
        pose[2] += ego_pose_z
        semsegs = [self.semseg_model.pred(rgb)[0, 0] for rgb in rgbs]
        fid = self.cloud.integrate_records(pc, pc_cam_idx, [np.asarray(r) for r in rgbs], semsegs,
                                           T_ego_world, self.semseg_filters, 255.)
        return fid, pose, semsegs

    @staticmethod
    def cal_pose_change(p0, p1):
        return np.linalg.norm(p1 - p0)

    @staticmethod
    def get_obj_inst_poses_ts(obs_list):
        poses, tss = zip(*obs_list)
        return poses, tss

    # -- trajectories of dynamic objects (:272-414) ----------------------------
    @staticmethod
    def find_nearest_ge_idx(array, target_val):
        for idx, val in enumerate(array):
            if val >= target_val:
                return idx
        raise ValueError(f'Value {target_val} not in array {array}')

    @staticmethod
    def find_nearest_le_idx(array, target_val):
        if array[0] > target_val:
            raise ValueError(f'Value {target_val} not in array {array}')
        for idx in range(len(array) - 1):
            if array[idx + 1] > target_val:
                return idx
        return len(array) - 1

    @staticmethod
    def parse_seq_into_coherent_seqs(ts):
        runs = [[]]
        prev = ts[0] - 1
        for k, t in enumerate(ts):
            if t - prev != 1:
                runs.append([])
            runs[-1].append(k)
            prev = t
        return runs

    def parse_coherent_pose_seqs(self, poses, tss):
        return [[poses[k].tolist() for k in run]
                for run in self.parse_seq_into_coherent_seqs(tss)]

    def get_dyn_obj_trajs(self, ts_start: int = 0, ts_end: int = None, skip_ego_traj: bool = True):
        out = []
        for token, seen in self.instances.items():
            if token not in self.dyn_instances:
                continue
            poses, tss = zip(*seen)
            try:
                i0 = self.find_nearest_ge_idx(tss, ts_start)
                i1 = None if ts_end is None else self.find_nearest_le_idx(tss, ts_end) + 1
            except ValueError:
                continue
            poses, tss = poses[i0:i1], tss[i0:i1]
            out += [s for s in self.parse_coherent_pose_seqs(poses, tss) if len(s) >= 2]
        if not skip_ego_traj:
            out.append(self.poses)
        return out

    def get_split_dyn_obj_trajs(self, split_idx: int, skip_ego_traj: bool = True):
        return (self.get_dyn_obj_trajs(ts_end=split_idx), self.get_dyn_obj_trajs(ts_start=split_idx),
                self.get_dyn_obj_trajs())

    def generate_bev(self, present_idx: int = None, bev_num: int = 1, gen_future: bool = False):
        other = self.get_split_dyn_obj_trajs(present_idx)
        pcs, trajs = self._window_inputs(present_idx, gen_future, other, self.gt_lane_poses)
        return self._generate(pcs, trajs, bev_num)
