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

import bisect

import numpy as np

from . import _lib
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
        self._dyn_set = set()            # the same tokens, for O(1) membership
        self._first_xy = {}              # token -> (x, y) of its first sighting (Python floats)
        self._pose_lists = {}            # token -> ([[x, y, z], ...], [ts, ...]) mirror of `instances`
        self.staging = _lib.STAGE_AUTO   # host staging mode of integrate (pcacc.h PCACC_STAGE_*)
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

    def _boxes_to_world(self, centres):
        """homo_transform(T_global_world, expand_dims(c, 0))[0] for every box centre in ONE
        numpy call.  A single-point product takes numpy's matrix-vector path, whose rounding
        differs from the matrix-matrix one; a stacked (n,4,1) operand
        runs that same matrix-vector kernel once per box, so the centres — which reach the
        output through `trajs_*` — keep the reference's bits (tests/test_gpu_api.py)."""
        c = np.asarray(centres, dtype=np.float64).reshape(-1, 3)
        col = np.ones((c.shape[0], 4, 1))
        col[:, :3, 0] = c
        return np.matmul(self.T_global_world, col)[:, :3, 0]

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

        # fake detector / tracker (nuscenes_oracle_sem_pc_accum.py:191-250): host bookkeeping,
        # device flag updates
        ts = self.ts
        seen_now = {'ts': ts}
        self.token2idx.append(seen_now)
        tokens = obs['inst_tokens']
        if len(tokens):
            inst_cls = obs['inst_cls']
            centres = self._boxes_to_world(obs['inst_center'])
            xyz = centres.tolist()
            tracked, instances, dyn = self.track_inst_clss, self.instances, self._dyn_set
            thresh = self.dyn_obj_trans_thresh
            mark_f, mark_i = [], []
            for idx, token in enumerate(tokens):
                if inst_cls[idx] not in tracked:
                    continue
                seen = instances.get(token)
                if seen is None:
                    seen = instances[token] = []
                    self._first_xy[token] = xyz[idx][:2]
                    self._pose_lists[token] = ([], [])
                seen.append((centres[idx], ts))
                cached = self._pose_lists.get(token)
                if cached is not None:
                    cached[0].append(xyz[idx])
                    cached[1].append(ts)
                seen_now[token] = idx
                if token in dyn:                     # known mover: flag the new sweep
                    mark_f.append(fid)
                    mark_i.append(idx)
                    continue
                if len(seen) < 2:
                    continue
                # moved = ||last - first|| > thresh, decided on the squared distance; only
                # within 1e-9 of the threshold is the reference's own expression evaluated
                first = self._first_xy.get(token)
                if first is None:        # `instances` was filled by the caller
                    first = self._first_xy[token] = seen[0][0][:2].tolist()
                x0, y0 = first
                dx, dy = xyz[idx][0] - x0, xyz[idx][1] - y0
                d2, t2 = dx * dx + dy * dy, thresh * thresh
                if abs(d2 - t2) <= 1e-9 * max(1.0, t2):
                    is_dyn = self.cal_pose_change(seen[0][0][:2], seen[-1][0][:2]) > thresh
                else:
                    is_dyn = d2 > t2
                if is_dyn:                           # newly dynamic: flag every sighting
                    self.dyn_instances.append(token)
                    dyn.add(token)
                    for pc_ts, f in enumerate(self._fids):
                        prev = self.token2idx[pc_ts].get(token)
                        if prev is not None:
                            mark_f.append(f)
                            mark_i.append(prev)
            if mark_f:
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
        if isinstance(pc, np.ndarray) and not any(hasattr(r, 'is_cuda') for r in rgbs):
            # host arrays: one C call; page-locked arrays are read in place, pageable ones go
            # through the sparse staging mode (device.py: integrate_records_host)
            fid = self.cloud.integrate_records_host(
                pc, pc_cam_idx, [r if type(r) is np.ndarray else np.asarray(r) for r in rgbs],
                semsegs, T_ego_world, self.semseg_filters, 255., self.staging)
        else:
            fid = self.cloud.integrate_records(pc, pc_cam_idx, [np.asarray(r) for r in rgbs],
                                               semsegs, T_ego_world, self.semseg_filters, 255.)
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
    def get_tf_pose(inst_tf):
        """(x, y, z) of a (4,4) instance pose matrix (nuscenes_oracle_sem_pc_accum.py:701-709)."""
        return inst_tf[:3, -1]

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
        """nuscenes_oracle_sem_pc_accum.py:272-347: for every dynamic token the sightings with
        ts_start <= ts <= ts_end, split into runs of consecutive time steps, runs of >= 2
        poses kept (lists of [x, y, z])."""
        out = []
        dyn = self._dyn_set if len(self._dyn_set) == len(self.dyn_instances) else set(self.dyn_instances)
        for token, seen in self.instances.items():
            if token not in dyn:
                continue
            cached = self._pose_lists.get(token)
            if cached is not None and len(cached[1]) == len(seen):
                plist, tss = cached          # [x, y, z] lists kept since integrate()
            else:                            # `instances` was filled by the caller
                plist, tss = [p.tolist() for p, _ in seen], [t for _, t in seen]
            # find_nearest_ge_idx / find_nearest_le_idx on an increasing sequence
            i0 = bisect.bisect_left(tss, ts_start)
            if i0 >= len(tss):
                continue
            if ts_end is None:
                i1 = len(tss)
            else:
                if tss[0] > ts_end:
                    continue
                i1 = bisect.bisect_right(tss, ts_end)
            if i0 >= i1:
                # a gap in the sightings covers the whole interval: the reference indexes an
                # empty tuple here (parse_seq_into_coherent_seqs, :400)
                raise IndexError('tuple index out of range')
            # runs of consecutive time steps; the usual case is one unbroken run
            if tss[i1 - 1] - tss[i0] == i1 - 1 - i0:
                if i1 - i0 >= 2:
                    out.append([p[:] for p in plist[i0:i1]])
                continue
            a = i0
            for k in range(i0 + 1, i1 + 1):
                if k == i1 or tss[k] - tss[k - 1] != 1:
                    if k - a >= 2:
                        out.append([p[:] for p in plist[a:k]])
                    a = k
        if not skip_ego_traj:
            out.append(self.poses)
        return out

    def get_split_dyn_obj_trajs(self, split_idx: int, skip_ego_traj: bool = True):
        return (self.get_dyn_obj_trajs(ts_end=split_idx), self.get_dyn_obj_trajs(ts_start=split_idx),
                self.get_dyn_obj_trajs())

    def reset(self):
        """New scene on the same accumulator (run_nuscenes_bev_gen.py builds a new accumulator per
        scene, :165,203; this keeps the ring, its workspace and the staging buffers)."""
        self.cloud.reset()
        self._fids, self.poses, self.seg_dists, self.rgbs, self.semsegs = [], [], [], [], []
        self.T_global_world = None
        self.instances, self.dyn_instances, self.token2idx, self.ts = {}, [], [], 0
        self._dyn_set, self._first_xy, self._pose_lists = set(), {}, {}
        self.ego_global_xs, self.ego_global_ys = [], []

    def generate_bev(self, present_idx: int = None, bev_num: int = 1, gen_future: bool = False):
        pcs, trajs = self._window_inputs(present_idx, gen_future,
                                         lambda: self.get_split_dyn_obj_trajs(present_idx), self.gt_lane_poses)
        return self._generate(pcs, trajs, bev_num)
