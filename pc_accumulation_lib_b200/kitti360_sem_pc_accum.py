"""Kitti360SemanticPointCloudAccumulator — drop-in for the reference's
`kitti360_sem_pc_accum.py`: same constructor, `integrate(observations)`
returning the number of evicted frames, `generate_bev(present_idx, bev_num,
gen_future)` returning a list of BEV dicts.

One departure, forced by scope: the reference obtains the
ego-motion `T_new_prev` from Open3D point-to-plane ICP
(kitti360_sem_pc_accum.py:123-127).  ICP is a third-party dependency outside
the hot path, so the transform is an input here: either a 4th element of the
observation tuple `(rgb, pc, sem_gt, T_new_prev)` or `accumulator.pose_source`
(a callable `pc -> 4x4`, or an iterator of 4x4 matrices).
"""
from __future__ import annotations

import numpy as np

from .sem_pc_accum import SemanticPointCloudAccumulator


class Kitti360SemanticPointCloudAccumulator(SemanticPointCloudAccumulator):
    def __init__(self, horizon_dist: float, calib_params: dict, icp_threshold: float,
                 semseg_onnx_path, semseg_filters: list, sem_idxs: dict, use_gt_sem: bool,
                 bev_params: dict, **ring_kwargs):
        super().__init__(horizon_dist, icp_threshold, semseg_onnx_path, semseg_filters, sem_idxs,
                         use_gt_sem, bev_params, **ring_kwargs)
        self.H_velo_cam = calib_params['h_velo_cam']
        self.P_cam_frame = calib_params['p_cam_frame']
        self.P_velo_frame = calib_params['p_velo_frame']
        self.pose_source = None

    def _next_transform(self, pc):
        src = self.pose_source
        if src is None:
            raise NotImplementedError(
                'Open3D ICP is outside the B200 hot path: pass T_new_prev as the 4th element of '
                'the observation tuple or set accumulator.pose_source')
        return np.asarray(src(pc) if callable(src) else next(src), dtype=np.float64)

    def integrate(self, observations: list):
        obs = observations[0]
        if len(obs) >= 4:
            rgb, pc, sem_gt, T_new_prev = obs[:4]
            T_new_prev = np.asarray(T_new_prev, dtype=np.float64)
        else:
            rgb, pc, sem_gt = obs
            T_new_prev = self._next_transform(pc)
        if not self.use_gt_sem:
            sem_gt = None

        # previous poses / clouds move into the new ego frame first: the new
        # frame is not affected, so the order relative to the append is free
        if len(self.poses) > 0:
            self.update_poses(T_new_prev)
            self.update_sem_pcs(T_new_prev)

        fid, pose, semseg, _ = self.obs2sem_vec_space(rgb, pc, sem_gt, T_new_prev)
        self._fids.append(fid)
        self.poses.append(pose)
        self.rgbs.append(rgb)
        self.semsegs.append(semseg)

        idx = 0
        if len(self.poses) > 1:
            idx, path_length = self.remove_observations()
        if self.sync_each_integrate:
            self._sync()
        return idx

    def obs2sem_vec_space(self, rgb, pc, sem_gt=None, T_new_prev=None):
        """Appends the observation's semantic cloud to the device ring; returns
        (frame id, pose, semseg, T_new_prev).  K1-K4 of SURVEY.md §2.1 in one
        kernel: projection, in-image mask, RGB + class gather, class filter,
        inst = 0, dyn = 0, order-preserving compaction."""
        T_new_origin = np.matmul(self.T_prev_origin, T_new_prev)
        if sem_gt is None:
            semseg = self.semseg_model.pred(rgb)[0, 0]
            fid = self.cloud.integrate_frustum(pc, self.P_velo_frame, np.asarray(rgb), semseg,
                                               self.semseg_filters)
        else:
            semseg = None
            fid = self.cloud.integrate_gt(pc, sem_gt, self.semseg_filters)
        self.T_prev_origin = T_new_origin
        return fid, [0., 0., 0.], semseg, T_new_prev

    def generate_bev(self, present_idx: int = None, bev_num: int = 1, gen_future: bool = False):
        pcs, trajs = self._window_inputs(present_idx, gen_future)
        return self._generate(pcs, trajs, bev_num)
