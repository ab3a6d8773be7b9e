"""Host side of the BEV generator: same class, method names and argument
meaning as the reference's `bev_generator/bev_generator.py`, with every
point-cloud operation (rotate / translate / crop / height filter / pos2grid and
the per-cell reductions) executed by libpcacc on the device.

What stays on the host, as in the reference, is the tiny per-BEV scalar and
trajectory work: the heading angle (bev_generator.py:87-93), the rotation
matrix (:732-735, numpy cos/sin so the device sees the same matrix bits as the
reference), trajectory rotate / translate / crop with midpoint bisection
(:224-237,257-371) and their pos2grid (:737-747).

Point clouds reach `generate()` in one of two forms:
  * `DeviceWindow` views of an accumulator's device-resident ring (what the
    accumulators in this package pass): no cloud bytes move;
  * numpy (M,10) arrays, exactly as the reference's callers pass them: they are
    uploaded to a scratch ring first (drop-in path).
"""
from __future__ import annotations

import ctypes as C
import os
import time
from abc import ABC, abstractmethod

import numpy as np

from .. import _lib
from ..device import DeviceCloud, make_bev_params, make_bev_params_batch

WINDOWS = ('present', 'future', 'full')


class DeviceWindow:
    """Frames [frame_begin, frame_end) of a DeviceCloud, to be shifted by
    `origin` — the device-resident stand-in for
    `np.concatenate(sem_pcs[a:b])[:, :3] - bev_frame_coords`
    (kitti360_sem_pc_accum.py:189-193)."""

    def __init__(self, cloud: DeviceCloud, frame_begin: int, frame_end: int, origin):
        self.cloud = cloud
        self.frame_begin = int(frame_begin)
        self.frame_end = int(frame_end)
        self.origin = np.asarray(origin, dtype=np.float64).reshape(3)

    @property
    def shape(self):
        self.cloud.refresh()     # exact counts; data-error flags stay pending for the caller's sync
        n = sum(self.cloud.frame_count(f) for f in range(self.frame_begin, self.frame_end))
        return (n, 10)


_OPS = {}


def _ops_cloud() -> DeviceCloud:
    """One small DeviceCloud per (process, thread, device) whose handle serves the stand-alone
    operators (static methods of the reference have no generator instance to hang it on; a handle's
    scan state must not be shared by two threads)."""
    import threading
    import torch
    key = (os.getpid(), threading.get_ident(), torch.cuda.current_device())
    c = _OPS.get(key)
    if c is None:
        c = _OPS[key] = DeviceCloud(1024, max_frames=8)
    return c


def _to_like(like, t):
    """Results go back in the kind of the input: numpy in -> numpy out, CUDA tensor in -> tensor out."""
    import torch
    return t if isinstance(like, torch.Tensor) else t.cpu().numpy()


class BEVGenerator(ABC):
    def __init__(self,
                 view_size: int,
                 pixel_size: int,
                 max_trans_radius: float = 0.,
                 zoom_thresh: float = 0.,
                 do_warp: bool = False,
                 int_scaler: float = 1.,
                 int_sep_scaler: float = 1.,
                 int_mid_threshold: float = 0.5,
                 height_filter=None):
        self.view_size = view_size
        self.pixel_size = pixel_size
        self.max_trans_radius = max_trans_radius
        self.zoom_thresh = zoom_thresh
        self.do_warp = do_warp
        self.do_aug = bool(self.max_trans_radius > 0. or self.zoom_thresh > 0.)
        self.int_scaler = int_scaler
        self.int_sep_scaler = int_sep_scaler
        self.int_mid_threshold = int_mid_threshold
        self.sem_idx = 7
        self.height_filter = height_filter
        # injectable RNG for generate_rand_aug (the reference seeds numpy's
        # global RNG from pid*time, bev_generator.py:168)
        self.rng = None
        self.py_rng = None             # `random.Random` for the warp's sign draws
        self.elevation_max = False      # north-star variant; reference = per-cell min z
        self._scratch = None            # DeviceCloud for host-array inputs

    # ------------------------------------------------------------------
    # abstract parts
    # ------------------------------------------------------------------
    @abstractmethod
    def generate_bev(self, *args, **kwargs):
        pass

    @abstractmethod
    def _assemble(self, planes, v, trajs_by_window, gt_lane_trajs, has_future):
        """Builds the output dict of variant v from the (V,3,7,P,P) planes."""

    def viz_bev(self, *a, **k):
        raise NotImplementedError('plotting is outside the B200 hot path')

    # ------------------------------------------------------------------
    # host helpers with the reference's names
    # ------------------------------------------------------------------
    @staticmethod
    def rotation_matrix_3d(ang):
        return np.array([[np.cos(ang), -np.sin(ang), 0],
                         [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])

    def pos2grid(self, mat, view_size):
        mat[:, 0:2] = np.floor(mat[:, 0:2] / view_size * self.pixel_size
                               + 0.5 * self.pixel_size)
        return mat

    @staticmethod
    def point_in_box(x, y, bx0, by0, bx1, by1):
        return (bx0 < x and x < bx1) and (by0 < y and y < by1)

    def cal_intersec_pnt(self, x0, y0, x1, y1, bbox, thresh=1e-4):
        """Midpoint refinement towards the view boundary; returns
        (x_mid, y_mid, iterations)."""
        bx0, by0, bx1, by1 = bbox
        moved = np.inf
        iters = 0
        while moved > thresh:
            xm, ym = 0.5 * (x0 + x1), 0.5 * (y0 + y1)
            first_in = self.point_in_box(x0, y0, bx0, by0, bx1, by1)
            mid_in = self.point_in_box(xm, ym, bx0, by0, bx1, by1)
            if mid_in == first_in:      # midpoint on the same side as endpoint 0
                moved = np.sqrt((xm - x0) ** 2 + (ym - y0) ** 2)
                x0, y0 = xm, ym
            else:
                moved = np.sqrt((xm - x1) ** 2 + (ym - y1) ** 2)
                x1, y1 = xm, ym
            iters += 1
        return xm, ym, iters

    def crop_trajectory(self, traj, aug_view_size, thresh=1e-4):
        """Edges of the polyline against the open view box: a vertex is kept if it starts
        an edge and lies inside; an edge that crosses the boundary contributes its
        bisection point.  (The last in-view vertex is never emitted — as in the reference.)
        Runs in libpcacc's host helper `pcacc_crop_trajectory` (same double arithmetic)."""
        traj = np.ascontiguousarray(traj[:, :3], dtype=np.float64)
        n = traj.shape[0]
        if n < 2:
            return np.zeros((0, 3))
        out = np.empty((2 * n, 3))
        m = C.c_int(0)
        _lib.check(_lib.load().pcacc_crop_trajectory(
            traj.ctypes.data_as(C.c_void_p), n, float(aug_view_size), float(thresh),
            out.ctypes.data_as(C.c_void_p), C.byref(m)))
        return out[:m.value].copy() if m.value else np.zeros((0, 3))

    def geometric_transform(self, pc_mat, rot_ang, trans_dx, trans_dy, aug_view_size,
                            is_traj=False):
        """bev_generator.py:207-237.  Trajectories are rotated and shifted on the host and cropped by
        `pcacc_crop_trajectory`; clouds go through `pcacc_preprocess_pc` (rotation, translation and
        crop in one kernel; rows of any width, kept rows in input order).  The reference also leaves
        the rotated coordinates in the caller's array; the cloud branch here does not touch it."""
        R = self.rotation_matrix_3d(rot_ang)
        if not is_traj:
            return _to_like(pc_mat, _ops_cloud().preprocess_pc(pc_mat, R, trans_dx, trans_dy, aug_view_size))
        pc_mat[:, :3] = np.matmul(R, pc_mat[:, :3].T).T
        pc_mat[:, 0] += trans_dx
        pc_mat[:, 1] += trans_dy
        return self.crop_trajectory(pc_mat, aug_view_size)

    @staticmethod
    def crop_view(pc_mat, aug_view_size: float):
        """bev_generator.py:239-256: rows with |x| and |y| strictly inside aug_view_size / 2."""
        return _to_like(pc_mat, _ops_cloud().preprocess_pc(pc_mat, crop_view=aug_view_size))

    def preprocess_pc_and_trajs(self, pc, trajs, rot_ang, trans_dx, trans_dy, aug_view_size):
        """bev_generator.py:127-160: the cloud through one `pcacc_preprocess_pc` launch (rotation,
        translation, crop, height filter, pos2grid), the trajectories through `preprocess_trajs`."""
        pc = _to_like(pc, _ops_cloud().preprocess_pc(
            pc, self.rotation_matrix_3d(rot_ang), trans_dx, trans_dy, aug_view_size, self.height_filter,
            aug_view_size, self.pixel_size))
        return pc, self.preprocess_trajs(trajs, rot_ang, trans_dx, trans_dy, aug_view_size)

    # ------------------------------------------------------------------
    # per-plane steps of generate_bev as stand-alone operators (fused in pcacc_rasterise)
    # ------------------------------------------------------------------
    def gen_sem_probmap(self, pc, sem_clss: list):
        """bev_generator.py:373-391: Dirichlet expectation of (points of sem_clss, other points) per cell."""
        sems = [self.sem_idxs[c] for c in sem_clss]
        m = _ops_cloud().cell_stats(pc, self.pixel_size, sems, self.sem_idx, finish=1, want=('sel', 'rest'))
        return _to_like(pc, m['sel'])

    def gen_intensity_map(self, pc, sem_cls: str):
        """bev_generator.py:393-409: summed intensity (column 3) of the class's points / (count + 1)."""
        m = _ops_cloud().cell_stats(pc, self.pixel_size, [self.sem_idxs[sem_cls]], self.sem_idx,
                                    weight_col=3, finish=2, want=('sel', 'wsum'))
        return _to_like(pc, m['wsum'])

    @staticmethod
    def partition_semantic_pc(pc_mat, sems: list, sem_idx: int):
        """bev_generator.py:411-432 -> (rows whose column sem_idx is in sems, the other rows), both in
        input order (`pcacc_partition_semantic_pc`: one scan ranks both halves)."""
        a, b = _ops_cloud().partition_semantic_pc(pc_mat, sems, sem_idx)
        return _to_like(pc_mat, a), _to_like(pc_mat, b)

    def gen_gridmap_count_map(self, pc, weights=None):
        """bev_generator.py:434-453: np.histogram2d of the grid coordinates (optionally weighted),
        flipped to image rows."""
        if weights is None:
            m = _ops_cloud().cell_stats(pc, self.pixel_size, want=('sel',))
            return _to_like(pc, m['sel'])
        m = _ops_cloud().cell_stats(pc, self.pixel_size, weights=weights, want=('wsum',))
        return _to_like(pc, m['wsum'])

    @staticmethod
    def dirichlet_dist_expectation(gridmaps, obs_weight=1):
        """bev_generator.py:455-481 -> list of posterior maps."""
        m = _ops_cloud().dirichlet_expectation(gridmaps, obs_weight)
        like = gridmaps if not isinstance(gridmaps, (list, tuple)) else gridmaps[0]
        return [_to_like(like, m[g]) for g in range(m.shape[0])]

    def get_rgb_maps(self, pc):
        """sem_bev.py:619-669 == rgb_bev.py:133-183: per-cell medians of the r, g, b columns (4..6) of a cloud in grid
        coordinates, `rgb_fill` where a cell is empty -> (red_map, green_map, blue_map) float64.
        The medians come out of the rasteriser's RGB planes (float64 copy, median / 255): a median of
        0..255 integers is a multiple of 1/2, so round(2 * 255 * plane) / 2 recovers it exactly; empty
        cells are told apart by rasterising with a fill value no colour can produce."""
        P = self.pixel_size
        c = self._pad10(np.array(pc.cpu().numpy() if hasattr(pc, 'cpu') else pc, dtype=np.float64))
        c[:, 7:] = 0.
        c[:, 0] = c[:, 0] + 0.5 - 0.5 * P
        c[:, 1] = c[:, 1] + 0.5 - 0.5 * P
        fill = getattr(self, 'rgb_fill', 0)
        saved = self.view_size, self.height_filter
        self.view_size, self.height_filter, self.rgb_fill = float(P), None, -1.
        try:
            cloud = self._scratch_cloud(c.shape[0])
            fid = cloud.integrate_cloud(c)
            prm = self._bev_params(fid, fid + 1, fid + 1, np.zeros(3), 0., 0., 0., float(P))
            _, p64, _ = cloud.rasterise([prm], P, want_f64=True)
            if cloud.sync() & _lib.FLAG_ATTR_RANGE:
                raise ValueError('r, g, b must be integers in 0..255')
        finally:
            (self.view_size, self.height_filter), self.rgb_fill = saved, fill
        med = (p64[0, 0, 2:5] * 510.).round() / 2.
        med = med.where(med >= 0., med.new_tensor(float(fill)))
        med = med if hasattr(pc, 'cpu') else med.cpu().numpy()
        return med[0], med[1], med[2]

    @staticmethod
    def extract_aug_dict(augs: dict):
        return augs['max_translation_radius'], augs['zoom_threshold']

    def preprocess_trajs(self, trajs, rot_ang, trans_dx, trans_dy, aug_view_size):
        out = []
        for t in trajs:
            t = np.array(t, dtype=float)
            t = t.reshape(-1, 3) if t.size else np.zeros((0, 3))
            t = self.geometric_transform(t, rot_ang, trans_dx, trans_dy, aug_view_size,
                                         is_traj=True)
            out.append(self.pos2grid(t, aug_view_size))
        return out

    @staticmethod
    def extract_pc_dict(pcs):
        return pcs['pc_present'], pcs['pc_future'], pcs['pc_full']

    @staticmethod
    def extract_ego_traj_dict(trajs):
        return trajs['ego_traj_present'], trajs['ego_traj_future'], trajs['ego_traj_full']

    @staticmethod
    def extract_other_traj_dicts(trajs):
        return (trajs['other_trajs_present'], trajs['other_trajs_future'],
                trajs['other_trajs_full'])

    @staticmethod
    def extract_gt_lane_dicts(trajs):
        return trajs['gt_lanes']

    @staticmethod
    def heading_angle(ego_traj_present):
        rot_ang = 0.5 * np.pi
        if len(ego_traj_present) > 1:
            dx = ego_traj_present[-1][0] - ego_traj_present[-2][0]
            dy = ego_traj_present[-1][1] - ego_traj_present[-2][1]
            rot_ang += np.arctan2(dy, dx)
        return np.pi - rot_ang

    # ------------------------------------------------------------------
    # device plumbing
    # ------------------------------------------------------------------
    def _bev_params(self, fb, fs, fe, origin, rot_ang, dx, dy, view):
        sem_idxs = getattr(self, 'sem_idxs', None) or {
            'road': -1, 'car': -1, 'truck': -1, 'bus': -1, 'motorcycle': -1}
        return make_bev_params(fb, fs, fe, origin, self.rotation_matrix_3d(rot_ang), dx, dy, view,
                               self.height_filter, self.int_scaler, self.int_sep_scaler,
                               self.int_mid_threshold, getattr(self, 'rgb_fill', 0), sem_idxs,
                               self.elevation_max)

    def _rot_of(self, aug):
        """rotation_matrix_3d(aug['rot_ang']), computed once per augmentation dict."""
        R = aug.get('_R')
        if R is None:
            R = aug['_R'] = self.rotation_matrix_3d(aug['rot_ang'])
        return R

    def _scratch_cloud(self, n_pts):
        if self._scratch is None or self._scratch.capacity < n_pts + 16:
            if self._scratch is not None:
                self._scratch.close()
            self._scratch = DeviceCloud(int(n_pts * 1.25) + 1024, max_frames=8)
        self._scratch.reset()
        return self._scratch

    def _rasterise_windows(self, pcs, augs, warps=None):
        """Rasterise (and optionally warp) on the device, then copy the planes to the host."""
        pending, has_future, cloud = self._rasterise_windows_begin(pcs, augs, warps)
        return cloud.planes_to_host_finish(pending), has_future

    def _rasterise_windows_begin(self, pcs, augs, warps=None):
        """Enqueue rasterise (+ warp) and the device -> host copy; returns (pending, has_future,
        cloud) for `cloud.planes_to_host_finish(pending)`."""
        planes, has_future, cloud = self._rasterise_device(pcs, augs)
        if warps is not None:
            planes = cloud.warp_planes(planes.contiguous(), [w['imap'] for w in warps],
                                       [w['jmap'] for w in warps])
        return cloud.planes_to_host_begin(planes.contiguous()), has_future, cloud

    def _rasterise_device(self, pcs, augs):
        """pcs: the reference's dict. Returns (planes (V,3,7,P,P) float16 device tensor,
        has_future, the DeviceCloud that produced them)."""
        P = self.pixel_size
        pp, pf, pa = self.extract_pc_dict(pcs)
        has_future = pf is not None
        V = len(augs)

        def params(cloud, fb, fs, fe, origin):
            sem_idxs = getattr(self, 'sem_idxs', None) or {
                'road': -1, 'car': -1, 'truck': -1, 'bus': -1, 'motorcycle': -1}
            return make_bev_params_batch(
                fb, fs, fe, origin, [self._rot_of(a) for a in augs],
                [a['trans_dx'] for a in augs], [a['trans_dy'] for a in augs],
                [a['zoom_scalar'] * self.view_size for a in augs], self.height_filter,
                self.int_scaler, self.int_sep_scaler, self.int_mid_threshold,
                getattr(self, 'rgb_fill', 0), sem_idxs, self.elevation_max)

        if isinstance(pp, DeviceWindow):
            cloud = pp.cloud
            if has_future:
                same = (pf.cloud is cloud and pa.cloud is cloud
                        and pf.frame_begin == pp.frame_end and pa.frame_begin == pp.frame_begin
                        and pa.frame_end == pf.frame_end
                        and np.array_equal(pp.origin, pf.origin)
                        and np.array_equal(pp.origin, pa.origin))
                if same:        # full = present ++ future: one pass does all three
                    planes, _, _ = cloud.rasterise(
                        params(cloud, pp.frame_begin, pp.frame_end, pf.frame_end, pp.origin), P)
                    return planes, True, cloud
                # general case: each window rasterised as a 'present' window
                outs = []
                for w in (pp, pf, pa):
                    pl, _, _ = w.cloud.rasterise(
                        params(w.cloud, w.frame_begin, w.frame_end, w.frame_end, w.origin), P)
                    outs.append(pl[:, 0])
                import torch
                return torch.stack(outs, dim=1), True, pp.cloud
            planes, _, _ = cloud.rasterise(
                params(cloud, pp.frame_begin, pp.frame_end, pp.frame_end, pp.origin), P)
            return planes, False, cloud

        # host arrays: upload into a scratch ring as frames
        clouds = [np.asarray(pp, dtype=np.float64)]
        if has_future:
            clouds += [np.asarray(pf, dtype=np.float64), np.asarray(pa, dtype=np.float64)]
        clouds = [self._pad10(c) for c in clouds]
        cloud = self._scratch_cloud(sum(c.shape[0] for c in clouds))
        fids = [cloud.integrate_cloud(c) for c in clouds]
        zero = np.zeros(3)
        pl, _, _ = cloud.rasterise(
            params(cloud, fids[0], fids[1] if has_future else fids[0] + 1,
                   fids[1] + 1 if has_future else fids[0] + 1, zero), P)
        if not has_future:
            return pl, False, cloud
        # pc_full is an input of its own in the reference API: rasterise it as well
        plf, _, _ = cloud.rasterise(params(cloud, fids[2], fids[2] + 1, fids[2] + 1, zero), P)
        pl[:, 2] = plf[:, 0]
        return pl, True, cloud

    @staticmethod
    def _pad10(c):
        c = c.reshape(-1, c.shape[-1]) if c.ndim == 2 else c.reshape(0, 10)
        if c.shape[1] < 10:
            c = np.concatenate([c, np.zeros((c.shape[0], 10 - c.shape[1]))], axis=1)
        return np.ascontiguousarray(c[:, :10])

    # ------------------------------------------------------------------
    # public generation API (reference names)
    # ------------------------------------------------------------------
    def generate_batch(self, pcs: dict, trajs: dict, augs: list) -> list:
        """All variants in `augs` (dicts with rot_ang, trans_dx, trans_dy,
        zoom_scalar, do_warping) rasterised in ONE batch of launches — the
        device-side replacement of the reference's Pool(bev_num).map
        (kitti360_sem_pc_accum.py:236-241)."""
        ego_p, ego_f, ego_a = self.extract_ego_traj_dict(trajs)
        _, pc_future, _ = self.extract_pc_dict(pcs)
        if pc_future is None:
            # the reference dies here too: trajs_future is never bound
            # (bev_generator.py:111-123); every caller passes gen_future=True
            raise UnboundLocalError(
                "cannot access local variable 'trajs_future': generate() requires pc_future "
                '(call generate_bev(..., gen_future=True))')
        full = []
        for a in augs:
            a = dict(a)
            a.setdefault('rot_ang', 0.)
            a.setdefault('trans_dx', 0.)
            a.setdefault('trans_dy', 0.)
            a.setdefault('zoom_scalar', 1.)
            if not a.get('do_warping', False):
                a['rot_ang'] = self.heading_angle(ego_p)
            full.append(a)
        warps = None
        if self.do_warp:
            # one polynomial warp per BEV, drawn like sem_bev.py:121-129
            warps = [self.draw_warp() for _ in full]
        # the device works (rasterise, warp, copy to pinned memory) while the host prepares
        # the trajectories
        pending, has_future, cloud = self._rasterise_windows_begin(pcs, full, warps)
        # (a lazily filled trajs dict extracts the other objects' trajectories here, under the kernels)
        oth_p, oth_f, oth_a = self.extract_other_traj_dicts(trajs)
        groups = [[ego_p] + list(oth_p), [ego_f] + list(oth_f), [ego_a] + list(oth_a)]
        if 'gt_lanes' in trajs:
            groups.append(list(self.extract_gt_lane_dicts(trajs)))
        done = self.preprocess_trajs_batch(groups, full)
        planes = cloud.planes_to_host_finish(pending)
        bevs = []
        for v, a in enumerate(full):
            tw = {'present': done[v][0], 'future': done[v][1], 'full': done[v][2]}
            lanes = None
            if 'gt_lanes' in trajs:
                lanes = [ln for ln in done[v][3] if ln.shape[0] > 0]
            if warps is not None:
                w = warps[v]
                for k in tw:
                    tw[k] = self.warp_trajs(tw[k], w['a_1'], w['a_2'], w['b_1'], w['b_2'], w['i_mid'],
                                            w['j_mid'], w['i_warp'], w['j_warp'])
                if lanes is not None:
                    lanes = self.warp_trajs(lanes, w['a_1'], w['a_2'], w['b_1'], w['b_2'], w['i_mid'],
                                            w['j_mid'], w['i_warp'], w['j_warp'])
            bevs.append(self._assemble(planes, v, tw, lanes, has_future))
        return bevs

    def preprocess_trajs_batch(self, groups, augs):
        """preprocess_trajs for several lists of polylines and several augmentation variants in
        one call of libpcacc's host helper `pcacc_preprocess_trajectories` (same double
        arithmetic as geometric_transform -> crop_trajectory -> pos2grid).
        Returns out[v][g] = list of (m,3) arrays."""
        arrs = []
        for g in groups:
            for t in g:
                if not (type(t) is np.ndarray and t.dtype == np.float64 and t.ndim == 2 and t.shape[1] == 3):
                    t = np.asarray(t, dtype=np.float64)
                    t = t.reshape(-1, 3) if t.size else np.zeros((0, 3))
                arrs.append(t)
        n_traj, n_var = len(arrs), len(augs)
        lens = [a.shape[0] for a in arrs]
        off_l = [0]
        for n in lens:
            off_l.append(off_l[-1] + n)
        off = np.array(off_l, dtype=np.int32)
        N = off_l[-1]
        pts = np.concatenate(arrs, axis=0) if N else np.zeros((0, 3))
        var = np.empty((n_var, 12), dtype=np.float64)
        for v, a in enumerate(augs):
            var[v, :9] = self._rot_of(a).reshape(9)
            var[v, 9], var[v, 10] = a['trans_dx'], a['trans_dy']
            var[v, 11] = a['zoom_scalar'] * self.view_size
        out = np.empty((n_var, 2 * N, 3), dtype=np.float64)
        cnt = np.zeros((n_var, max(n_traj, 1)), dtype=np.int32)
        _lib.check(_lib.load().pcacc_preprocess_trajectories(
            pts.ctypes.data, off.ctypes.data, n_traj, var.ctypes.data, n_var, int(self.pixel_size), 1e-4,
            out.ctypes.data, cnt.ctypes.data))
        cnt_l = cnt.tolist()
        res = []
        for v in range(n_var):
            per_group, t = [], 0
            for g in groups:
                lst = []
                ov, cv = out[v], cnt_l[v]
                for _ in g:
                    b = 2 * off_l[t]
                    lst.append(ov[b:b + cv[t]])
                    t += 1
                per_group.append(lst)
            res.append(per_group)
        return res

    # ------------------------------------------------------------------
    # polynomial warp (bev_generator.py:482-698): parameters and trajectories on the
    # host, the dense gather on the device (pcacc_warp_planes)
    # ------------------------------------------------------------------
    @staticmethod
    def cal_warp_params(idx_0, idx_1, idx_max):
        a_1 = (idx_1 - idx_0 ** 2 / idx_max) / (idx_0 * (1.0 - idx_0 / idx_max))
        a_2 = (1.0 - a_1) / idx_max
        return (a_1, a_2)

    def get_random_warp_params(self, mean_ratio, max_ratio, I, J):
        """Two normal draws, then two uniform draws for the signs.  `self.rng` /
        `self.py_rng` replace the reference's global numpy / `random` state when set."""
        import random as _random
        np_rng = self.rng if self.rng is not None else np.random
        py_rng = getattr(self, 'py_rng', None) or _random
        max_val = max_ratio * (I / 2.0)
        mean_val = mean_ratio * max_val
        i_warp = np_rng.normal(mean_val, max_val)
        j_warp = np_rng.normal(mean_val, max_val)
        if abs(i_warp) > max_val:
            i_warp = max_val
        if abs(j_warp) > max_val:
            j_warp = max_val
        if py_rng.random() < 0.5:
            i_warp = -i_warp
        if py_rng.random() < 0.5:
            j_warp = -j_warp
        return (int(I / 2) + i_warp, int(J / 2) + j_warp)

    @staticmethod
    def warp_index_map(c_1, c_2, n):
        """Source index of every warped index: clamp(rint(c_1 k + c_2 k^2), 0, n-1)."""
        k = np.arange(n)
        return np.clip(np.rint(c_1 * k + c_2 * k ** 2).astype(np.int64), 0, n - 1)

    def draw_warp(self):
        P = self.pixel_size
        i_mid = j_mid = int(P / 2)
        i_warp, j_warp = self.get_random_warp_params(0.15, 0.30, P, P)
        a_1, a_2 = self.cal_warp_params(i_warp, i_mid, P - 1)
        b_1, b_2 = self.cal_warp_params(j_warp, j_mid, P - 1)
        return dict(a_1=a_1, a_2=a_2, b_1=b_1, b_2=b_2, i_mid=i_mid, j_mid=j_mid, i_warp=i_warp,
                    j_warp=j_warp, imap=self.warp_index_map(a_1, a_2, P),
                    jmap=self.warp_index_map(b_1, b_2, P))

    def warp_dense_probmaps(self, probmaps, a_1, a_2, b_1, b_2):
        """(n,h,w) maps -> warped maps, B[:, jw, iw] = A[:, j(jw), i(iw)].  A gather: done with
        index tensors on the device for host arrays of any dtype; the rasteriser's own float16
        planes go through pcacc_warp_planes instead."""
        import torch
        probmaps = np.asarray(probmaps)
        n, I, J = probmaps.shape
        imap = torch.from_numpy(self.warp_index_map(a_1, a_2, I)).cuda()
        jmap = torch.from_numpy(self.warp_index_map(b_1, b_2, J)).cuda()
        t = torch.from_numpy(np.ascontiguousarray(probmaps)).cuda()
        return t[:, jmap[:, None], imap[None, :]].cpu().numpy()

    @staticmethod
    def warp_point(x, y, a_1, a_2, b_1, b_2, I, J):
        import math

        def inv(v, c_1, c_2, n):
            if math.isclose(c_2, 0.0, abs_tol=1e-6):
                w = v
            else:
                w = int(np.rint((-c_1 + np.sqrt(c_1 ** 2 + 4.0 * c_2 * v)) / (2 * c_2)))
            return 0 if w < 0 else (n - 1 if w >= n else w)
        return (inv(x, a_1, a_2, I), inv(y, b_1, b_2, J))

    def warp_points(self, pnt_list, a_1, a_2, b_1, b_2, I, J):
        return [self.warp_point(p[0], p[1], a_1, a_2, b_1, b_2, I, J) for p in pnt_list]

    @staticmethod
    def _warp_axis(v, c_1, c_2, n):
        """warp_point's inverse map for a whole coordinate column: the same numpy ufuncs applied
        element-wise give the same bits as the reference's per-point calls."""
        import math
        if math.isclose(c_2, 0.0, abs_tol=1e-6):
            w = np.array(v, dtype=np.float64)
        else:
            w = np.rint((-c_1 + np.sqrt(c_1 ** 2 + 4.0 * c_2 * v)) / (2 * c_2))
            if np.isnan(w).any():
                raise ValueError('cannot convert float NaN to integer')   # int(np.rint(nan)) in the reference
        return np.where(w < 0, 0, np.where(w >= n, n - 1, w))

    def warp_sparse_points(self, pnts, a_1, a_2, b_1, b_2, i_mid, j_mid, i_warp, j_warp):
        # the j warp is applied reversed, as in the reference (bev_generator.py:530-533)
        b_1_rev, b_2_rev = self.cal_warp_params(self.pixel_size - j_warp, j_mid, self.pixel_size - 1)
        if pnts.shape[0]:
            x = self._warp_axis(pnts[:, 0], a_1, a_2, self.pixel_size)
            y = self._warp_axis(pnts[:, 1], b_1_rev, b_2_rev, self.pixel_size)
            pnts[:, 0] = x
            pnts[:, 1] = y
        return pnts

    def warp_trajs(self, trajs, a_1, a_2, b_1, b_2, i_mid, j_mid, i_warp, j_warp):
        return [self.warp_sparse_points(t, a_1, a_2, b_1, b_2, i_mid, j_mid, i_warp, j_warp)
                for t in trajs]

    def generate(self, pcs: dict, trajs: dict, rot_ang: float = 0., trans_dx: float = 0.,
                 trans_dy: float = 0., zoom_scalar: float = 1., do_warping: bool = False):
        return self.generate_batch(pcs, trajs, [dict(
            rot_ang=rot_ang, trans_dx=trans_dx, trans_dy=trans_dy, zoom_scalar=zoom_scalar,
            do_warping=do_warping)])[0]

    def rand_aug_params(self, do_warping: bool = True) -> dict:
        """Draws (rot, trans, zoom) in the reference's order
        (bev_generator.py:168-180). `self.rng` (np.random.Generator or
        RandomState) makes it reproducible; default = the reference's seeding."""
        rng = self.rng
        if rng is None:
            rng = np.random.RandomState((os.getpid() * int(time.time())) % 123456789)
        rnd = rng.random if hasattr(rng, 'random') else rng.random_sample
        rot_ang = 2 * np.pi * rnd()
        trans_r = self.max_trans_radius * rnd()
        trans_ang = 2 * np.pi * rnd()
        zoom = rng.normal(0, 0.1)
        zoom = min(max(zoom, -self.zoom_thresh), self.zoom_thresh)
        return dict(rot_ang=rot_ang, trans_dx=trans_r * np.cos(trans_ang),
                    trans_dy=trans_r * np.sin(trans_ang), zoom_scalar=1 + zoom,
                    do_warping=do_warping)

    def generate_rand_aug(self, pcs: dict, trajs: dict, do_warping: bool = True):
        return self.generate_batch(pcs, trajs, [self.rand_aug_params(do_warping)])[0]

    def generate_multiproc(self, bev_gen_inputs):
        pcs, trajs = bev_gen_inputs
        if self.do_aug:
            return self.generate_rand_aug(pcs, trajs)
        return self.generate(pcs, trajs)

    def generate_rand_aug_multiproc(self, bev_gen_inputs):
        pcs, trajs = bev_gen_inputs
        return self.generate_rand_aug(pcs, trajs, do_warping=True)
