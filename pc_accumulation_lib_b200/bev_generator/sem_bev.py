"""SemBEVGenerator — same constructor, `generate_bev` signature and output dict
as the reference's `bev_generator/sem_bev.py` (keys road_/intensity_/rgb_/
dynamic_/elevation_/trajs_ x present/future/full, float16, rgb (3,P,P)).
The seven grids of each window come out of libpcacc's rasteriser."""
from __future__ import annotations

import numpy as np

from .. import _lib
from .bev_generator import BEVGenerator, WINDOWS, _ops_cloud, _to_like


class SemBEVGenerator(BEVGenerator):
    def __init__(self,
                 sem_idxs: dict,
                 view_size: int,
                 pixel_size: int,
                 max_trans_radius: float = 0.,
                 zoom_thresh: float = 0.,
                 do_warp: bool = False,
                 int_scaler: float = 1.,
                 int_sep_scaler: float = 1.,
                 int_mid_threshold: float = 0.5,
                 height_filter=None,
                 rgb_fill: int = 0):
        super().__init__(view_size, pixel_size, max_trans_radius, zoom_thresh, do_warp, int_scaler,
                         int_sep_scaler, int_mid_threshold, height_filter)
        self.sem_idxs = sem_idxs
        self.dyn_idx = 9
        self.rgb_fill = rgb_fill

    def _assemble(self, planes, v, trajs_by_window, gt_lane_trajs, has_future):
        bev = {}
        for wi, w in enumerate(WINDOWS if has_future else WINDOWS[:1]):
            pl = planes[v, wi]
            bev[f'road_{w}'] = pl[0]
            bev[f'trajs_{w}'] = trajs_by_window[w]
            bev[f'intensity_{w}'] = pl[1]
            bev[f'rgb_{w}'] = pl[2:5]
            bev[f'dynamic_{w}'] = pl[5]
            bev[f'elevation_{w}'] = pl[6]
        if gt_lane_trajs is not None:
            bev['gt_lanes'] = gt_lane_trajs
        return bev

    def generate_bev(self, pc_present, pc_future, pc_full, trajs_present, trajs_future,
                     trajs_full, gt_lane_trajs=None):
        """Reference entry point on PRE-PROCESSED clouds: columns 0,1 already
        hold integer grid coordinates (output of pos2grid).  The clouds are
        uploaded and binned by the same kernels: a point at grid (i, j) is placed
        at the centre of its cell of a P-metre, P-pixel view."""
        P = self.pixel_size

        def centre(pc):
            if pc is None:
                return None
            pc = self._pad10(np.array(pc, dtype=np.float64))
            pc[:, 0] = pc[:, 0] + 0.5 - 0.5 * P
            pc[:, 1] = pc[:, 1] + 0.5 - 0.5 * P
            return pc

        pcs = {'pc_present': centre(pc_present), 'pc_future': centre(pc_future),
               'pc_full': centre(pc_full)}
        saved = self.view_size, self.height_filter
        self.view_size, self.height_filter = float(P), None
        try:
            aug = [dict(rot_ang=0., trans_dx=0., trans_dy=0., zoom_scalar=1., do_warping=True)]
            planes, has_future = self._rasterise_windows(pcs, aug)
        finally:
            self.view_size, self.height_filter = saved
        tw = {'present': trajs_present, 'future': trajs_future, 'full': trajs_full}
        return self._assemble(planes, 0, tw, gt_lane_trajs, has_future)

    def static_obj_partitioning_by_elev(self, pc, elev_thresh: float):
        """bev_generator/sem_bev.py:556-591 (unused by the reference's own pipeline): per-cell
        minimum z of a cloud in grid coordinates, points more than `elev_thresh` above it get
        column 8 set to 1 IN PLACE -> (pc_static, pc_dynamic, elevmap, elevmap_obs_mask).
        Runs in `pcacc_static_obj_partitioning`."""
        from .bev_generator import DeviceCloud
        if self._scratch is None:
            self._scratch = DeviceCloud(1024, max_frames=8)
        flagged, elev, obs = self._scratch.static_obj_partitioning(pc, self.pixel_size, elev_thresh)
        if self._scratch.sync() & _lib.FLAG_ATTR_RANGE:
            raise IndexError('grid index out of bounds for the elevation map (or NaN z)')
        out = flagged.cpu().numpy()
        pc[:, 8] = out[:, 8]
        return pc[pc[:, 8] == 0], pc[pc[:, 8] == 1], elev.cpu().numpy(), obs.cpu().numpy()

    # -- the per-plane steps of generate_bev as stand-alone operators ----------------------
    def road_marking_transform(self, intensity_map, int_scaler: float, int_sep_scaler: float,
                               int_mid_threshold: float):
        """sem_bev.py:593-613: int_scaler * sigmoid(int_sep_scaler * (I - int_mid_threshold)) capped
        at 1, elementwise (`pcacc_road_marking`)."""
        return _to_like(intensity_map, _ops_cloud().road_marking(intensity_map, int_scaler, int_sep_scaler,
                                                                 int_mid_threshold))

    @staticmethod
    def sigmoid(z):
        """sem_bev.py:615-617."""
        return _to_like(z, _ops_cloud().road_marking(z, sigmoid_only=True))

    def get_elevation_map(self, pc):
        """sem_bev.py:535-553: per-cell minimum z of a cloud in grid coordinates -> (elevmap,
        elevmap_obs_mask) (`pcacc_elevation_map`; index rules of static_obj_partitioning_by_elev)."""
        ops = _ops_cloud()
        elev, obs = ops.elevation_map(pc, self.pixel_size)
        if ops.sync() & _lib.FLAG_ATTR_RANGE:
            raise IndexError('grid index out of bounds for the elevation map (or NaN z)')
        return _to_like(pc, elev), _to_like(pc, obs)
