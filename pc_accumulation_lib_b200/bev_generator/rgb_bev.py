"""RGBBEVGenerator — signature and output keys of the reference's
`bev_generator/rgb_bev.py` (`rgb_present`, `rgb_future`, `poses_present`,
`poses_future`).  The reference accumulator never constructs it
(sem_pc_accum.py:120-121 raises), so only `generate_bev` on pre-processed
clouds is offered (with and without the polynomial warp); the per-cell median
raster is the same kernel as SemBEVGenerator's RGB planes (rgb_bev.py:133-183 ==
sem_bev.py:619-669), the warp is `pcacc_warp_planes`."""
from __future__ import annotations

import numpy as np

from .bev_generator import BEVGenerator


class RGBBEVGenerator(BEVGenerator):
    def __init__(self, view_size: int, pixel_size: int, rgb_fill: int = 0,
                 max_trans_radius: float = 0., zoom_thresh: float = 0., do_warp: bool = False):
        super().__init__(view_size, pixel_size, max_trans_radius, zoom_thresh, do_warp)
        self.rgb_fill = rgb_fill

    def _assemble(self, planes, v, trajs_by_window, gt_lane_trajs, has_future):
        raise NotImplementedError('RGBBEVGenerator only offers generate_bev()')

    def generate_bev(self, pc_present, pc_future, poses_present, poses_future, do_warping=False):
        """rgb_bev.py:27-95: per-cell RGB medians / 255 of two pre-processed clouds (grid coordinates
        in columns 0, 1; r, g, b in columns 4..6), optionally pushed through one polynomial warp
        together with the pose lists (drawn like the reference: two normal draws, two sign draws)."""
        P = self.pixel_size

        def centre(pc):
            pc = np.array(pc, dtype=np.float64)
            out = np.zeros((pc.shape[0], 10))
            out[:, :min(pc.shape[1], 7)] = pc[:, :7]      # x,y,z,i,r,g,b; sem/inst/dyn = 0
            out[:, 0] += 0.5 - 0.5 * P
            out[:, 1] += 0.5 - 0.5 * P
            return out

        pres, fut = centre(pc_present), centre(pc_future)
        pcs = {'pc_present': pres, 'pc_future': fut,
               'pc_full': np.zeros((0, 10))}
        saved = self.view_size, self.height_filter
        self.view_size, self.height_filter = float(P), None
        try:
            aug = [dict(rot_ang=0., trans_dx=0., trans_dy=0., zoom_scalar=1., do_warping=True)]
            warp = self.draw_warp() if do_warping else None
            planes, _ = self._rasterise_windows(pcs, aug, [warp] if warp else None)
        finally:
            self.view_size, self.height_filter = saved
        if warp is not None:
            w = warp
            poses_present = self.warp_sparse_points(poses_present, w['a_1'], w['a_2'], w['b_1'], w['b_2'],
                                                    w['i_mid'], w['j_mid'], w['i_warp'], w['j_warp'])
            poses_future = self.warp_sparse_points(poses_future, w['a_1'], w['a_2'], w['b_1'], w['b_2'],
                                                   w['i_mid'], w['j_mid'], w['i_warp'], w['j_warp'])
        return {'rgb_present': planes[0, 0, 2:5], 'rgb_future': planes[0, 1, 2:5],
                'poses_present': poses_present, 'poses_future': poses_future}
