"""Mirror of the parts of the reference's datasets/nuscenes_utils.py (and of the projection
loop of obs_dataloaders/nuscenes_obs_dataloader.py:176-198) that produce the inputs of
`NuScenesOracleSemanticPointCloudAccumulator.integrate`: the (N,7) rows `[x, y, z, intensity,
u, v, inst]` and `pc_cam_idx`.  Same function names and argument meaning; the per-point work
runs in libpcacc (`pcacc_assign_boxes`, `pcacc_project_cameras`), the 4x4 inverses and the
small bookkeeping stay in numpy exactly as the reference has them.

`pts_feat_from_img` (both branches) runs in `pcacc_pts_feat_from_img`; the accumulator's own
'nearest' gather is fused into `pcacc_integrate_records`.
"""
from __future__ import annotations

import numpy as np

from ..device import DeviceCloud

_cloud = None


def _dev() -> DeviceCloud:
    """A small handle that only provides the parameter arena / stream plumbing."""
    global _cloud
    if _cloud is None:
        _cloud = DeviceCloud(capacity_pts=1024, max_frames=4)
    return _cloud


def _affine_rows(mat, xyz):
    """Rows [x y z 1] @ mat.T restricted to the first three output columns, i.e. the 3-D part
    of a 4x4 homogeneous transform applied to every row of `xyz` (float64 matmul, the same
    dgemm the reference calls)."""
    rows = np.empty((xyz.shape[0], 4), dtype=np.result_type(xyz.dtype, np.float64))
    rows[:, :3] = xyz
    rows[:, 3] = 1.0
    return rows


def homo_transform(tf, points):
    """Behaviour of datasets/nuscenes_utils.py:46-60 (host; a handful of poses per
    observation): `(tf @ [p 1]^T)[:3]^T` with the reference's two shape assertions."""
    if tf.shape != (4, 4):
        raise AssertionError(f'{tf.shape} is not (4, 4)')
    if points.ndim != 2 or points.shape[1] != 3:
        raise AssertionError(f'{points.shape} is not (N, 3)')
    return np.matmul(tf, _affine_rows(tf, points).T)[:3].T


def apply_tf(tf, points, in_place=False):
    """Behaviour of datasets/nuscenes_utils.py:233-243 (host): `[p 1] @ tf^T`, written back
    into `points[:, :3]` when `in_place`, returned otherwise."""
    if points.shape[1] < 3:
        raise AssertionError(f'expect points.shape[1] >= 3, get {points.shape[1]}')
    if tf.shape != (4, 4):
        raise AssertionError(f'expect tf.shape == 4, get {tf.shape}')
    moved = np.matmul(_affine_rows(tf, points[:, :3]), tf.T)[:, :3]
    if not in_place:
        return moved
    points[:, :3] = moved


def pts_feat_from_img(pts_uv, img, method='bilinear'):
    """datasets/nuscenes_utils.py:181-214 -> (N, C) array ((N,) for a 2-D image).  'nearest':
    `img[round(v), round(u)]`; 'bilinear': the reference's four weights in its operation order.
    The reference multiplies (N,) weights with (N, C) features, which numpy only broadcasts for a
    2-D image — there the results are bit-equal (tests); for C channels every channel gets the
    per-point weights.  Coordinates outside `1 < uv < wh - 1` raise the reference's AssertionError."""
    if not isinstance(img, np.ndarray):
        raise AssertionError(f'{type(img)} is not supported')
    if method not in ('bilinear', 'nearest'):
        raise AssertionError(f'{method} is not supported')
    dev = _dev()
    out = dev.pts_feat_from_img(np.ascontiguousarray(pts_uv, dtype=np.float64), img, method)
    if dev.sync() & 1:          # PCACC_FLAG_UV_OUT_OF_IMAGE
        raise AssertionError('pts_uv must be all inside image')
    res = out.cpu().numpy()
    if method == 'nearest':
        return res.astype(img.dtype)      # a pure gather keeps the image's dtype
    return res


def find_points_in_box(points, target_from_box, dxdydz, tolerance):
    """datasets/nuscenes_utils.py:317-329 -> (N,) bool."""
    box, _ = _dev().assign_boxes(points, [target_from_box], [dxdydz], tolerance)
    return (box >= 0).cpu().numpy()


def assign_points_to_boxes(points, target_from_boxes, sizes, tolerance):
    """All boxes of a sweep in one launch: the effect of the loop at
    datasets/nuscenes_utils.py:412-470 on the points.  -> (index of the last box containing
    each point, -1 if none; points inside each box).  The caller skips boxes whose count is
    zero and maps box -> instance / class index as the reference does."""
    box, cnt = _dev().assign_boxes(points, target_from_boxes, sizes, tolerance)
    return box.cpu().numpy(), cnt.cpu().numpy()


def project_pts3d(pc, cam_K, img_wh, depth_thres=1e-3):
    """NuScenesCamera.project_pts3d, datasets/nuscenes_utils.py:112-136, for points already in
    the camera frame -> (uv (N,2) with -10 rows for invalid points, mask_in_img (N,) bool)."""
    cam = dict(glob_from_self=np.eye(4), cam_K=cam_K, img_wh=img_wh)
    pc = np.ascontiguousarray(pc, dtype=np.float64)
    uv, idx = _dev().project_cameras(pc, np.eye(4), [cam], depth_thres)
    mask = (idx >= 0).cpu().numpy()
    # the device reports uv only for points inside the image; the reference also returns the
    # off-image projections of valid points, which its only caller discards
    # (nuscenes_obs_dataloader.py:191-196)
    out = np.zeros((pc.shape[0], 2)) - 10.
    out[mask] = uv.cpu().numpy()[mask]
    return out, mask


def project_to_cameras(pc_in_ego, glob_from_ego, cams, depth_thres=1e-3):
    """obs_dataloaders/nuscenes_obs_dataloader.py:176-198 -> (pc_uv (N,2) float64,
    pc_cam_idx (N,) int64); cams: list of dicts glob_from_self (4,4), cam_K (3,3), img_wh."""
    uv, idx = _dev().project_cameras(pc_in_ego, glob_from_ego, cams, depth_thres)
    return uv.cpu().numpy(), idx.cpu().numpy()
