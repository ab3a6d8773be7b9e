"""TEST INFRASTRUCTURE — CPU restatement of the reference's hot path.

This module is the *checker*: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.  It is
never on the product path (the product fails loudly without its CUDA library).

It restates, in vectorised numpy plus a small C helper for exact FMA chains
(oracle/fmachain.c), what robin-karlsson0/pc-accumulation-lib computes on the
path  project -> mask -> gather -> class filter -> pose transform -> accumulate
-> BEV preprocess -> per-cell reductions -> finalise.  Each function cites the
reference file:line it follows (paths relative to /root/reference).

Pinning: the reference ships no tests or golden vectors ("parity unpinned" at
the source).  This restatement is therefore pinned against
outputs of the reference itself, run in the build container under the stub
recipe of oracle/ref_loader.py: `tests/golden/*.npz` (made by
tests/golden/make_golden.py) and oracle/validate_against_reference.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, 'liboracle_fma.so')
        src = os.path.join(_HERE, 'fmachain.c')
        if (not os.path.exists(so)
                or os.path.getmtime(so) < os.path.getmtime(src)):
            subprocess.check_call(['make', '-C', _HERE, '-s'])
        _LIB = ctypes.CDLL(so)
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------------------
#  Exact small-matrix transforms (FMA chains)
# ---------------------------------------------------------------------------
def affine(M: np.ndarray, pts: np.ndarray) -> np.ndarray:
    """rows 0..2 of  M @ [pts 1]^T  as (N,3) f64, FMA chain over k=0..3.
    Equals np.matmul for N >= 2 on FMA-capable x86 hosts (checked in tests)."""
    M = np.ascontiguousarray(M, dtype=np.float64)
    assert M.shape in ((3, 4), (4, 4))
    n = pts.shape[0]
    out = np.empty((n, 3), dtype=np.float64)
    if n == 0:
        return out
    if pts.dtype == np.float32:
        p = np.ascontiguousarray(pts[:, :3])
        _lib().orc_affine_f32(_ptr(M), ctypes.c_int(4), _ptr(p),
                              ctypes.c_int64(n), ctypes.c_int64(3), _ptr(out))
    else:
        p = np.ascontiguousarray(pts[:, :3], dtype=np.float64)
        _lib().orc_affine_f64(_ptr(M), ctypes.c_int(4), _ptr(p),
                              ctypes.c_int64(n), ctypes.c_int64(3), _ptr(out))
    return out


def rot33(R: np.ndarray, pts: np.ndarray) -> np.ndarray:
    R = np.ascontiguousarray(R, dtype=np.float64)
    n = pts.shape[0]
    out = np.empty((n, 3), dtype=np.float64)
    if n == 0:
        return out
    p = np.ascontiguousarray(pts[:, :3], dtype=np.float64)
    _lib().orc_rot33_f64(_ptr(R), _ptr(p), ctypes.c_int64(n),
                         ctypes.c_int64(3), _ptr(out))
    return out


# ---------------------------------------------------------------------------
#  a1-a4: projection, mask, gather, class filter
# ---------------------------------------------------------------------------
def velo2frame(pc_xyz: np.ndarray, P_velo_frame: np.ndarray) -> np.ndarray:
    """sem_pc_accum.py:347-365 — (N,3) -> (N,3) f64 homogeneous image coords."""
    return affine(np.asarray(P_velo_frame, dtype=np.float64), pc_xyz)


def project(pc_velo: np.ndarray, P_velo_frame, img_h: int, img_w: int,
            max_depth=np.inf):
    """sem_pc_accum.py:382-396 — returns (u, v, mask) for ALL N points.
    u, v are int64 (np.round = half-to-even, then astype(int))."""
    frame = velo2frame(pc_velo[:, :3], P_velo_frame)
    depth = frame[:, 2].copy()
    depth[depth == 0] = -1e-6
    with np.errstate(invalid='ignore', over='ignore'):
        u = np.round(frame[:, 0] / np.abs(depth)).astype(int)
        v = np.round(frame[:, 1] / np.abs(depth)).astype(int)
    mask = (u >= 0) & (u < img_w) & (v >= 0) & (v < img_h)
    mask = mask & (depth > 0) & (depth < max_depth)
    return u, v, mask


def velo2img(pc_velo, P_velo_frame, img_h, img_w, max_depth=np.inf):
    """sem_pc_accum.py:367-402 — (M,6) f64 [x,y,z,i,u,v] of in-image points,
    original order."""
    u, v, mask = project(pc_velo, P_velo_frame, img_h, img_w, max_depth)
    out = np.concatenate([pc_velo, u[:, None], v[:, None]], axis=1)
    return out[mask]


def gen_semantic_pc(pc_velo, semantic_map, P_velo_frame):
    """sem_pc_accum.py:323-345 — (M,4+K) [x,y,z,i,feat_1..K]."""
    img_h, img_w, _ = semantic_map.shape
    pc_img = velo2img(pc_velo, P_velo_frame, img_h, img_w)
    u = pc_img[:, -2].astype(int)
    v = pc_img[:, -1].astype(int)
    return np.concatenate([pc_img[:, :4], semantic_map[v, u, :]], axis=1)


def filter_semseg_pc(pc, filters):
    """sem_pc_accum.py:317-321."""
    keep = np.ones(pc.shape[0], dtype=bool)
    for f in filters:
        keep &= pc[:, -1] != f
    return pc[keep]


def kitti_obs2sem(pc, rgb, class_map, P_velo_frame, filters, sem_gt=None):
    """kitti360_sem_pc_accum.py:129-156 without ICP — (M',10) f64 record
    [x,y,z,i,r,g,b,sem,inst=0,dyn=0]."""
    if sem_gt is None:
        pc_rgb = gen_semantic_pc(pc, np.asarray(rgb), P_velo_frame)
        pc_sem = gen_semantic_pc(pc, class_map[..., None], P_velo_frame)
        rec = np.concatenate([pc_rgb, pc_sem[:, -1:]], axis=1)
    else:
        n = sem_gt.shape[0]
        rec = np.concatenate([pc, np.zeros((n, 3)), sem_gt[:, -1:]], axis=1)
    rec = filter_semseg_pc(rec, filters)
    z = np.zeros((rec.shape[0], 2))
    return np.concatenate([rec, z], axis=1)


# ---------------------------------------------------------------------------
#  a9-a11: nuScenes nearest gather, homogeneous transform, record build
# ---------------------------------------------------------------------------
def homo_transform(tf, points):
    """datasets/nuscenes_utils.py:46-60."""
    assert tf.shape == (4, 4), f"{tf.shape} is not (4, 4)"
    assert points.shape == (points.shape[0], 3), \
        f"{points.shape} is not (N, 3)"
    if points.shape[0] == 1:
        # single-point products take numpy's matrix-vector path, whose
        # arithmetic differs from the FMA chain (DESIGN.md "N=1"): mirror it
        _p = np.concatenate([points, np.ones((1, 1))], axis=1)
        return (tf @ _p.T)[:3, :].T
    return affine(tf, points)


def pts_feat_nearest(pts_uv, img):
    """datasets/nuscenes_utils.py:181-214 with method='nearest'."""
    img_wh = np.array([img.shape[1], img.shape[0]], dtype=float)
    inside = (pts_uv > 1) & (pts_uv < img_wh - 1)
    assert np.all(inside), "pts_uv must be all inside image"
    uv = np.round(pts_uv).astype(int)
    return img[uv[:, 1], uv[:, 0]]


def nusc_obs2sem(pc, pc_cam_idx, rgbs, class_maps, T_ego_world, filters):
    """nuscenes_oracle_sem_pc_accum.py:454-501 — (M,10) f64 world-frame
    records; also returns the boolean keep mask over the N inputs."""
    feat = -np.ones((pc.shape[0], 4), dtype=float)
    for cam_idx, (rgb, cls) in enumerate(zip(rgbs, class_maps)):
        m = pc_cam_idx == cam_idx
        img = np.concatenate([np.asarray(rgb), cls[..., None]], axis=2)
        feat[m] = pts_feat_nearest(pc[m, 4:6], img)
    invalid = np.any(feat < 0, axis=1)
    for c in filters:
        invalid |= feat[:, -1] == c
    keep = ~invalid
    pc_k, feat_k = pc[keep], feat[keep]
    xyz = homo_transform(T_ego_world, pc_k[:, :3])
    rec = np.concatenate([xyz, pc_k[:, 3:4] / 255., feat_k, pc_k[:, 6:7],
                          np.zeros((pc_k.shape[0], 1))], axis=1)
    return rec, keep


# ---------------------------------------------------------------------------
#  the dataloader's input side (multi-camera projection,
#  box -> point instance assignment)
# ---------------------------------------------------------------------------
def apply_tf(tf, points):
    """datasets/nuscenes_utils.py:233-243: ([xyz 1] @ tf.T)[:, :3]; float32 points are
    promoted to float64 by the product.  Same FMA chain as homo_transform [measured]."""
    assert points.shape[1] >= 3 and tf.shape == (4, 4)
    if points.shape[0] == 1:
        xyz1 = np.pad(points[:, :3], pad_width=[(0, 0), (0, 1)], constant_values=1.0)
        return (xyz1 @ tf.T)[:, :3]
    return affine(tf, points[:, :3])


def find_points_in_box(points, target_from_box, dxdydz, tolerance):
    """datasets/nuscenes_utils.py:317-329."""
    box_points = apply_tf(np.linalg.inv(target_from_box), points[:, :3])
    return np.all(np.abs(box_points / np.asarray(dxdydz, dtype=float)) < (0.5 + tolerance), axis=1)


def assign_boxes(points, target_from_boxes, sizes, tolerance):
    """The box loop of inst_centric_get_sweeps (datasets/nuscenes_utils.py:412-470) reduced
    to what it does to the points: boxes are visited in order and every point inside takes
    the box's index, later boxes overwriting earlier ones.  Returns (box index per point, -1
    if none; number of points inside each box — the loop skips boxes with none)."""
    box = -np.ones(points.shape[0], dtype=np.int32)
    cnt = np.zeros(len(target_from_boxes), dtype=np.int32)
    for b, (tfb, size) in enumerate(zip(target_from_boxes, sizes)):
        m = find_points_in_box(points, tfb, size, tolerance)
        cnt[b] = int(m.sum())
        box[m] = b
    return box, cnt


def view_points(points, view, normalize):
    """nuscenes-devkit (version unpinned by the reference, README.md:29)
    nuscenes/utils/geometry_utils.py `view_points`, restated from its published source:
    pad `view` into a 4x4 identity, multiply the homogeneous points, keep three rows and,
    when `normalize`, divide every row by the third."""
    assert view.shape[0] <= 4 and view.shape[1] <= 4 and points.shape[0] == 3
    n = points.shape[1]
    if n == 1:
        viewpad = np.eye(4)
        viewpad[:view.shape[0], :view.shape[1]] = view
        pts = np.dot(viewpad, np.concatenate((points, np.ones((1, n)))))[:3, :]
    else:
        viewpad = np.eye(4)
        viewpad[:view.shape[0], :view.shape[1]] = view
        pts = affine(viewpad, np.ascontiguousarray(points.T)).T
    if normalize:
        pts = pts / pts[2:3, :].repeat(3, 0).reshape(3, n)
    return pts


def project_pts3d(pc, cam_K, img_wh, depth_thres=1e-3):
    """NuScenesCamera.project_pts3d, datasets/nuscenes_utils.py:112-136."""
    mask_valid = pc[:, 2] > depth_thres
    out = np.zeros((pc.shape[0], 2), dtype=float) - 10
    uv = view_points(np.ascontiguousarray(pc[mask_valid].T), cam_K, normalize=True)
    out[mask_valid] = uv[:2, :].T
    mask_in_img = (out > 1) & (out < np.asarray(img_wh, dtype=float) - 1)
    return out, np.all(mask_in_img, axis=1) & mask_valid


def project_to_cameras(pc_in_ego, glob_from_ego, cams, depth_thres=1e-3):
    """obs_dataloaders/nuscenes_obs_dataloader.py:176-198: every camera in turn, a later
    camera overwriting an earlier one.  cams: list of dicts glob_from_self (4,4), cam_K
    (3,3), img_wh (2,).  Returns (pc_uv (N,2) float64, pc_cam_idx (N,) int)."""
    pc_in_glob = homo_transform(glob_from_ego, pc_in_ego)
    pc_uv = np.zeros((pc_in_ego.shape[0], 2), dtype=float)
    pc_cam_idx = -np.ones(pc_in_ego.shape[0], dtype=int)
    for j, cam in enumerate(cams):
        pc_in_cam = homo_transform(np.linalg.inv(cam['glob_from_self']), pc_in_glob)
        uv, mask = project_pts3d(pc_in_cam, cam['cam_K'], cam['img_wh'], depth_thres)
        pc_uv[mask] = uv[mask]
        pc_cam_idx[mask] = j
    return pc_uv, pc_cam_idx


def pts_feat_from_img(pts_uv, img, method='bilinear'):
    """datasets/nuscenes_utils.py:181-214, both branches.  Bilinear: the reference's expression
    per channel (it only broadcasts for a 2-D image; there this IS its arithmetic)."""
    img_wh = np.array([img.shape[1], img.shape[0]], dtype=float)
    assert np.all((pts_uv > 1) & (pts_uv < img_wh - 1)), "pts_uv must be all inside image"
    if method == 'nearest':
        uv_ = np.round(pts_uv).astype(int)
        return img[uv_[:, 1], uv_[:, 0]]
    u, v = pts_uv[:, 0], pts_uv[:, 1]
    u_floor, u_ceil = np.floor(u), np.ceil(u)
    v_floor, v_ceil = np.floor(v), np.ceil(v)
    with np.errstate(divide='ignore', invalid='ignore'):
        total = (u_ceil - u_floor) * (v_ceil - v_floor)
        w_ff = (u_ceil - u) * (v_ceil - v) / total
        w_cc = (u - u_floor) * (v - v_floor) / total
        w_fc = (u - u_floor) * (v_ceil - v) / total
        w_cf = 1. - (w_ff + w_cc + w_fc)
        uf, vf, uc, vc = u_floor.astype(int), v_floor.astype(int), u_ceil.astype(int), v_ceil.astype(int)
        if img.ndim == 2:
            return w_ff * img[vf, uf] + w_cc * img[vc, uc] + w_cf * img[vc, uf] + w_fc * img[vf, uc]
        return (w_ff[:, None] * img[vf, uf] + w_cc[:, None] * img[vc, uc] + w_cf[:, None] * img[vc, uf]
                + w_fc[:, None] * img[vf, uc])


def static_obj_partitioning_by_elev(pc, P, elev_thresh):
    """bev_generator/sem_bev.py:556-591, vectorised: per-cell minimum z (row P-1-j, column i,
    indices truncated like .astype(int) and wrapped once like numpy's negative indices), then
    column 8 = 1 where z > min + thresh.  Mutates pc like the reference."""
    i = pc[:, 0].astype(int)
    j_rev = P - 1 - pc[:, 1].astype(int)
    if np.any((i < -P) | (i >= P) | (j_rev < -P) | (j_rev >= P)):
        raise IndexError('index out of bounds')
    i = np.where(i < 0, i + P, i)
    j_rev = np.where(j_rev < 0, j_rev + P, j_rev)
    cell = j_rev * P + i
    zmin = np.full(P * P, np.inf)
    np.minimum.at(zmin, cell, pc[:, 2])
    obs = np.zeros(P * P, dtype=bool)
    obs[cell] = True
    elevmap = np.where(obs, zmin, 0.).reshape(P, P)
    pc[pc[:, 2] > zmin[cell] + elev_thresh, 8] = 1
    return pc[pc[:, 8] == 0], pc[pc[:, 8] == 1], elevmap, obs.reshape(P, P)


# ---------------------------------------------------------------------------
#  a14-a20: the BEV generator
# ---------------------------------------------------------------------------
def rotation_matrix_3d(ang):
    """bev_generator/bev_generator.py:732-735."""
    return np.array([[np.cos(ang), -np.sin(ang), 0],
                     [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])


def heading_rot_ang(ego_traj_present):
    """bev_generator/bev_generator.py:87-93."""
    rot_ang = 0.5 * np.pi
    if len(ego_traj_present) > 1:
        dx = ego_traj_present[-1][0] - ego_traj_present[-2][0]
        dy = ego_traj_present[-1][1] - ego_traj_present[-2][1]
        rot_ang += np.arctan2(dy, dx)
    return np.pi - rot_ang


def point_in_box(x, y, bx0, by0, bx1, by1):
    """bev_generator/bev_generator.py:317-320."""
    return (bx0 < x and x < bx1) and (by0 < y and y < by1)


def bisect_boundary(x0, y0, x1, y1, bbox, thresh=1e-4):
    """bev_generator/bev_generator.py:322-371 — midpoint refinement until the
    replaced end moved by <= thresh."""
    bx0, by0, bx1, by1 = bbox
    diff = np.inf
    while diff > thresh:
        xm = 0.5 * (x0 + x1)
        ym = 0.5 * (y0 + y1)
        p0_in = point_in_box(x0, y0, bx0, by0, bx1, by1)
        mid_in = point_in_box(xm, ym, bx0, by0, bx1, by1)
        if mid_in == p0_in:
            diff = np.sqrt((xm - x0) ** 2 + (ym - y0) ** 2)
            x0, y0 = xm, ym
        else:
            diff = np.sqrt((xm - x1) ** 2 + (ym - y1) ** 2)
            x1, y1 = xm, ym
    return xm, ym


def crop_trajectory(traj, view, thresh=1e-4):
    """bev_generator/bev_generator.py:257-315."""
    b = (-0.5 * view, -0.5 * view, 0.5 * view, 0.5 * view)
    out = []
    for k in range(traj.shape[0] - 1):
        x0, y0 = list(traj[k][:2])
        x1, y1 = list(traj[k + 1][:2])
        z0 = traj[k][2]
        in0 = point_in_box(x0, y0, *b)
        in1 = point_in_box(x1, y1, *b)
        if not in0 and not in1:
            continue
        if in0:
            out.append([x0, y0, z0])
        if in0 != in1:
            xi, yi = bisect_boundary(x0, y0, x1, y1, list(b), thresh)
            out.append([xi, yi, z0])
    return np.array(out) if out else np.zeros((0, 3))


def pos2grid(mat, view, P):
    """bev_generator/bev_generator.py:737-747 (in place on columns 0:2)."""
    mat[:, 0:2] = np.floor(mat[:, 0:2] / view * P + 0.5 * P)
    return mat


def transform_traj(traj, R, dx, dy, view, P):
    """geometric_transform(is_traj=True) + pos2grid,
    bev_generator/bev_generator.py:141-158,224-237."""
    t = np.array(traj, dtype=float)
    t = t.reshape(-1, 3) if t.size else np.zeros((0, 3))
    t[:, :3] = np.matmul(R, t[:, :3].T).T
    t[:, 0] += dx
    t[:, 1] += dy
    t = crop_trajectory(t, view)
    return pos2grid(t, view, P)


def preprocess_pc(pc, R, dx, dy, view, P, height_filter):
    """bev_generator/bev_generator.py:127-160,207-255 for the point cloud:
    rotate (FMA chain), translate, strict crop, height filter, pos2grid.
    Returns the (Mc,10) cloud with integer-valued xy."""
    pc = pc.copy()
    pc[:, :3] = rot33(R, pc[:, :3]) if pc.shape[0] != 1 else \
        np.matmul(R, pc[:, :3].T).T
    pc[:, 0] += dx
    pc[:, 1] += dy
    m = (pc[:, 0] > -0.5 * view) & (pc[:, 0] < 0.5 * view)
    pc = pc[m]
    m = (pc[:, 1] > -0.5 * view) & (pc[:, 1] < 0.5 * view)
    pc = pc[m]
    if height_filter is not None:
        pc = pc[pc[:, 2] < height_filter]
    return pos2grid(pc, view, P)


def _segment_median(cell, val, P2):
    """Per-cell np.median (mean of the two middle values for even counts),
    bev_generator/sem_bev.py:657-667; NaN where the cell is empty."""
    out = np.full(P2, np.nan)
    if cell.size == 0:
        return out
    order = np.lexsort((val, cell))
    c = cell[order]
    v = val[order]
    starts = np.flatnonzero(np.r_[True, c[1:] != c[:-1]])
    counts = np.diff(np.r_[starts, c.size])
    hi = starts + counts // 2
    lo = np.where(counts % 2 == 1, hi, hi - 1)
    out[c[starts]] = (v[lo] + v[hi]) / 2.0
    return out


def raster_window(pc_grid, P, sem_idxs, int_scaler, int_sep_scaler,
                  int_mid_threshold, rgb_fill=0, return_f64=False,
                  elevation_mode='min'):
    """One window of SemBEVGenerator.generate_bev
    (bev_generator/sem_bev.py:54-118,196-257) on a preprocessed cloud:
    7 planes (road, intensity, r, g, b, dynamic, elevation), each (P,P).

    `pc_grid` columns: [i, j, z, intensity, r, g, b, sem, inst, dyn] with
    integer-valued i, j (output of preprocess_pc)."""
    P2 = P * P
    pc = pc_grid[pc_grid[:, 9] != 1]                       # sem_bev.py:54-58
    col = pc[:, 0].astype(np.int64)
    row = P - 1 - pc[:, 1].astype(np.int64)                # flip, :453 / :546
    ok = (col >= 0) & (col < P) & (row >= 0) & (row < P)
    pc, col, row = pc[ok], col[ok], row[ok]
    cell = row * P + col
    sem = pc[:, 7]

    n = np.bincount(cell, minlength=P2).astype(np.float64)
    road = sem == sem_idxs['road']
    veh = np.zeros(sem.shape, dtype=bool)
    for name in ('car', 'truck', 'bus', 'motorcycle'):     # sem_bev.py:55
        veh |= sem == sem_idxs[name]
    n_road = np.bincount(cell[road], minlength=P2).astype(np.float64)
    n_veh = np.bincount(cell[veh], minlength=P2).astype(np.float64)

    # Dirichlet expectation with a uniform prior, bev_generator.py:457-480
    p_road = (n_road + 1.) / ((n_road + 1.) + ((n - n_road) + 1.))
    p_veh = (n_veh + 1.) / ((n_veh + 1.) + ((n - n_veh) + 1.))

    # intensity: sequential f64 sum in point order, bev_generator.py:396-415
    int_sum = np.bincount(cell[road], weights=pc[road, 3], minlength=P2)
    inten = int_sum / (n_road + 1.)
    # road_marking_transform, sem_bev.py:593-617
    inten = int_scaler * (1 / (1 + np.exp(-(int_sep_scaler *
                                            (inten - int_mid_threshold)))))
    inten[inten > 1.] = 1.

    # elevation: per-cell min z, 0 where unobserved, sem_bev.py:535-554
    elev = np.zeros(P2)
    if cell.size:
        if elevation_mode == 'min':
            ext = np.full(P2, np.inf)
            np.minimum.at(ext, cell, pc[:, 2])
        else:                       # north-star variant, not the reference
            ext = np.full(P2, -np.inf)
            np.maximum.at(ext, cell, pc[:, 2])
        obs = n > 0
        elev[obs] = ext[obs]

    # rgb: per-cell per-channel median, fill where empty, /255
    planes = [p_road, inten]
    for ch in (4, 5, 6):
        med = _segment_median(cell, pc[:, ch], P2)
        med[np.isnan(med)] = rgb_fill
        planes.append(med / 255.)
    planes += [p_veh, elev]
    out = np.stack([p.reshape(P, P) for p in planes])
    return out if return_f64 else out.astype(np.float16)


PLANES = ('road', 'intensity', 'r', 'g', 'b', 'dynamic', 'elevation')


# ---------------------------------------------------------------------------
#  the per-plane steps of generate_bev as the stand-alone methods the reference exposes
# ---------------------------------------------------------------------------
def crop_view(pc, view):
    """bev_generator/bev_generator.py:239-256: strict crop of x, then of y."""
    pc = pc[(pc[:, 0] > -0.5 * view) & (pc[:, 0] < 0.5 * view)]
    return pc[(pc[:, 1] > -0.5 * view) & (pc[:, 1] < 0.5 * view)]


def geometric_transform_pc(pc, R, dx, dy, view):
    """bev_generator/bev_generator.py:207-237, cloud branch (on a copy)."""
    pc = pc.copy()
    pc[:, :3] = rot33(R, pc[:, :3]) if pc.shape[0] != 1 else np.matmul(R, pc[:, :3].T).T
    pc[:, 0] += dx
    pc[:, 1] += dy
    return crop_view(pc, view)


def partition_semantic_pc(pc, sems, sem_col):
    """bev_generator/bev_generator.py:411-432."""
    hit = np.isin(pc[:, sem_col], np.asarray(sems, dtype=np.float64))
    return pc[hit], pc[~hit]


def gridmap_count_map(pc, P, weights=None):
    """bev_generator/bev_generator.py:434-453 without np.histogram2d: P unit bins over [0, P] a side
    (the right edge belongs to the last bin, everything else outside and NaN is dropped), first
    axis = column 1 flipped, second axis = column 0; weights are summed in point order."""
    def bins(v):
        ok = (v >= 0) & (v <= P)
        b = np.where(ok, np.minimum(np.nan_to_num(v, nan=0.), P - 1), 0).astype(np.int64)
        return ok, b
    ok_j, bj = bins(pc[:, 1])
    ok_i, bi = bins(pc[:, 0])
    ok = ok_j & ok_i
    cell = ((P - 1 - bj) * P + bi)[ok]
    w = None if weights is None else np.asarray(weights, dtype=np.float64)[ok]
    return np.bincount(cell, weights=w, minlength=P * P).astype(np.float64).reshape(P, P)


def dirichlet_expectation(gridmaps, obs_weight=1):
    """bev_generator/bev_generator.py:455-481."""
    g = np.stack(gridmaps).astype(np.float64) * obs_weight + 1.
    a0 = g[0].copy()
    for k in range(1, g.shape[0]):
        a0 = a0 + g[k]
    return [g[k] / a0 for k in range(g.shape[0])]


def sem_probmap(pc, P, sems, sem_col=7):
    """bev_generator/bev_generator.py:373-391."""
    a, b = partition_semantic_pc(pc, sems, sem_col)
    return dirichlet_expectation([gridmap_count_map(a, P), gridmap_count_map(b, P)])[0]


def intensity_map(pc, P, sem, sem_col=7):
    """bev_generator/bev_generator.py:393-409."""
    a, _ = partition_semantic_pc(pc, [sem], sem_col)
    return gridmap_count_map(a, P, weights=a[:, 3]) / (gridmap_count_map(a, P) + 1)


def sigmoid(z):
    """bev_generator/sem_bev.py:615-617."""
    return 1 / (1 + np.exp(-np.asarray(z, dtype=np.float64)))


def road_marking_transform(inten, int_scaler, int_sep_scaler, int_mid_threshold):
    """bev_generator/sem_bev.py:593-613."""
    out = int_scaler * sigmoid(int_sep_scaler * (np.asarray(inten, dtype=np.float64) - int_mid_threshold))
    out[out > 1.] = 1.
    return out


def elevation_map(pc, P):
    """bev_generator/sem_bev.py:535-553 (per-cell minimum z, observed mask)."""
    q = np.zeros((pc.shape[0], 10))
    q[:, :3] = pc[:, :3]
    _, _, elev, obs = static_obj_partitioning_by_elev(q, P, np.inf)
    return elev, obs


def rgb_maps(pc, P, rgb_fill=0):
    """bev_generator/sem_bev.py:619-669 == rgb_bev.py:133-183: per-cell medians of columns 4..6."""
    i = pc[:, 0].astype(int)
    j_rev = P - 1 - pc[:, 1].astype(int)
    if np.any((i < -P) | (i >= P) | (j_rev < -P) | (j_rev >= P)):
        raise IndexError('index out of bounds')
    cell = np.where(j_rev < 0, j_rev + P, j_rev) * P + np.where(i < 0, i + P, i)
    out = []
    for ch in (4, 5, 6):
        med = _segment_median(cell, pc[:, ch], P * P)
        med[np.isnan(med)] = rgb_fill
        out.append(med.reshape(P, P))
    return tuple(out)


# ---------------------------------------------------------------------------
#  polynomial warp (bev_generator/bev_generator.py:482-698), SURVEY §8f rank 1
# ---------------------------------------------------------------------------
def cal_warp_params(idx_0, idx_1, idx_max):
    """bev_generator.py:663-683."""
    a_1 = (idx_1 - idx_0 ** 2 / idx_max) / (idx_0 * (1.0 - idx_0 / idx_max))
    a_2 = (1.0 - a_1) / idx_max
    return a_1, a_2


def random_warp_params(np_rng, py_rng, mean_ratio, max_ratio, I, J):
    """get_random_warp_params, bev_generator.py:621-661: two np.random.normal draws, then
    two random.random() draws for the signs."""
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
    return int(I / 2) + i_warp, int(J / 2) + j_warp


def warp_index_map(c_1, c_2, n):
    """Source index of every warped index: clamp(rint(c_1 k + c_2 k^2)), bev_generator.py:508-520."""
    k = np.arange(n)
    src = np.rint(c_1 * k + c_2 * k ** 2).astype(np.int64)
    return np.clip(src, 0, n - 1)


def warp_dense(maps, a_1, a_2, b_1, b_2):
    """warp_dense_probmaps, bev_generator.py:482-525: B[:, jw, iw] = A[:, j(jw), i(iw)]."""
    n, I, J = maps.shape
    imap = warp_index_map(a_1, a_2, I)
    jmap = warp_index_map(b_1, b_2, J)
    return maps[:, jmap[:, None], imap[None, :]]


def warp_traj(traj, a_1, a_2, j_warp, P):
    """warp_sparse_points + warp_point, bev_generator.py:527-590 (the j warp is reversed)."""
    import math
    b_1, b_2 = cal_warp_params(P - j_warp, int(P / 2), P - 1)
    out = np.array(traj, dtype=float).reshape(-1, 3)

    def one(v, c_1, c_2):
        if math.isclose(c_2, 0.0, abs_tol=1e-6):
            w = v
        else:
            with np.errstate(invalid='ignore'):
                w = int(np.rint((-c_1 + np.sqrt(c_1 ** 2 + 4.0 * c_2 * v)) / (2 * c_2)))
        if w < 0:
            w = 0
        elif w >= P:
            w = P - 1
        return w

    for k in range(out.shape[0]):
        out[k, 0] = one(out[k, 0], a_1, a_2)
        out[k, 1] = one(out[k, 1], b_1, b_2)
    return out


def generate(pcs, trajs, gen_params, rot_ang=0., trans_dx=0., trans_dy=0.,
             zoom_scalar=1., do_warping=False, return_f64=False, warp_rngs=None):
    """BEVGenerator.generate + SemBEVGenerator.generate_bev
    (bev_generator/bev_generator.py:63-125, sem_bev.py:36-262), warp excluded.

    gen_params: dict(sem_idxs, view_size, pixel_size, int_scaler,
                     int_sep_scaler, int_mid_threshold, height_filter,
                     rgb_fill)
    Returns the 18-key dict; with return_f64 also '<name>_f64' stacks
    (3 windows x 7 planes before the fp16 cast) and 'cells_<w>' index lists.
    """
    P = gen_params['pixel_size']
    view = zoom_scalar * gen_params['view_size']
    ego_present = trajs['ego_traj_present']
    if do_warping is False:
        rot_ang = heading_rot_ang(ego_present)
    R = rotation_matrix_3d(rot_ang)

    bev = {}
    dbg = {}
    warp = None
    if gen_params.get('do_warp'):
        # sem_bev.py:121-190: one set of warp parameters for all maps and trajectories
        i_mid = int(P / 2)
        i_warp, j_warp = random_warp_params(warp_rngs[0], warp_rngs[1], 0.15, 0.30, P, P)
        a_1, a_2 = cal_warp_params(i_warp, i_mid, P - 1)
        b_1, b_2 = cal_warp_params(j_warp, i_mid, P - 1)
        warp = (a_1, a_2, b_1, b_2, j_warp)
    for w in ('present', 'future', 'full'):
        pc = pcs[f'pc_{w}']
        if pc is None:
            continue
        pcg = preprocess_pc(pc, R, trans_dx, trans_dy, view, P,
                            gen_params.get('height_filter'))
        tr = [trajs[f'ego_traj_{w}']] + list(trajs[f'other_trajs_{w}'])
        tr = [transform_traj(t, R, trans_dx, trans_dy, view, P) for t in tr]
        planes = raster_window(
            pcg, P, gen_params['sem_idxs'], gen_params['int_scaler'],
            gen_params['int_sep_scaler'], gen_params['int_mid_threshold'],
            gen_params.get('rgb_fill', 0), return_f64=True,
            elevation_mode=gen_params.get('elevation_mode', 'min'))
        if warp is not None:
            planes = warp_dense(planes, *warp[:4])
            tr = [warp_traj(t, warp[0], warp[1], warp[4], P) for t in tr]
        h = planes.astype(np.float16)
        bev[f'road_{w}'] = h[0]
        bev[f'trajs_{w}'] = tr
        bev[f'intensity_{w}'] = h[1]
        bev[f'rgb_{w}'] = h[2:5]
        bev[f'dynamic_{w}'] = h[5]
        bev[f'elevation_{w}'] = h[6]
        if return_f64:
            dbg[f'planes_f64_{w}'] = planes
            st = pcg[pcg[:, 9] != 1]
            dbg[f'cells_{w}'] = np.stack(
                [P - 1 - st[:, 1].astype(np.int64), st[:, 0].astype(np.int64)],
                axis=1)
    if 'gt_lanes' in trajs:
        lanes = [transform_traj(t, R, trans_dx, trans_dy, view, P)
                 for t in trajs['gt_lanes']]
        lanes = [l for l in lanes if l.shape[0] > 0]
        if warp is not None:
            lanes = [warp_traj(t, warp[0], warp[1], warp[4], P) for t in lanes]
        bev['gt_lanes'] = lanes
    if return_f64:
        bev['_debug'] = dbg
    return bev


def split_windows(sem_pcs, poses, present_idx, other_trajs=None):
    """generate_bev of the accumulators (kitti360_sem_pc_accum.py:179-228,
    nuscenes_oracle_sem_pc_accum.py:521-595): window split + origin shift."""
    origin = np.array(poses[-1] if present_idx is None else poses[present_idx])
    pcs, trajs = {}, {}
    if other_trajs is None:
        other_trajs = ([], [], [])
    for w, sl, ot in (('present', slice(None, present_idx), other_trajs[0]),
                      ('future', slice(present_idx, None), other_trajs[1]),
                      ('full', slice(None), other_trajs[2])):
        pc = np.concatenate(sem_pcs[sl])
        pc[:, :3] = pc[:, :3] - origin
        pcs[f'pc_{w}'] = pc
        trajs[f'ego_traj_{w}'] = np.concatenate([poses[sl]]) - origin
        trajs[f'other_trajs_{w}'] = [np.concatenate([t]) - origin for t in ot]
    return pcs, trajs


# ---------------------------------------------------------------------------
#  a5-a8, a13: KITTI-360 accumulator (ICP replaced by an injected transform)
# ---------------------------------------------------------------------------
class KittiOracle:
    """Restates Kitti360SemanticPointCloudAccumulator
    (kitti360_sem_pc_accum.py:41-243 over sem_pc_accum.py:156-228) with the
    ICP result `T_new_prev` injected per frame."""

    def __init__(self, horizon_dist, P_velo_frame, filters, gen_params,
                 use_gt_sem=False):
        self.horizon_dist = horizon_dist
        self.P = np.asarray(P_velo_frame, dtype=np.float64)
        self.filters = list(filters)
        self.gen_params = gen_params
        self.use_gt_sem = use_gt_sem
        self.sem_pcs, self.poses, self.seg_dists = [], [], []

    def integrate(self, pc, rgb, class_map, T_new_prev, sem_gt=None):
        rec = kitti_obs2sem(pc, rgb, class_map, self.P, self.filters,
                            sem_gt if self.use_gt_sem else None)
        if len(self.poses) > 0:
            # update_poses, sem_pc_accum.py:156-165 (one 4x1 product per pose)
            self.poses = [
                list(np.matmul(T_new_prev, np.array([p + [1]]).T)[:, 0][:-1])
                for p in self.poses]
            # update_sem_pcs, sem_pc_accum.py:167-183
            for s in self.sem_pcs:
                if s.shape[0] == 0:
                    continue
                if s.shape[0] == 1:
                    h = np.concatenate((s[:, :3], np.ones((1, 1))), axis=1)
                    s[:, :3] = np.matmul(T_new_prev, h.T).T[:, :3]
                else:
                    s[:, :3] = affine(T_new_prev, s[:, :3])
        self.sem_pcs.append(rec)
        self.poses.append([0., 0., 0.])
        idx = 0
        if len(self.poses) > 1:
            idx = self._remove_observations()
        return idx

    def _remove_observations(self):
        """sem_pc_accum.py:185-209."""
        idx = 0
        d = np.sqrt(np.sum((np.array(self.poses[-2])
                            - np.array(self.poses[-1])) ** 2))
        self.seg_dists.append(d)
        path_length = np.sum(self.seg_dists)
        if path_length > self.horizon_dist:
            incr = np.matmul(np.tri(len(self.seg_dists)),
                             np.array(self.seg_dists))
            incr -= path_length - self.horizon_dist
            idx = (incr > 0.).argmax()
            self.sem_pcs = self.sem_pcs[idx:]
            self.poses = self.poses[idx:]
            self.seg_dists = self.seg_dists[idx:]
        return idx

    def generate_bev(self, present_idx=None, return_f64=False, **aug):
        pcs, trajs = split_windows(self.sem_pcs, self.poses, present_idx)
        return generate(pcs, trajs, self.gen_params, return_f64=return_f64,
                        **aug)


# ---------------------------------------------------------------------------
#  a11-a13: nuScenes oracle-pose accumulator
# ---------------------------------------------------------------------------
class NuscOracle:
    """Restates NuScenesOracleSemanticPointCloudAccumulator
    (nuscenes_oracle_sem_pc_accum.py:139-270,272-414,416-610)."""

    TRACKED = [0, 1, 2, 3, 5]
    DYN_THRESH = 1.0

    def __init__(self, filters, gen_params, ego_pose_z=1.):
        self.filters = list(filters)
        self.gen_params = gen_params
        self.ego_pose_z = ego_pose_z
        self.T_global_world = None
        self.sem_pcs, self.poses, self.seg_dists = [], [], []
        self.instances, self.dyn_instances, self.token2idx = {}, [], []
        self.ts = 0

    def integrate(self, obs, class_maps):
        T_ego_global = obs['ego_at_lidar_ts']
        if self.T_global_world is None:
            self.T_global_world = np.linalg.inv(T_ego_global)
        T_ego_world = self.T_global_world @ T_ego_global
        pose = T_ego_world[:3, -1].tolist()
        pose[2] += self.ego_pose_z
        rec, _ = nusc_obs2sem(obs['pc'], obs['pc_cam_idx'], obs['images'],
                              class_maps, T_ego_world, self.filters)
        self.sem_pcs.append(rec)
        self.poses.append(pose)

        # fake detector / tracker, :191-250
        self.token2idx.append({'ts': self.ts})
        for idx, token in enumerate(obs['inst_tokens']):
            if obs['inst_cls'][idx] not in self.TRACKED:
                continue
            c = homo_transform(self.T_global_world,
                               np.expand_dims(obs['inst_center'][idx], 0))[0]
            self.instances.setdefault(token, []).append((c, self.ts))
            self.token2idx[-1][token] = idx
            if token in self.dyn_instances:
                s = self.sem_pcs[-1]
                s[s[:, 8] == idx, 9] = 1
                continue
            obs_list = self.instances[token]
            if len(obs_list) < 2:
                continue
            d = obs_list[-1][0][:2] - obs_list[0][0][:2]
            if np.linalg.norm(d) > self.DYN_THRESH:
                self.dyn_instances.append(token)
                for ts, s in enumerate(self.sem_pcs):
                    if token in self.token2idx[ts]:
                        s[s[:, 8] == self.token2idx[ts][token], 9] = 1
        if len(self.poses) > 1:
            self.seg_dists.append(np.sqrt(np.sum(
                (np.array(self.poses[-1]) - np.array(self.poses[-2])) ** 2)))
        self.ts += 1

    # trajectory bookkeeping, :272-414
    def dyn_obj_trajs(self, ts_start=0, ts_end=None):
        out = []
        for token, obs_list in self.instances.items():
            if token not in self.dyn_instances:
                continue
            poses, tss = zip(*obs_list)
            ge = [k for k, t in enumerate(tss) if t >= ts_start]
            if not ge:
                continue
            i0 = ge[0]
            if ts_end is None:
                i1 = None
            else:
                if tss[0] > ts_end:
                    continue
                le = [k for k in range(len(tss) - 1) if tss[k + 1] > ts_end]
                i1 = (le[0] if le else len(tss) - 1) + 1
            poses, tss = poses[i0:i1], tss[i0:i1]
            if len(tss) == 0:
                # reference: parse_seq_into_coherent_seqs reads ts[0] ->
                # IndexError propagates; generators avoid empty splices
                raise IndexError('tuple index out of range')
            seqs, prev = [[]], tss[0] - 1
            for k, t in enumerate(tss):
                if t - prev != 1:
                    seqs.append([])
                seqs[-1].append(poses[k].tolist())
                prev = t
            out += [s for s in seqs if len(s) >= 2]
        return out

    def generate_bev(self, present_idx=None, return_f64=False, **aug):
        other = (self.dyn_obj_trajs(ts_end=present_idx),
                 self.dyn_obj_trajs(ts_start=present_idx),
                 self.dyn_obj_trajs())
        pcs, trajs = split_windows(self.sem_pcs, self.poses, present_idx,
                                   other)
        return generate(pcs, trajs, self.gen_params, return_f64=return_f64,
                        **aug)
