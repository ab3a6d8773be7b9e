"""SemanticPointCloudAccumulator — the reference's accumulator base class
(`sem_pc_accum.py`) with the accumulated cloud held in a device-resident SoA
ring (libpcacc) instead of a Python list of numpy arrays.

Same constructor arguments, attributes (`poses`, `seg_dists`, `rgbs`,
`semsegs`, `sem_pcs`) and helper names as the reference.  `sem_pcs` is a lazy
view: indexing it exports that frame from the device as the reference's
(M,10) float64 record array.

Out of scope, as in SURVEY.md §2: the ONNX segmentation network (pass any
object with `.pred(rgb) -> (1,1,H,W)` as `semseg_onnx_path`), Open3D ICP and
visualisation.
"""
from __future__ import annotations

import gzip
import os
import pickle
import time
from collections.abc import Sequence

import numpy as np
import torch

from . import _lib
from .bev_generator import DeviceWindow, SemBEVGenerator
from .device import DeviceCloud


class _LazyTrajs(dict):
    """The `trajs` dict of BEVGenerator.generate whose other-object / lane entries are computed
    on first use (`_fill`); the ego entries are there from the start."""
    _fill = None

    def _materialise(self):
        fill, self._fill = self._fill, None
        if fill is not None:
            self.update(fill())

    def __missing__(self, key):
        if self._fill is None:
            raise KeyError(key)
        self._materialise()
        return dict.__getitem__(self, key)

    def __contains__(self, key):
        if not dict.__contains__(self, key) and self._fill is not None:
            self._materialise()
        return dict.__contains__(self, key)

    def keys(self):
        self._materialise()
        return dict.keys(self)

    def items(self):
        self._materialise()
        return dict.items(self)

    def __iter__(self):
        self._materialise()
        return dict.__iter__(self)

    def __len__(self):
        self._materialise()
        return dict.__len__(self)


class SemPcsView(Sequence):
    """`accumulator.sem_pcs`: list-like view of the live frames."""

    def __init__(self, acc):
        self._acc = acc

    def __len__(self):
        return len(self._acc._fids)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._acc._export(f) for f in self._acc._fids[i]]
        return self._acc._export(self._acc._fids[i])


class SemanticPointCloudAccumulator:
    def __init__(self, horizon_dist: float, icp_threshold: float, semseg_onnx_path,
                 semseg_filters: list, sem_idxs: dict, use_gt_sem: bool, bev_params: dict,
                 ring_capacity_pts: int = 8_000_000, ring_max_frames: int = 2048,
                 device: int | None = None):
        self.semseg_model = None
        if use_gt_sem is False:
            if hasattr(semseg_onnx_path, 'pred'):
                self.semseg_model = semseg_onnx_path
            else:
                raise NotImplementedError(
                    'the ONNX semseg network is outside the B200 hot path: pass an object with '
                    '.pred(rgb) -> (1,1,H,W) class map as semseg_onnx_path '
                    '(utils/onnx_utils.py:32-44 contract)')
        self.semseg_filters = semseg_filters
        self.sem_idxs = sem_idxs
        self.use_gt_sem = use_gt_sem
        self.icp_threshold = icp_threshold
        self.icp_trans_init = np.eye(4)
        self.T_prev_origin = np.eye(4)
        self.pcd_prev = None
        self.horizon_dist = horizon_dist

        self.poses = []
        self.seg_dists = []
        self.rgbs = []
        self.semsegs = []

        # device-resident cloud
        self.cloud = DeviceCloud(ring_capacity_pts, ring_max_frames, device)
        self._fids = []              # absolute frame ids of the live frames
        self.eager_rebase = False    # True = rewrite xyz every frame like the reference
        self.sync_each_integrate = True

        self.sem_bev_generator = None
        if bev_params['type'] == 'sem':
            self.sem_bev_generator = SemBEVGenerator(
                self.sem_idxs, bev_params['view_size'], bev_params['pixel_size'],
                bev_params['max_trans_radius'], bev_params['zoom_thresh'], bev_params['do_warp'],
                bev_params['int_scaler'], bev_params['int_sep_scaler'],
                bev_params['int_mid_threshold'], bev_params['height_filter'])
        elif bev_params['type'] == 'rgb':
            raise NotImplementedError('Needs refactoring')

    # ------------------------------------------------------------------
    @property
    def sem_pcs(self):
        return SemPcsView(self)

    def _export(self, fid):
        self._sync()
        return self.cloud.export_frame(fid)

    def _sync(self):
        if not self.cloud._dirty:
            return 0          # nothing was enqueued since the last sync: table and flags are current
        flags = self.cloud.sync()
        if flags & _lib.FLAG_UV_OUT_OF_IMAGE:
            raise AssertionError('pts_uv must be all inside image')
        if flags & (_lib.FLAG_ATTR_RANGE | _lib.FLAG_CELL_OVERFLOW):
            raise ValueError(f'libpcacc data error flags 0x{flags:x} '
                             '(class / colour outside 0..255 or instance outside int32)')
        return flags

    def integrate(self, observations: list):
        raise NotImplementedError()

    def obs2sem_vec_space(self, *a, **k):
        raise NotImplementedError()

    def generate_bev(self, *a, **k):
        """Dataset-specific (kitti360_sem_pc_accum.py:166-243, nuscenes_oracle_sem_pc_accum.py:421-530)."""
        raise NotImplementedError()

    # -- pose / cloud updates (sem_pc_accum.py:156-209) ------------------
    def update_poses(self, T_new_prev):
        """sem_pc_accum.py:156-165: every stored pose <- T_new_prev @ [pose 1].  One numpy call for all
        poses: a stacked (F,4,1) operand runs the reference's (4,4) @ (4,1) product once per pose, so
        the bits are the reference's (tests/test_cabi_cpu.py), without F interpreter round trips per
        integrated frame."""
        if not self.poses:
            return
        col = np.ones((len(self.poses), 4, 1))
        col[:, :3, 0] = self.poses
        self.poses = np.matmul(np.asarray(T_new_prev, dtype=np.float64), col)[:, :3, 0].tolist()

    def update_sem_pcs(self, T_new_prev):
        """Every stored point <- T_new_prev @ point.  On the device this is
        either an in-place kernel (eager) or one more link of the lazy chain."""
        self.cloud.rebase(T_new_prev, eager=self.eager_rebase)

    def remove_observations(self):
        idx = 0
        self.seg_dists.append(self.dist(np.array(self.poses[-1]), np.array(self.poses[-2])))
        path_length = np.sum(self.seg_dists)
        if path_length > self.horizon_dist:
            incr = self.get_incremental_path_dists()
            incr -= path_length - self.horizon_dist
            idx = (incr > 0.).argmax()
            self.cloud.evict(int(idx))
            self._fids = self._fids[idx:]
            self.poses = self.poses[idx:]
            self.seg_dists = self.seg_dists[idx:]
            self.rgbs = self.rgbs[idx:]
            self.semsegs = self.semsegs[idx:]
        return idx, path_length

    @staticmethod
    def comp_incr_path_dist(seg_dists):
        return np.matmul(np.tri(len(seg_dists)), np.array(seg_dists))

    def get_segment_dists(self):
        return self.seg_dists

    def get_incremental_path_dists(self):
        return self.comp_incr_path_dist(np.array(self.seg_dists))

    def get_pose(self, idx: int = None):
        return np.array(self.poses) if idx is None else np.array(self.poses[idx])

    def get_rgb(self, idx: int = None):
        return self.rgbs if idx is None else [self.rgbs[idx]]

    def get_semseg(self, idx: int = None):
        return self.semsegs if idx is None else [self.semsegs[idx]]

    @staticmethod
    def dist(pose_0, pose_1):
        return np.sqrt(np.sum((pose_1 - pose_0) ** 2))

    # -- output files (sem_pc_accum.py:280-308) ----------------------------
    @staticmethod
    def write_compressed_pickle(obj, filename, write_dir):
        path = os.path.join(write_dir, f'{filename}.gz')
        try:
            with gzip.open(path, 'wb') as f:
                f.write(pickle.dumps(obj))
        except IOError as error:
            print(error)

    @staticmethod
    def async_writer(n_threads: int = 8, compresslevel: int = 1, max_pending: int = 64):
        """Threaded write_compressed_pickle for dataset generation (run_*_bev_gen.py write one
        .gz per BEV): at device rates the gzip of 2.75 MB per BEV is the bottleneck of a
        single-threaded writer.  Same file format as write_compressed_pickle (any gzip level reads
        back identically); the default level is 1: measured on the bench host (16 threads, bench.py
        extra "writer") level 1 writes 366 BEVs/s at 452 KB per BEV, level 9 (gzip.open's default,
        what write_compressed_pickle uses) 31 BEVs/s at 364 KB."""
        return AsyncBevWriter(n_threads, compresslevel, max_pending)

    @staticmethod
    def read_compressed_pickle(path):
        try:
            with gzip.open(path, 'rb') as f:
                return pickle.loads(f.read())
        except IOError as error:
            print(error)

    # -- stand-alone operators (sem_pc_accum.py:317-402), device-backed ------
    def filter_semseg_pc(self, pc):
        """sem_pc_accum.py:317-321: rows whose last column is none of the filtered classes, input
        order (`pcacc_partition_semantic_pc`: the rows that are NOT selected)."""
        pc = np.ascontiguousarray(pc, dtype=np.float64)
        _, rest = self.cloud.partition_semantic_pc(pc, self.semseg_filters, pc.shape[1] - 1)
        return rest.cpu().numpy()

    def gen_semantic_pc(self, pc_velo, semantic_map, P_velo_frame):
        return self.cloud.gen_semantic_pc(pc_velo, semantic_map, P_velo_frame).cpu().numpy()

    def velo2img(self, pc_velo, P_velo_frame, img_h, img_w, max_depth=np.inf):
        """sem_pc_accum.py:367-402 (`pcacc_velo2img`)."""
        return self.cloud.velo2img(pc_velo, P_velo_frame, img_h, img_w, max_depth).cpu().numpy()

    @staticmethod
    def velo2frame(pc_velo, P_velo_frame):
        """sem_pc_accum.py:347-366: (N,3) points times the (3,4) matrix in homogeneous
        coordinates -> (N,3) float64 (`pcacc_velo2frame`)."""
        from .bev_generator.bev_generator import _ops_cloud, _to_like
        return _to_like(pc_velo, _ops_cloud().velo2frame(pc_velo, P_velo_frame))

    def viz_sem_vec_space(self, *a, **k):
        raise NotImplementedError('Open3D visualisation is outside the B200 hot path')

    def viz_sem_pc(self, *a, **k):
        raise NotImplementedError('Open3D visualisation is outside the B200 hot path')

    def viz_bev(self, *a, **k):
        raise NotImplementedError('plotting is outside the B200 hot path')

    def pc2pcd(self, *a, **k):
        raise NotImplementedError('Open3D point cloud objects are outside the B200 hot path')

    # -- shared part of generate_bev (kitti360_sem_pc_accum.py:166-243) -------
    def _window_inputs(self, present_idx, gen_future, other_trajs=None, gt_lanes=None):
        """The reference's window split (kitti360_sem_pc_accum.py:179-213): frame ranges instead
        of concatenated clouds; trajectories shifted by the present pose.  `other_trajs` may be a
        callable returning the (present, future, full) lists: the dict then produces the
        other-object entries on first access, which the generator makes after it has enqueued the
        device work (their extraction is host-only work that overlaps the kernels)."""
        n = len(self.poses)
        poses = np.array(self.poses, dtype=np.float64).reshape(n, 3)
        origin = poses[-1 if present_idx is None else present_idx].copy()
        lo, hi, _ = slice(None, present_idx).indices(n)
        flo, fhi, _ = slice(present_idx, None).indices(n)
        if hi <= lo:
            raise ValueError('need at least one array to concatenate')
        first = self._fids[0]
        shifted = poses - origin
        pcs = {'pc_present': DeviceWindow(self.cloud, first + lo, first + hi, origin)}
        trajs = _LazyTrajs()
        trajs['ego_traj_present'] = shifted[:present_idx]
        if gen_future:
            if fhi <= flo:
                raise ValueError('need at least one array to concatenate')
            pcs['pc_future'] = DeviceWindow(self.cloud, first + flo, first + fhi, origin)
            pcs['pc_full'] = DeviceWindow(self.cloud, first, first + n, origin)
            trajs['ego_traj_future'] = shifted[present_idx:]
            trajs['ego_traj_full'] = shifted
        else:
            pcs['pc_future'] = pcs['pc_full'] = None
            trajs['ego_traj_future'] = trajs['ego_traj_full'] = None

        def fill():
            others = other_trajs() if callable(other_trajs) else other_trajs

            def shift(ts):
                return [np.array(t, dtype=np.float64) - origin for t in ts] if ts else []

            out = {'other_trajs_present': shift(others[0] if others else None)}
            if gt_lanes is not None:
                out['gt_lanes'] = [lane - origin for lane in gt_lanes]
            if gen_future:
                out['other_trajs_future'] = shift(others[1] if others else None)
                out['other_trajs_full'] = shift(others[2] if others else None)
            else:
                out['other_trajs_future'] = out['other_trajs_full'] = None
            return out

        trajs._fill = fill
        return pcs, trajs

    def _generate(self, pcs, trajs, bev_num):
        self._sync()
        gen = self.sem_bev_generator
        if bev_num == 1:
            return [gen.generate_multiproc((pcs, trajs))]
        # bev_num variants of the same sample: one batched launch instead of the
        # reference's Pool(processes=bev_num) (kitti360_sem_pc_accum.py:236-241)
        if gen.do_aug:
            saved = gen.rng
            if saved is None:       # one stream for the whole batch, seeded like the reference
                gen.rng = np.random.RandomState((os.getpid() * int(time.time())) % 123456789)
            augs = [gen.rand_aug_params() for _ in range(bev_num)]
            gen.rng = saved
        else:
            augs = [dict(do_warping=False) for _ in range(bev_num)]
        return gen.generate_batch(pcs, trajs, augs)


class AsyncBevWriter:
    """write_compressed_pickle (sem_pc_accum.py:280-294) on a pool of worker threads.

    `submit(bev, filename, write_dir)` returns at once; the pickle + gzip + file write run in
    a worker (zlib releases the GIL).  The files are what the reference writes — a gzip
    stream of `pickle.dumps(obj)` named `<filename>.gz`, read back by `read_compressed_pickle`
    — at `compresslevel` (default 1; 9 is gzip.open's default and 12x slower for 20 % smaller
    files).  At most `max_pending` BEVs are queued
    (back-pressure instead of unbounded host memory).  Use as a context manager or call
    `close()`; errors of the workers are re-raised there (the reference prints IOError and
    goes on: pass `raise_errors=False` to `close` for that behaviour)."""

    def __init__(self, n_threads: int = 8, compresslevel: int = 1, max_pending: int = 64):
        import concurrent.futures
        import threading
        self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=max(1, int(n_threads)))
        self._level = int(compresslevel)
        self._slots = threading.BoundedSemaphore(max(1, int(max_pending)))
        self._futures = []
        self.n_written = 0
        self.bytes_written = 0

    def _work(self, obj, path):
        try:
            data = pickle.dumps(obj)
            with gzip.open(path, 'wb', compresslevel=self._level) as f:
                f.write(data)
            return os.path.getsize(path)
        finally:
            self._slots.release()

    def submit(self, obj, filename, write_dir):
        self._slots.acquire()
        path = os.path.join(write_dir, f'{filename}.gz')
        self._futures.append(self._pool.submit(self._work, obj, path))
        if len(self._futures) > 256:
            self._collect(only_done=True)

    def _collect(self, only_done=False, raise_errors=True):
        keep, first_error = [], None
        for f in self._futures:
            if only_done and not f.done():
                keep.append(f)
                continue
            try:
                self.bytes_written += f.result()
                self.n_written += 1
            except IOError as error:
                if first_error is None:
                    first_error = error
                if not raise_errors:
                    print(error)
        self._futures = keep
        if first_error is not None and raise_errors:
            raise first_error

    def close(self, raise_errors=True):
        try:
            self._collect(raise_errors=raise_errors)
        finally:
            self._pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close(raise_errors=exc[0] is None)
        return False
