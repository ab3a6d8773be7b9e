"""Device-resident accumulated cloud: the thin host layer between the Python
mirror of the reference classes and libpcacc's C ABI.

PyTorch is plumbing here (pinned staging, device buffers for inputs/outputs,
the current stream); every computation happens in libpcacc's sm_100a kernels.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import BevParams, PcaccError, check

_SEM_DTYPES = {np.dtype(np.uint8): _lib.SEM_U8, np.dtype(np.int32): _lib.SEM_I32,
               np.dtype(np.int64): _lib.SEM_I64, np.dtype(np.int16): _lib.SEM_I16}
_TORCH_SEM = {torch.uint8: _lib.SEM_U8, torch.int32: _lib.SEM_I32,
              torch.int64: _lib.SEM_I64, torch.int16: _lib.SEM_I16}


def _narrow_class_map(sem: np.ndarray) -> torch.Tensor:
    """Host class maps arrive as int64 (the ONNX argmax): 8 B/pixel over PCIe for values
    that fit a byte.  Clamp to [-1, 256] and send int16: negatives stay 'invalid', anything
    above 255 still trips the kernel's class-range flag."""
    return torch.clamp(torch.from_numpy(np.ascontiguousarray(sem)), -1, 256).to(torch.int16)


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array in page-locked host memory (cudaHostAlloc through torch's pinned allocator).
    Observations written into such arrays — by a dataloader, a semseg network's output copy —
    are read by the integrate kernels in place (PCACC_STAGE_DIRECT): no host-side copy at all."""
    require_cuda()
    shape = (int(shape),) if np.isscalar(shape) else tuple(int(d) for d in shape)
    t = torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
    return t.numpy()        # the array keeps the tensor (and its pinned block) alive


def pinned_like(arr) -> np.ndarray:
    """Page-locked copy of `arr` (same shape and dtype)."""
    arr = np.asarray(arr)
    out = pinned_empty(arr.shape, arr.dtype)
    out[...] = arr
    return out


def pin_observation(obs: dict) -> dict:
    """A nuScenes observation dict (obs_dataloaders/nuscenes_obs_dataloader.py:103-123) whose
    bulk arrays — pc, pc_cam_idx, images and, if present, the `_semseg` class maps of the
    synthetic stand-in — live in page-locked memory; everything else is passed through."""
    out = dict(obs)
    out['pc'] = pinned_like(np.asarray(obs['pc'], dtype=np.float64))
    out['pc_cam_idx'] = pinned_like(np.asarray(obs['pc_cam_idx'], dtype=np.int64))
    out['images'] = [pinned_like(np.asarray(i, dtype=np.uint8)) for i in obs['images']]
    if '_semseg' in obs:
        out['_semseg'] = [pinned_like(c) for c in obs['_semseg']]
    return out


def require_cuda():
    if not torch.cuda.is_available():
        raise PcaccError(_lib.ERR_CUDA,
                         'no CUDA device: pc_accumulation_lib_b200 has no CPU path')


class _PinnedOutPool:
    """Pinned float16 buffers for results handed to the caller as numpy arrays.  cudaHostAlloc
    costs about a millisecond per 10 MB, so buffers are recycled: a buffer goes back to the
    pool when the numpy array created over it (and thereby every view of it) has been
    collected.  `max_bytes` bounds the pinned memory held by live results; beyond it the
    caller falls back to a pageable copy."""

    def __init__(self, max_bytes=2 << 30, quantum=1 << 20):
        self.max_bytes, self.quantum = max_bytes, quantum
        self.free = {}
        self.total = 0

    def take(self, nbytes):
        size = (max(int(nbytes), 1) + self.quantum - 1) // self.quantum * self.quantum
        lst = self.free.get(size)
        if lst:
            return lst.pop()
        if self.total + size > self.max_bytes:
            return None
        self.total += size
        return torch.empty(size // 2, dtype=torch.float16, pin_memory=True)

    def give(self, buf):
        self.free.setdefault(buf.numel() * 2, []).append(buf)


_OUT_POOL = _PinnedOutPool()


class Stager:
    """Host -> device staging of inputs through reusable pinned buffers (two per key, used
    alternately, so the copy of the next observation can start while the kernels of the
    previous one still read theirs).

    `put(..., mapped=True)` stops at the pinned buffer and hands its address to the kernel:
    pinned memory is device-accessible under unified addressing, and the integrate kernels
    only sample a few thousand pixels of each camera map — reading those over PCIe beats
    DMA-ing the whole image into HBM first."""

    def __init__(self, device):
        self.device = device
        self._slots = {}     # key -> [[pinned, device, event, ...], [pinned, device, event, ...], turn]
        self._touched = []
        self.keepalive = []   # caller-owned pinned arrays handed to kernels / DMA; cleared by the owner after a sync
        # bare CUDA events from libpcacc (see pcacc_event_record): four, used in turn
        self._lib = _lib.load()
        self._events = []
        for _ in range(8):
            ev = C.c_void_p()
            check(self._lib.pcacc_event_create(C.byref(ev)))
            self._events.append(ev)
        self._ev_turn = 0

    def __del__(self):
        try:
            for ev in self._events:
                self._lib.pcacc_event_destroy(ev)
        except Exception:
            pass

    def _slot(self, key, nbytes, need_dev):
        ring = self._slots.get(key)
        if ring is None:
            ring = self._slots[key] = [None, None, 0]
        turn = ring[2]
        ring[2] = turn ^ 1
        sl = ring[turn]
        if sl is None or sl[0].numel() < nbytes or (need_dev and sl[1] is None):
            if sl is not None and sl[2] is not None:
                # kernels may still read the old pinned block through a raw pointer torch knows
                # nothing about: wait for its last reader before it returns to the allocator
                self._lib.pcacc_event_sync(sl[2])
            cap = max(nbytes, 1) * 5 // 4 + 64
            sl = ring[turn] = [torch.empty(cap, dtype=torch.uint8, pin_memory=True),
                               torch.empty(cap, dtype=torch.uint8, device=self.device)
                               if need_dev else None, None]     # no reader yet: nothing to wait for
        return sl

    def put(self, key, arr, mapped=False):
        """numpy array / CPU tensor -> tensor of the same dtype and shape whose data_ptr() the
        kernels may dereference: a device copy, or (mapped) the pinned host buffer itself.
        CUDA tensors pass through."""
        if isinstance(arr, torch.Tensor):
            if arr.is_cuda:
                return arr.contiguous()
            t = arr.contiguous()
        else:
            t = torch.from_numpy(np.ascontiguousarray(arr))
        nbytes = t.numel() * t.element_size()
        if nbytes and self._lib.pcacc_host_is_pinned(t.data_ptr()):
            # the caller's array is page-locked already (pinned_empty / pin_memory): no staging
            # copy.  mapped: the kernel reads it in place; otherwise one DMA straight from it.
            # The caller keeps `keepalive` entries until the stream has passed the consumers.
            self.keepalive.append(t)
            if mapped:
                return t
            sl = self._slot(key, nbytes, True)
            dev_v = sl[1][:nbytes].view(t.dtype).view(t.shape)
            if sl[2] is not None:
                self._lib.pcacc_event_sync(sl[2])
            dev_v.copy_(t, non_blocking=True)
            self._touched.append(sl)
            return dev_v
        sl = self._slot(key, nbytes, not mapped)
        # typed views of the slot's buffers are cached: observations of one stream keep
        # their shapes, and a dozen tensor-view calls per put cost more than the copy
        sig = (t.dtype, tuple(t.shape), mapped)
        if len(sl) < 5 or sl[3] != sig:
            pin_v = sl[0][:nbytes].view(t.dtype).view(t.shape)
            dev_v = None if mapped else sl[1][:nbytes].view(t.dtype).view(t.shape)
            del sl[3:]
            sl += [sig, pin_v, dev_v]
        pin_v, dev_v = sl[4], sl[5]
        if nbytes:
            if sl[2] is not None:   # everything that read this buffer two puts ago is done
                self._lib.pcacc_event_sync(sl[2])
            pin_v.copy_(t)          # torch splits large host copies over its thread pool
            if not mapped:
                dev_v.copy_(pin_v, non_blocking=True)
            self._touched.append(sl)
        return pin_v if mapped else dev_v

    def next_event(self):
        """A bare CUDA event from the ring (valid until PCACC ring-size further requests)."""
        ev = self._events[self._ev_turn]
        self._ev_turn = (self._ev_turn + 1) % len(self._events)
        return ev

    def fence(self):
        """Call after enqueuing the kernels that consume the buffers put since the last fence.
        One event per fence (a ring of four: a slot is reused two fences later), recorded on
        torch's current stream through libpcacc — Event.record() costs a
        torch.cuda.current_stream() call (~25 us)."""
        if not self._touched:
            return
        ring = self._events
        ev = ring[self._ev_turn]
        self._ev_turn = (self._ev_turn + 1) % len(ring)
        check(self._lib.pcacc_event_record(ev, _stream()))
        for sl in self._touched:
            sl[2] = ev
        self._touched = []


def _stream():
    """cudaStream_t of torch's current stream.  torch.cuda.current_stream() costs ~25 us
    per call (more than the kernels it launches); the raw accessor is ~100x cheaper."""
    try:
        return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))
    except AttributeError:      # older / newer torch without the private accessor
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr() if t is not None and t.numel() else 0)


def _hostd(a, n):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    assert a.size == n, (a.size, n)
    return a


class DeviceCloud:
    """Owns one pcacc handle: the SoA ring with its frame table on `device`."""

    def __init__(self, capacity_pts: int, max_frames: int = 1024, device: int | None = None):
        require_cuda()
        self.lib = _lib.load()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device('cuda', self.device_index)
        self.capacity = int(capacity_pts)
        self.max_frames = int(max_frames)
        h = C.c_void_p()
        check(self.lib.pcacc_create(self.device_index, self.capacity, self.max_frames,
                                    C.byref(h)))
        self.h = h
        self.stage = Stager(self.device)
        self._keep = self.stage.keepalive   # one list: everything in-flight work still reads
        # camera maps handed over as host arrays are read in place from pinned memory
        self.map_images = True
        self._filt_cache = None
        self._fid_out, self._staging_out = C.c_int64(-1), C.c_int(0)
        self._fid_ref, self._staging_ref = C.byref(self._fid_out), C.byref(self._staging_out)
        self.last_staging = None
        self._dirty = True    # frames or flags may have changed since the last sync()
        self._mark_f, self._mark_i = [], []   # queued dynamic-flag updates

    def close(self):
        if getattr(self, 'h', None) is not None and self.h:
            self.lib.pcacc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers -------------------------------------------------------------
    def _check(self, status):
        check(status, self.h)

    def _filters(self, filters):
        f = np.ascontiguousarray(np.asarray(list(filters or []), dtype=np.int32))
        return f, f.ctypes.data_as(C.c_void_p), int(f.size)

    def _sem_arg(self, sem, key, mapped=False):
        """class map (H,W) of u8/i16/i32/i64, or (H,W,K) float32 probabilities."""
        if mapped and not isinstance(sem, torch.Tensor):
            sem = np.asarray(sem)
            if sem.dtype in _SEM_DTYPES:     # read in place over PCIe: no narrowing needed
                return self.stage.put(key, sem, mapped=True), _SEM_DTYPES[sem.dtype], 1
        if isinstance(sem, torch.Tensor):
            if sem.dtype == torch.float32:
                return self.stage.put(key, sem), _lib.SEM_F32_PROB, int(sem.shape[-1])
            if sem.dtype not in _TORCH_SEM:
                sem = sem.to(torch.int64)
            return self.stage.put(key, sem), _TORCH_SEM[sem.dtype], 1
        sem = np.asarray(sem)
        if sem.dtype == np.float32 and sem.ndim == 3:
            return self.stage.put(key, sem), _lib.SEM_F32_PROB, int(sem.shape[-1])
        if sem.dtype in (np.int64, np.int32):
            return self.stage.put(key, _narrow_class_map(sem)), _lib.SEM_I16, 1
        if sem.dtype not in _SEM_DTYPES:
            return self.stage.put(key, _narrow_class_map(sem.astype(np.int64))), _lib.SEM_I16, 1
        return self.stage.put(key, sem), _SEM_DTYPES[sem.dtype], 1

    # -- lifetime --------------------------------------------------------------
    def reset(self):
        self._dirty = True
        self._mark_f, self._mark_i = [], []
        self._check(self.lib.pcacc_reset(self.h, _stream()))

    def sync(self) -> int:
        """Waits for the stream, refreshes the frame table, returns and clears
        the data-error flags."""
        self.flush_marks()
        fl = C.c_uint32(0)
        self._check(self.lib.pcacc_sync(self.h, C.byref(fl), _stream()))
        self._keep.clear()
        self._dirty = False
        return int(fl.value)

    def refresh(self):
        """Waits for the stream and refreshes the frame table like sync(), but leaves the
        data-error flags pending for the next sync() (they must reach the caller that checks them)."""
        self.flush_marks()
        self._check(self.lib.pcacc_sync(self.h, None, _stream()))
        self._keep.clear()

    def live_frames(self):
        first, n = C.c_int64(0), C.c_int(0)
        self._check(self.lib.pcacc_num_frames(self.h, C.byref(first), C.byref(n)))
        return int(first.value), int(n.value)

    def frame_count(self, fid: int) -> int:
        c = C.c_int64(0)
        self._check(self.lib.pcacc_frame_count(self.h, int(fid), C.byref(c)))
        return int(c.value)

    def frame_offset(self, fid: int) -> int:
        c = C.c_int64(0)
        self._check(self.lib.pcacc_frame_offset(self.h, int(fid), C.byref(c)))
        return int(c.value)

    def resident_points(self) -> int:
        return int(self.lib.pcacc_resident_points(self.h))

    # -- integrate ---------------------------------------------------------------
    @staticmethod
    def _f32_cloud(pc):
        """KITTI-360 clouds are float32 (the reference reads them from .bin files); the frustum
        kernels take float32 input.  A wider array is accepted only if the cast loses nothing —
        otherwise the projection would silently differ from the reference's float64 result."""
        if isinstance(pc, torch.Tensor):
            return pc
        pc = np.asarray(pc)
        if pc.dtype == np.float32:
            return pc
        pc32 = pc.astype(np.float32)
        if not np.array_equal(pc32.astype(pc.dtype), pc, equal_nan=True):
            raise ValueError('point cloud values are not representable in float32: the frustum '
                             'path takes float32 clouds (as the KITTI-360 reader delivers them)')
        return pc32

    def integrate_frustum(self, pc, P, rgb, sem, filters, max_depth=np.inf) -> int:
        self._dirty = True
        pts = self.stage.put('pc', self._f32_cloud(pc))
        assert pts.dim() == 2 and pts.shape[1] == 4 and pts.dtype == torch.float32
        rgb_d = self.stage.put('rgb', rgb if isinstance(rgb, torch.Tensor)
                               else np.asarray(rgb, dtype=np.uint8), mapped=self.map_images)
        sem_d, sem_dt, K = self._sem_arg(sem, 'sem', mapped=self.map_images)
        h, w = int(rgb_d.shape[0]), int(rgb_d.shape[1])
        assert tuple(sem_d.shape[:2]) == (h, w), (sem_d.shape, rgb_d.shape)
        Pm = _hostd(P, 12)
        f, fp, nf = self._filters(filters)
        fid = C.c_int64(-1)
        self._check(self.lib.pcacc_integrate_frustum(
            self.h, _ptr(pts), int(pts.shape[0]), Pm.ctypes.data_as(C.c_void_p), _ptr(rgb_d),
            _ptr(sem_d), sem_dt, K, h, w, float(max_depth), fp, nf, C.byref(fid), _stream()))
        self.stage.fence()
        return int(fid.value)

    def integrate_gt(self, pc, sem_gt, filters) -> int:
        self._dirty = True
        pts = self.stage.put('pc', self._f32_cloud(pc))
        assert pts.dim() == 2 and pts.shape[1] == 4 and pts.dtype == torch.float32
        if isinstance(sem_gt, torch.Tensor):
            sg = self.stage.put('sem_gt', sem_gt.reshape(-1).to(torch.int16))
        else:
            sg = self.stage.put('sem_gt', np.asarray(sem_gt)[:, -1].astype(np.int16))
        assert sg.numel() == pts.shape[0]
        f, fp, nf = self._filters(filters)
        fid = C.c_int64(-1)
        self._check(self.lib.pcacc_integrate_gt(self.h, _ptr(pts), int(pts.shape[0]), _ptr(sg),
                                                fp, nf, C.byref(fid), _stream()))
        self.stage.fence()
        return int(fid.value)

    def integrate_records(self, pc, cam_idx, rgbs, sems, T_ego_world, filters,
                          intensity_div=255.) -> int:
        self._dirty = True
        pcd = self.stage.put('pc', pc if isinstance(pc, torch.Tensor)
                             else np.asarray(pc, dtype=np.float64))
        assert pcd.dim() == 2 and pcd.shape[1] == 7 and pcd.dtype == torch.float64
        cam = self.stage.put('cam', cam_idx if isinstance(cam_idx, torch.Tensor)
                             else np.asarray(cam_idx, dtype=np.int64))
        assert cam.dtype == torch.int64 and cam.numel() == pcd.shape[0]
        n_cams = len(rgbs)
        rgb_d, sem_d, sem_dt = [], [], None
        for k in range(n_cams):
            r = self.stage.put(f'rgb{k}', rgbs[k] if isinstance(rgbs[k], torch.Tensor)
                               else np.asarray(rgbs[k], dtype=np.uint8), mapped=self.map_images)
            s, dt, K = self._sem_arg(sems[k], f'sem{k}', mapped=self.map_images)
            assert K == 1 and dt != _lib.SEM_F32_PROB
            assert sem_dt in (None, dt), 'all class maps must share one dtype'
            sem_dt = dt
            rgb_d.append(r)
            sem_d.append(s)
        h, w = (int(rgb_d[0].shape[0]), int(rgb_d[0].shape[1])) if n_cams else (1, 1)
        rp = (C.c_void_p * max(n_cams, 1))(*[r.data_ptr() for r in rgb_d])
        sp = (C.c_void_p * max(n_cams, 1))(*[s.data_ptr() for s in sem_d])
        T = _hostd(T_ego_world, 16)
        f, fp, nf = self._filters(filters)
        fid = C.c_int64(-1)
        self._check(self.lib.pcacc_integrate_records(
            self.h, _ptr(pcd), _ptr(cam), int(pcd.shape[0]), C.cast(rp, C.c_void_p),
            C.cast(sp, C.c_void_p), n_cams, sem_dt if sem_dt is not None else _lib.SEM_U8, h, w,
            T.ctypes.data_as(C.c_void_p), float(intensity_div), fp, nf, C.byref(fid), _stream()))
        self.stage.fence()
        return int(fid.value)

    def integrate_records_host(self, pc, cam_idx, rgbs, sems, T_ego_world, filters,
                               intensity_div=255., staging=_lib.STAGE_AUTO) -> int:
        """integrate_records for HOST numpy arrays through ONE C call
        (pcacc_integrate_records_host): page-locked arrays are read by the kernel in place,
        pageable ones go through the sparse staging mode.  Returns the frame id;
        `self.last_staging` tells which mode ran."""
        self._dirty = True
        if not (type(pc) is np.ndarray and pc.dtype == np.float64 and pc.flags.c_contiguous):
            pc = np.ascontiguousarray(pc, dtype=np.float64)
        if not (type(cam_idx) is np.ndarray and cam_idx.dtype == np.int64 and cam_idx.flags.c_contiguous):
            cam_idx = np.ascontiguousarray(cam_idx, dtype=np.int64)
        n = pc.shape[0]
        assert pc.ndim == 2 and pc.shape[1] == 7 and cam_idx.size == n, (pc.shape, cam_idx.shape)
        n_cams = len(rgbs)
        sem_dt, h, w = _lib.SEM_U8, 1, 1
        if n_cams:
            rr, ss = [], []
            for r, m in zip(rgbs, sems):
                if not (type(r) is np.ndarray and r.dtype == np.uint8 and r.flags.c_contiguous):
                    r = np.ascontiguousarray(r, dtype=np.uint8)
                if not (type(m) is np.ndarray and m.flags.c_contiguous and m.dtype in _SEM_DTYPES):
                    m = np.ascontiguousarray(m)
                    if m.dtype not in _SEM_DTYPES:
                        m = m.astype(np.int64)
                rr.append(r)
                ss.append(m)
            rgbs, sems = rr, ss
            sem_dt = _SEM_DTYPES[sems[0].dtype]
            h, w = rgbs[0].shape[0], rgbs[0].shape[1]
            for r, m in zip(rgbs, sems):
                assert r.shape == (h, w, 3) and m.shape == (h, w) and _SEM_DTYPES[m.dtype] == sem_dt, \
                    (r.shape, m.shape, m.dtype)
        rp = (C.c_void_p * max(n_cams, 1))(*[r.__array_interface__['data'][0] for r in rgbs])
        sp = (C.c_void_p * max(n_cams, 1))(*[m.__array_interface__['data'][0] for m in sems])
        T = T_ego_world
        if not (type(T) is np.ndarray and T.dtype == np.float64 and T.flags.c_contiguous and T.size == 16):
            T = _hostd(T_ego_world, 16)
        fc = self._filt_cache
        if fc is None or fc[0] is not filters or fc[1] != len(filters or ()):
            f, _, nf = self._filters(filters)
            fc = self._filt_cache = (filters, len(filters or ()), f, f.__array_interface__['data'][0], nf)
        self._check(self.lib.pcacc_integrate_records_host(
            self.h, pc.__array_interface__['data'][0], cam_idx.__array_interface__['data'][0], n,
            rp, sp, n_cams, sem_dt, h, w, T.__array_interface__['data'][0], float(intensity_div),
            fc[3], fc[4], staging, self._staging_ref, self._fid_ref, _stream()))
        self.last_staging = self._staging_out.value
        if self.last_staging == _lib.STAGE_DIRECT:
            # the kernel reads these arrays in place: keep them alive until the next sync
            self._keep.append((pc, cam_idx, rgbs, sems))
        return self._fid_out.value

    def integrate_records_batch(self, sweeps, filters, intensity_div=255.) -> int:
        """sweeps: list of dicts with DEVICE tensors pc (n,7) f64, cam (n,) i64,
        rgb [ (H,W,3) u8 ], sem [ (H,W) u8|i32|i64 ] and host T (4,4).  One launch
        for the whole list; returns the first frame id."""
        return self.integrate_prepared(self.prepare_records_batch(sweeps, filters, intensity_div))

    def prepare_records_batch(self, sweeps, filters, intensity_div=255.):
        """The argument block of integrate_records_batch, built once for inputs that stay
        resident (offline dataset generation re-integrates the same device buffers)."""
        n_s = len(sweeps)
        n_cams = len(sweeps[0]['rgb'])
        sem_dt = _TORCH_SEM[sweeps[0]['sem'][0].dtype] if n_cams else _lib.SEM_U8
        h, w = ((int(sweeps[0]['rgb'][0].shape[0]), int(sweeps[0]['rgb'][0].shape[1]))
                if n_cams else (1, 1))
        pcp = (C.c_void_p * n_s)(*[s['pc'].data_ptr() for s in sweeps])
        cmp_ = (C.c_void_p * n_s)(*[s['cam'].data_ptr() for s in sweeps])
        nn = (C.c_int64 * n_s)(*[int(s['pc'].shape[0]) for s in sweeps])
        m = max(n_s * n_cams, 1)
        rp = (C.c_void_p * m)(*[r.data_ptr() for s in sweeps for r in s['rgb']])
        sp = (C.c_void_p * m)(*[r.data_ptr() for s in sweeps for r in s['sem']])
        T = np.ascontiguousarray(np.stack([np.asarray(s['T'], dtype=np.float64).reshape(4, 4)
                                           for s in sweeps]))
        f, fp, nf = self._filters(filters)
        args = (n_s, C.cast(pcp, C.c_void_p), C.cast(cmp_, C.c_void_p), C.cast(nn, C.c_void_p),
                C.cast(rp, C.c_void_p), C.cast(sp, C.c_void_p), n_cams, sem_dt, h, w,
                T.ctypes.data_as(C.c_void_p), float(intensity_div), fp, nf)
        return {'args': args, 'keep': (sweeps, pcp, cmp_, nn, rp, sp, T, f)}

    def integrate_prepared(self, prep) -> int:
        self._dirty = True
        fid = C.c_int64(-1)
        self._check(self.lib.pcacc_integrate_records_batch(self.h, *prep['args'], C.byref(fid),
                                                           _stream()))
        return int(fid.value)

    def integrate_cloud(self, rec) -> int:
        self._dirty = True
        r = self.stage.put('cloud', rec if isinstance(rec, torch.Tensor)
                           else np.asarray(rec, dtype=np.float64))
        assert r.dim() == 2 and r.shape[1] == 10 and r.dtype == torch.float64
        fid = C.c_int64(-1)
        self._check(self.lib.pcacc_integrate_cloud(self.h, _ptr(r), int(r.shape[0]),
                                                   C.byref(fid), _stream()))
        self.stage.fence()
        return int(fid.value)

    # -- state updates -------------------------------------------------------------
    def rebase(self, T_new_prev, eager=False):
        self._dirty = True
        T = _hostd(T_new_prev, 16)
        self._check(self.lib.pcacc_rebase(self.h, T.ctypes.data_as(C.c_void_p),
                                          1 if eager else 0, _stream()))

    def evict(self, n_frames: int):
        self._check(self.lib.pcacc_evict(self.h, int(n_frames)))

    def mark_dynamic(self, frame_ids, inst_idx):
        """dyn = 1 on the points of frame_ids[k] whose inst == inst_idx[k].  The
        flags are only read by rasterise / export, so the pairs are queued on the
        host and written by ONE launch right before the next reader."""
        assert len(frame_ids) == len(inst_idx)
        self._mark_f.extend(int(f) for f in frame_ids)
        self._mark_i.extend(int(i) for i in inst_idx)

    def mark_dynamic_now(self, frame_ids: np.ndarray, inst_idx: np.ndarray):
        """Unqueued variant for callers that already hold contiguous int64 / int32 arrays."""
        if frame_ids.size:
            self._check(self.lib.pcacc_mark_dynamic(self.h, frame_ids.ctypes.data_as(C.c_void_p),
                                                    inst_idx.ctypes.data_as(C.c_void_p),
                                                    int(frame_ids.size), _stream()))

    def flush_marks(self):
        if not self._mark_f:
            return
        first, n_live = self.live_frames()
        pairs = [(f, i) for f, i in zip(self._mark_f, self._mark_i) if f >= first]
        self._mark_f, self._mark_i = [], []
        if not pairs:
            return
        fi = np.ascontiguousarray(np.array([p[0] for p in pairs], dtype=np.int64))
        ii = np.ascontiguousarray(np.array([p[1] for p in pairs], dtype=np.int32))
        self._check(self.lib.pcacc_mark_dynamic(self.h, fi.ctypes.data_as(C.c_void_p),
                                                ii.ctypes.data_as(C.c_void_p), int(fi.size),
                                                _stream()))

    # -- export ----------------------------------------------------------------------
    def export_frame(self, fid: int, to_host=True):
        self.flush_marks()
        n = self.frame_count(fid)
        out = torch.empty((n, 10), dtype=torch.float64, device=self.device)
        self._check(self.lib.pcacc_export_frame(self.h, int(fid), _ptr(out), _stream()))
        return out.cpu().numpy() if to_host else out

    # -- stand-alone operators -----------------------------------------------------------
    def project(self, pc, P, img_h, img_w, max_depth=np.inf):
        pts = self.stage.put('pc', self._f32_cloud(pc))
        n, stride = int(pts.shape[0]), int(pts.shape[1])
        u = torch.empty(n, dtype=torch.int32, device=self.device)
        v = torch.empty(n, dtype=torch.int32, device=self.device)
        m = torch.empty(n, dtype=torch.uint8, device=self.device)
        Pm = _hostd(P, 12)
        self._check(self.lib.pcacc_project(_ptr(pts), n, stride, Pm.ctypes.data_as(C.c_void_p),
                                           int(img_h), int(img_w), float(max_depth), _ptr(u),
                                           _ptr(v), _ptr(m), _stream()))
        self.stage.fence()
        return u, v, m

    def velo2img(self, pc, P, img_h, img_w, max_depth=np.inf):
        """sem_pc_accum.py:367-402 -> (M, cols + 2) float64 CUDA tensor [pc_velo, u, v] of the points
        inside the image, input order (one kernel: projection + order-preserving compaction)."""
        pts = self.stage.put('pc', self._f32_cloud(pc)).contiguous()
        assert pts.dtype == torch.float32 and pts.dim() == 2 and pts.shape[1] >= 3
        n, stride = int(pts.shape[0]), int(pts.shape[1])
        out = torch.empty((n, stride + 2), dtype=torch.float64, device=self.device)
        nk = torch.zeros(1, dtype=torch.int64, device=self.device)
        Pm = _hostd(P, 12)
        self._check(self.lib.pcacc_velo2img(self.h, _ptr(pts), n, stride, Pm.ctypes.data_as(C.c_void_p),
                                            int(img_h), int(img_w), float(max_depth), _ptr(out), _ptr(nk),
                                            _stream()))
        self.stage.fence()
        return out[:int(nk.item())]

    def gen_semantic_pc(self, pc, semantic_map, P):
        """(M,4+K) float64 device tensor, rows in input order."""
        pts = self.stage.put('pc', self._f32_cloud(pc))
        assert pts.shape[1] == 4 and pts.dtype == torch.float32
        sm = semantic_map
        if not isinstance(sm, torch.Tensor):
            sm = np.asarray(sm)
            if sm.dtype not in _SEM_DTYPES and sm.dtype != np.float32:
                sm = sm.astype(np.float32 if sm.dtype.kind == 'f' else np.int64)
        mp = self.stage.put('map', sm)
        if mp.dtype == torch.float32:
            dt = _lib.SEM_F32_PROB
        else:
            dt = _TORCH_SEM[mp.dtype]
        h, w, K = int(mp.shape[0]), int(mp.shape[1]), int(mp.shape[2])
        n = int(pts.shape[0])
        out = torch.empty((n, 4 + K), dtype=torch.float64, device=self.device)
        nk = torch.zeros(1, dtype=torch.int64, device=self.device)
        Pm = _hostd(P, 12)
        self._check(self.lib.pcacc_gen_semantic_pc(
            self.h, _ptr(pts), n, Pm.ctypes.data_as(C.c_void_p), _ptr(mp), dt, h, w, K, _ptr(out),
            _ptr(nk), _stream()))
        self.stage.fence()
        return out[:int(nk.item())]

    # -- rasterise ------------------------------------------------------------------------
    def rasterise(self, params, P: int, want_f64=False, want_cells=False, out=None):
        """params: list of BevParams. Returns (planes f16 (V,3,7,P,P) device
        tensor, planes f64 or None, per-ring-position cell index or None)."""
        self.flush_marks()
        V = len(params)
        arr = params if isinstance(params, C.Array) else (BevParams * V)(*params)
        if out is None:
            out = torch.empty((V, 3, 7, P, P), dtype=torch.float16, device=self.device)
        o64 = (torch.empty((V, 3, 7, P, P), dtype=torch.float64, device=self.device)
               if want_f64 else None)
        cells = (torch.full((self.capacity + 4,), -2, dtype=torch.int32, device=self.device)
                 if want_cells else None)
        self._check(self.lib.pcacc_rasterise(self.h, arr, V, int(P), _ptr(out), _ptr(o64),
                                             _ptr(cells), _stream()))
        return out, o64, cells

    def set_option(self, option: int, value: int):
        self._check(self.lib.pcacc_set_option(self.h, int(option), int(value)))

    def profile(self, enable=True):
        """enable: False = off, True = every kernel class, or an iterable of class names
        (_lib.KERNEL_CLASSES) to time only those."""
        if enable is True:
            mask = (1 << len(_lib.KERNEL_CLASSES)) - 1
        elif not enable:
            mask = 0
        else:
            mask = sum(1 << _lib.KERNEL_CLASSES.index(k) for k in enable)
        self._check(self.lib.pcacc_profile(self.h, mask))

    def profile_read(self):
        """{kernel class: (device ms from CUDA events, launches)} since the last read."""
        ms, n = (C.c_double * 10)(), (C.c_int64 * 10)()
        self._check(self.lib.pcacc_profile_read(self.h, C.byref(ms), C.byref(n)))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(_lib.KERNEL_CLASSES)}

    def warp_planes(self, planes, imaps, jmaps):
        """planes (V,3,7,P,P) float16 device tensor, imaps / jmaps (V,P) int source indices
        -> warped planes (new tensor)."""
        V, P = int(planes.shape[0]), int(planes.shape[-1])
        n_planes = int(planes.numel() // (V * P * P))
        im = np.ascontiguousarray(np.asarray(imaps, dtype=np.int32).reshape(V, P))
        jm = np.ascontiguousarray(np.asarray(jmaps, dtype=np.int32).reshape(V, P))
        out = torch.empty_like(planes)
        self._check(self.lib.pcacc_warp_planes(self.h, _ptr(planes), _ptr(out), V, n_planes, P,
                                               im.ctypes.data_as(C.c_void_p),
                                               jm.ctypes.data_as(C.c_void_p), _stream()))
        return out

    def planes_to_host(self, planes):
        """(V,3,7,P,P) float16 device tensor -> numpy, through a reusable pinned buffer."""
        return self.planes_to_host_finish(self.planes_to_host_begin(planes))

    def planes_to_host_begin(self, planes):
        """Enqueue the device -> pinned copy; the caller may do host work before `_finish`.
        The planes land in a pinned buffer of their own (drawn from a pool, returned to it
        when the last numpy view of them is garbage-collected), so no second host copy is
        needed; beyond the pool's cap they go through one shared pinned buffer and a copy."""
        n = planes.numel()
        own = _OUT_POOL.take(n * 2)
        if own is not None:
            view = own[:n].view(planes.shape)
        else:
            buf = getattr(self, '_pin_out', None)
            if buf is None or buf.numel() < n:
                buf = self._pin_out = torch.empty(max(n, 1), dtype=torch.float16, pin_memory=True)
            view = buf[:n].view(planes.shape)
        st = _stream()
        check(self.lib.pcacc_memcpy_d2h_async(view.data_ptr(), planes.data_ptr(), n * 2, st))
        ev = self.stage.next_event()
        check(self.lib.pcacc_event_record(ev, st))
        return view, ev, planes, own      # `planes` stays referenced until the copy is done

    def planes_to_host_finish(self, pending):
        view, ev, _, own = pending
        check(self.lib.pcacc_event_sync(ev))
        if own is not None:
            out = view.numpy()              # shares the pinned buffer
            weakref.finalize(out, _OUT_POOL.give, own)
            return out
        out = np.empty(tuple(view.shape), dtype=np.float16)
        torch.from_numpy(out).copy_(view)       # multi-threaded for large blocks, unlike ndarray.copy
        return out

    # -- the dataloader's input side -----------------------
    def assign_boxes(self, points, target_from_boxes, sizes, tolerance):
        """Box loop of inst_centric_get_sweeps (datasets/nuscenes_utils.py:412-470).
        points: (n, >=3) float32/float64 numpy array or CUDA tensor; target_from_boxes: list of
        (4,4); sizes: list of dxdydz.  -> (box index per point int32 CUDA tensor, -1 = none;
        points inside each box int32 CUDA tensor)."""
        pts = self.stage.put('box_pts', points)
        if pts.dtype not in (torch.float32, torch.float64):
            pts = pts.double()
        pts = pts.contiguous()
        n, stride = int(pts.shape[0]), int(pts.shape[1])
        nb = len(target_from_boxes)
        inv = np.ascontiguousarray(np.stack([np.linalg.inv(np.asarray(T, dtype=np.float64))
                                             for T in target_from_boxes])) if nb else np.zeros((0, 4, 4))
        sz = np.ascontiguousarray(np.asarray(sizes, dtype=np.float64).reshape(nb, 3)) if nb else np.zeros((0, 3))
        box = torch.empty(n, dtype=torch.int32, device=self.device)
        cnt = torch.empty(nb, dtype=torch.int32, device=self.device)
        self._check(self.lib.pcacc_assign_boxes(
            self.h, _ptr(pts), int(pts.dtype == torch.float32), n, stride,
            inv.ctypes.data_as(C.c_void_p), sz.ctypes.data_as(C.c_void_p), nb, float(tolerance),
            _ptr(box), _ptr(cnt), _stream()))
        self.stage.fence()
        return box, cnt

    def project_cameras(self, pc_in_ego, glob_from_ego, cams, depth_thres=1e-3):
        """nuscenes_obs_dataloader.py:176-198.  pc_in_ego (n,3) float64; cams: list of dicts
        glob_from_self (4,4), cam_K (3,3), img_wh (2,).  -> (pc_uv (n,2) float64, pc_cam_idx
        (n,) int64) CUDA tensors."""
        pts = self.stage.put('cam_pts', pc_in_ego if isinstance(pc_in_ego, torch.Tensor)
                             else np.asarray(pc_in_ego, dtype=np.float64)).contiguous()
        assert pts.dtype == torch.float64 and pts.dim() == 2 and pts.shape[1] >= 3
        n, stride = int(pts.shape[0]), int(pts.shape[1])
        nc = len(cams)
        g = _hostd(glob_from_ego, 16)
        inv = np.ascontiguousarray(np.stack([np.linalg.inv(np.asarray(c['glob_from_self'], dtype=np.float64))
                                             for c in cams])) if nc else np.zeros((0, 4, 4))
        K = np.ascontiguousarray(np.stack([np.asarray(c['cam_K'], dtype=np.float64).reshape(3, 3)
                                           for c in cams])) if nc else np.zeros((0, 3, 3))
        wh = np.ascontiguousarray(np.stack([np.asarray(c['img_wh'], dtype=np.float64).reshape(2)
                                            for c in cams])) if nc else np.zeros((0, 2))
        uv = torch.empty((n, 2), dtype=torch.float64, device=self.device)
        cam = torch.empty(n, dtype=torch.int64, device=self.device)
        self._check(self.lib.pcacc_project_cameras(
            self.h, _ptr(pts), n, stride, g.ctypes.data_as(C.c_void_p), inv.ctypes.data_as(C.c_void_p),
            K.ctypes.data_as(C.c_void_p), wh.ctypes.data_as(C.c_void_p), nc, float(depth_thres),
            _ptr(uv), _ptr(cam), _stream()))
        self.stage.fence()
        return uv, cam

    def pts_feat_from_img(self, pts_uv, img, method='bilinear'):
        """datasets/nuscenes_utils.py:181-214 on the device -> (n, C) float64 CUDA tensor ((n,) for a
        2-D image, as the reference's fancy indexing gives)."""
        assert method in ('bilinear', 'nearest'), f'{method} is not supported'
        uv = self.stage.put('feat_uv', pts_uv if isinstance(pts_uv, torch.Tensor)
                            else np.asarray(pts_uv, dtype=np.float64)).contiguous()
        assert uv.dtype == torch.float64 and uv.dim() == 2 and uv.shape[1] == 2
        im = img if isinstance(img, torch.Tensor) else np.ascontiguousarray(img)
        codes = {torch.uint8: _lib.SEM_U8, torch.int16: _lib.SEM_I16, torch.int32: _lib.SEM_I32,
                 torch.int64: _lib.SEM_I64, torch.float32: _lib.IMG_F32, torch.float64: _lib.IMG_F64}
        imd = self.stage.put('feat_img', im).contiguous()
        if imd.dtype not in codes:
            imd = imd.double()
        two_d = imd.dim() == 2
        h, w = int(imd.shape[0]), int(imd.shape[1])
        c = 1 if two_d else int(imd.shape[2])
        n = int(uv.shape[0])
        out = torch.empty((n, c), dtype=torch.float64, device=self.device)
        self._check(self.lib.pcacc_pts_feat_from_img(self.h, _ptr(uv), n, _ptr(imd), codes[imd.dtype], h, w,
                                                     c, 1 if method == 'bilinear' else 0, _ptr(out),
                                                     _stream()))
        self.stage.fence()
        self._dirty = True
        return out[:, 0] if two_d else out

    def static_obj_partitioning(self, pc, P: int, elev_thresh: float):
        """bev_generator/sem_bev.py:556-591 on the device.  pc: (n,10) float64 rows with grid
        coordinates in columns 0, 1 -> (pc with column 8 flagged, elevmap (P,P) float64, observed
        mask (P,P) bool) as CUDA tensors."""
        t = pc if isinstance(pc, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(pc, dtype=np.float64))
        t = t.to(self.device).contiguous().clone()
        assert t.dtype == torch.float64 and t.dim() == 2 and t.shape[1] == 10
        elev = torch.empty((P, P), dtype=torch.float64, device=self.device)
        obs = torch.empty((P, P), dtype=torch.uint8, device=self.device)
        scratch = torch.empty(P * P, dtype=torch.int64, device=self.device)
        self._check(self.lib.pcacc_static_obj_partitioning(self.h, _ptr(t), int(t.shape[0]), int(P),
                                                           float(elev_thresh), _ptr(elev), _ptr(obs),
                                                           _ptr(scratch), _stream()))
        self._dirty = True
        return t, elev, obs.bool()

    # -- the rasteriser's steps as stand-alone operators (include/pcacc.h) ----------------
    def _rows64(self, pc, min_cols=3):
        """(n, cols) float64 contiguous CUDA tensor of a numpy array / tensor (no copy when it is one)."""
        t = pc if isinstance(pc, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(pc, dtype=np.float64))
        t = t.to(device=self.device, dtype=torch.float64).contiguous()
        assert t.dim() == 2 and t.shape[1] >= min_cols, f'expected (n, >={min_cols}) rows'
        return t

    def elevation_map(self, pc, P: int):
        """get_elevation_map, bev_generator/sem_bev.py:535-553 -> (elevmap (P,P) float64, observed
        mask (P,P) bool) CUDA tensors."""
        t = self._rows64(pc)
        elev = torch.empty((P, P), dtype=torch.float64, device=self.device)
        obs = torch.empty((P, P), dtype=torch.uint8, device=self.device)
        scratch = torch.empty(P * P, dtype=torch.int64, device=self.device)
        self._check(self.lib.pcacc_elevation_map(self.h, _ptr(t), int(t.shape[0]), int(t.shape[1]), int(P),
                                                 _ptr(elev), _ptr(obs), _ptr(scratch), _stream()))
        self._dirty = True
        return elev, obs.bool()

    def velo2frame(self, pc_velo, P_velo_frame):
        """sem_pc_accum.py:347-366 -> (n,3) float64 CUDA tensor."""
        t = pc_velo if isinstance(pc_velo, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(pc_velo))
        if t.dtype not in (torch.float32, torch.float64):
            t = t.double()
        t = t.to(self.device).contiguous()
        assert t.dim() == 2 and t.shape[1] >= 3
        n = int(t.shape[0])
        out = torch.empty((n, 3), dtype=torch.float64, device=self.device)
        Pm = _hostd(P_velo_frame, 12)
        self._check(self.lib.pcacc_velo2frame(self.h, _ptr(t), int(t.dtype == torch.float64), n,
                                              int(t.shape[1]), Pm.ctypes.data_as(C.c_void_p), _ptr(out),
                                              _stream()))
        return out

    def preprocess_pc(self, pc, rot=None, trans_dx=0., trans_dy=0., crop_view=None, height_filter=None,
                      grid_view=None, P=0):
        """bev_generator/bev_generator.py:127-160, 207-256, 737-747 for a cloud: rotate + translate
        (rot (3,3) or None), crop, height filter, pos2grid; None switches a step off.  -> kept rows,
        input order, CUDA tensor."""
        t = self._rows64(pc)
        n, cols = int(t.shape[0]), int(t.shape[1])
        out = torch.empty_like(t)
        nk = torch.zeros(1, dtype=torch.int64, device=self.device)
        nan = float('nan')
        R = None if rot is None else _hostd(rot, 9)
        self._check(self.lib.pcacc_preprocess_pc(
            self.h, _ptr(t), n, cols, None if R is None else R.ctypes.data_as(C.c_void_p),
            float(trans_dx), float(trans_dy), nan if crop_view is None else float(crop_view),
            nan if height_filter is None else float(height_filter),
            nan if grid_view is None else float(grid_view), int(P), _ptr(out), _ptr(nk), _stream()))
        return out[:int(nk.item())]

    def cell_stats(self, pc, P: int, sems=None, sem_col=7, weight_col=-1, weights=None, finish=0,
                   want=('sel', 'rest', 'wsum')):
        """pcacc_cell_stats: per-cell counts of the selected / other points and the selected points'
        weight sum -> dict of (P,P) float64 CUDA tensors for the names in `want`."""
        t = self._rows64(pc, 2)
        n, cols = int(t.shape[0]), int(t.shape[1])
        maps = {k: torch.empty((P, P), dtype=torch.float64, device=self.device) for k in want}
        if sems is None:
            sarr, ns = None, -1
        else:
            sarr = np.ascontiguousarray(np.asarray(list(sems), dtype=np.int32))
            ns = int(sarr.shape[0])
        w = None
        if weights is not None:
            w = weights if isinstance(weights, torch.Tensor) else torch.from_numpy(
                np.ascontiguousarray(weights, dtype=np.float64))
            w = w.to(device=self.device, dtype=torch.float64).contiguous()
            assert w.numel() == n, 'one weight per point'
        self._check(self.lib.pcacc_cell_stats(
            self.h, _ptr(t), n, cols, int(P), int(sem_col),
            None if sarr is None or ns == 0 else sarr.ctypes.data_as(C.c_void_p), ns, int(weight_col),
            _ptr(w), int(finish), _ptr(maps.get('sel')), _ptr(maps.get('rest')), _ptr(maps.get('wsum')),
            _stream()))
        return maps

    def partition_semantic_pc(self, pc, sems, sem_col: int):
        """bev_generator/bev_generator.py:411-432 -> (selected rows, other rows) CUDA tensors."""
        t = self._rows64(pc, 1)
        n, cols = int(t.shape[0]), int(t.shape[1])
        a, b = torch.empty_like(t), torch.empty_like(t)
        nk = torch.zeros(1, dtype=torch.int64, device=self.device)
        sarr = np.ascontiguousarray(np.asarray(list(sems), dtype=np.int32))
        self._check(self.lib.pcacc_partition_semantic_pc(
            self.h, _ptr(t), n, cols, int(sem_col),
            sarr.ctypes.data_as(C.c_void_p) if sarr.size else None, int(sarr.size), _ptr(a), _ptr(b),
            _ptr(nk), _stream()))
        k = int(nk.item())
        return a[:k], b[:n - k]

    def dirichlet_expectation(self, gridmaps, obs_weight=1.):
        """bev_generator/bev_generator.py:455-481 -> (G, ...) float64 CUDA tensor."""
        if isinstance(gridmaps, torch.Tensor):
            m = gridmaps.to(device=self.device, dtype=torch.float64).contiguous().clone()
        else:
            m = torch.from_numpy(np.ascontiguousarray(np.stack(gridmaps), dtype=np.float64)).to(self.device)
        G = int(m.shape[0])
        self._check(self.lib.pcacc_dirichlet_expectation(self.h, _ptr(m), G, int(m.numel() // max(G, 1)),
                                                         float(obs_weight), _stream()))
        return m

    def road_marking(self, values, int_scaler=1., int_sep_scaler=1., int_mid_threshold=0.5,
                     sigmoid_only=False):
        """sem_bev.py:593-617 elementwise -> float64 CUDA tensor of the input's shape."""
        v = values if isinstance(values, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(values, dtype=np.float64))
        v = v.to(device=self.device, dtype=torch.float64).contiguous()
        out = torch.empty_like(v)
        self._check(self.lib.pcacc_road_marking(self.h, _ptr(v), int(v.numel()), float(int_scaler),
                                                float(int_sep_scaler), float(int_mid_threshold),
                                                int(bool(sigmoid_only)), _ptr(out), _stream()))
        return out

    def raster_stats(self):
        s = (C.c_int64 * 3)()
        self._check(self.lib.pcacc_raster_stats(self.h, C.byref(s), _stream()))
        return {'visited': int(s[0]), 'binned': int(s[1]), 'replays': int(s[2])}


# numpy view of pcacc_bev_params (include/pcacc.h): V parameter blocks are filled with a dozen
# vectorised assignments instead of ~30 ctypes field stores per variant
BEV_DTYPE = np.dtype({
    'names': ['frame_begin', 'frame_split', 'frame_end', 'origin', 'R', 'trans_dx', 'trans_dy', 'view',
              'height_filter', 'int_scaler', 'int_sep_scaler', 'int_mid_threshold', 'rgb_fill',
              'road_cls', 'veh_cls', 'elevation_max'],
    'formats': ['<i8', '<i8', '<i8', ('<f8', 3), ('<f8', 9), '<f8', '<f8', '<f8', '<f8', '<f8', '<f8',
                '<f8', '<f8', '<i4', ('<i4', 4), '<i4'],
    'offsets': [getattr(BevParams, n).offset for n in (
        'frame_begin', 'frame_split', 'frame_end', 'origin', 'R', 'trans_dx', 'trans_dy', 'view',
        'height_filter', 'int_scaler', 'int_sep_scaler', 'int_mid_threshold', 'rgb_fill', 'road_cls',
        'veh_cls', 'elevation_max')],
    'itemsize': C.sizeof(BevParams)})


def make_bev_params_batch(frame_begin, frame_split, frame_end, origin, Rs, trans_dx, trans_dy, views,
                          height_filter, int_scaler, int_sep_scaler, int_mid_threshold, rgb_fill,
                          sem_idxs, elevation_max=False):
    """V parameter blocks that share the window and generator settings and differ in rotation
    (Rs: (V,3,3)), translation and view.  Returns a ctypes array (BevParams * V) over a numpy
    record array (kept alive by the ctypes object)."""
    Rs = np.asarray(Rs, dtype=np.float64).reshape(-1, 9)
    V = Rs.shape[0]
    rec = np.zeros(V, dtype=BEV_DTYPE)
    rec['frame_begin'], rec['frame_split'], rec['frame_end'] = int(frame_begin), int(frame_split), int(frame_end)
    rec['origin'] = np.asarray(origin, dtype=np.float64).reshape(3)
    rec['R'] = Rs
    rec['trans_dx'], rec['trans_dy'], rec['view'] = trans_dx, trans_dy, views
    rec['height_filter'] = np.nan if height_filter is None else float(height_filter)
    rec['int_scaler'], rec['int_sep_scaler'] = float(int_scaler), float(int_sep_scaler)
    rec['int_mid_threshold'], rec['rgb_fill'] = float(int_mid_threshold), float(rgb_fill)
    rec['road_cls'] = int(sem_idxs['road'])
    rec['veh_cls'] = [int(sem_idxs[k]) for k in ('car', 'truck', 'bus', 'motorcycle')]
    rec['elevation_max'] = 1 if elevation_max else 0
    return (BevParams * V).from_buffer(rec)


def make_bev_params(frame_begin, frame_split, frame_end, origin, R, trans_dx, trans_dy, view,
                    height_filter, int_scaler, int_sep_scaler, int_mid_threshold, rgb_fill,
                    sem_idxs, elevation_max=False) -> BevParams:
    p = BevParams()
    p.frame_begin, p.frame_split, p.frame_end = int(frame_begin), int(frame_split), int(frame_end)
    o = np.asarray(origin, dtype=np.float64).reshape(3)
    Rm = np.asarray(R, dtype=np.float64).reshape(9)
    for k in range(3):
        p.origin[k] = float(o[k])
    for k in range(9):
        p.R[k] = float(Rm[k])
    p.trans_dx, p.trans_dy = float(trans_dx), float(trans_dy)
    p.view = float(view)
    p.height_filter = float('nan') if height_filter is None else float(height_filter)
    p.int_scaler, p.int_sep_scaler = float(int_scaler), float(int_sep_scaler)
    p.int_mid_threshold = float(int_mid_threshold)
    p.rgb_fill = float(rgb_fill)
    p.road_cls = int(sem_idxs['road'])
    for k, name in enumerate(('car', 'truck', 'bus', 'motorcycle')):
        p.veh_cls[k] = int(sem_idxs[name])
    p.elevation_max = 1 if elevation_max else 0
    return p
