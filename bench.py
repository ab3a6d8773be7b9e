#!/usr/bin/env python
"""bench.py — the contract benchmark of the fusion + BEV rasterisation path.

    python bench.py --gpus N --steps K --warmup W            # this repo (libpcacc, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the path on the host CPU

Workload (BASELINE.json configs[1] shape, batched as configs[3]): nuScenes-shaped
oracle-pose scenes — 40 sweeps x 34,688 points, CAM_FRONT 1600x900, seeded
synthetic class maps standing in for the ONNX network — each scene accumulated
and rasterised into 32 BEVs (8 present indices x 4 augmentation variants,
256x256, present / future / full).  A *step* is SCENES_PER_STEP such scenes on
every GPU; scenes are sharded by rank with no data-path collective (weak
scaling).  metric = input lidar points fused + rasterised per second.

value : libpcacc's C ABI driven with inputs already resident in HBM.
e2e   : the reference-facing Python API (accumulator.integrate / generate_bev)
        with HOST numpy buffers in and out, copies inside the timed region.
Extra lines in the same JSON object: roofline (dominant kernel, CUDA-event
timed inside the timed region), roofline_path, cpu_baseline, clocks, and (N=1)
the long-horizon KITTI-360 window of BASELINE.json configs[2].
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pc_accumulation_lib_b200 import parallel, synth  # noqa: E402

N_SWEEPS = 40
PRESENT_IDXS = [6, 10, 14, 18, 22, 26, 30, 34]
BEVS_PER_PRESENT = 4
P = 256
N_DISTINCT = 4            # distinct synthetic scenes per rank (replicated on the device)
METRIC = 'lidar_points_fused_and_rasterised_per_s'
OUT = sys.stdout


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0}, 'fallback'


def bev_setup():
    bp = synth.nusc_bev_params(pixel_size=P)
    bp['max_trans_radius'] = 5.0       # run_nuscenes_bev_gen.py augmentation defaults
    bp['zoom_thresh'] = 0.1
    return bp


def make_scenes(rank, n, world=1):
    """The first n scenes of this rank's shard of the global scene list (scene id -> seed)."""
    ids = parallel.shard_units(n * world, rank, world)
    return [synth.nusc_scene(synth.seed_for(4, sid), N_SWEEPS) for sid in ids]


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every 5 ms from a
    thread (the timed region can be shorter than one `nvidia-smi -lms` period)."""
    REASONS = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20),
               ('sw_power_cap', 0x4))

    def __init__(self, gpu_index, period=0.005):
        self.idx = gpu_index
        self.period = period
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self.thread = None
        self.err = None

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            idx = self.idx
            if vis:
                try:
                    idx = int(vis.split(',')[self.idx])
                except Exception:
                    idx = self.idx
            hdl = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(hdl, pynvml.NVML_CLOCK_SM))
            while not self._stop.is_set():
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(hdl, pynvml.NVML_CLOCK_SM)))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(hdl)
                for name, bit in self.REASONS:
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(self.period)
        except Exception as e:      # pragma: no cover - depends on the box
            self.err = repr(e)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        time.sleep(0.02)

    def stop(self):
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        if not self.sm:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz,
                    'reasons': ['clock sampling unavailable: %s' % self.err]}
        return {'sm_mhz': float(np.median(self.sm)), 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.sm)}


# ---------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------
class HostSemseg:
    """Stand-in for the ONNX network (out of scope): returns the pre-generated
    class map of the image it is handed (utils/onnx_utils.py:32-44 contract)."""

    def __init__(self):
        self.by_id = {}

    def pred(self, rgb):
        return self.by_id[id(rgb)][None, None]


def variant_params(acc, rng):
    """The 32 pcacc_bev_params of one scene, exactly what generate_bev would use."""
    gen = acc.sem_bev_generator
    first, n = acc._fids[0], len(acc._fids)
    out, aug = [], []
    for p in PRESENT_IDXS:
        origin = np.array(acc.poses[p])
        for _ in range(BEVS_PER_PRESENT):
            gen.rng = rng
            a = gen.rand_aug_params()
            out.append(gen._bev_params(first + 0, first + p, first + n, origin, a['rot_ang'],
                                       a['trans_dx'], a['trans_dy'],
                                       a['zoom_scalar'] * gen.view_size))
            aug.append((p, a))
    return out, aug


def run_ours(args, rank, world, local_rank, dist):
    import torch
    from pc_accumulation_lib_b200 import NuScenesOracleSemanticPointCloudAccumulator
    from pc_accumulation_lib_b200.device import DeviceCloud
    dev = torch.device('cuda', local_rank)
    pk, pk_kind = peaks()

    scenes = make_scenes(rank, N_DISTINCT, world)
    n_in_scene = sum(o['pc'].shape[0] for o in scenes[0])
    S = args.scenes_per_step
    bevs_per_scene = len(PRESENT_IDXS) * BEVS_PER_PRESENT

    # ---- e2e objects: public API, host buffers ---------------------------------
    from pc_accumulation_lib_b200.device import pin_observation
    semseg = HostSemseg()
    # the same scenes in page-locked memory (what a dataloader / semseg stand-in that writes
    # into `pinned_empty` buffers hands over): read by the kernels in place, no host copy
    scenes_pin = [[pin_observation(o) for o in sc] for sc in scenes]
    for group in (scenes, scenes_pin):
        for sc in group:
            for o in sc:
                for img, cls in zip(o['images'], o['_semseg']):
                    semseg.by_id[id(img)] = cls

    def new_acc():
        a = NuScenesOracleSemanticPointCloudAccumulator(
            semseg, synth.NUSC_FILTERS, synth.SEM_IDXS, None, bev_setup(),
            ring_capacity_pts=n_in_scene + 4096, ring_max_frames=N_SWEEPS + 8, device=local_rank)
        # data-error flags are checked at the next synchronising call (generate_bev) instead of
        # after every sweep, so the host runs ahead of the kernels
        a.sync_each_integrate = False
        return a

    acc = new_acc()
    marks = []          # per distinct scene: per sweep (rel frame ids, inst idx)
    params = []         # per distinct scene: 32 BevParams (frame ids relative to 0)
    augs = []           # per distinct scene: the (present_idx, augmentation) of every BEV
    n_keep = []

    def e2e_scene(acc, sc, k, record=False):
        """One scene through the reference-facing API; returns #BEVs produced."""
        acc.reset()
        log = []
        if record:
            orig = acc.cloud.mark_dynamic
            acc.cloud.mark_dynamic = lambda f, i: (log.append((list(f), list(i))), orig(f, i))[1]
        for o in sc:
            acc.integrate([o])
        if record:
            acc.cloud.mark_dynamic = orig
            first = acc._fids[0]
            marks.append([([f - first for f in fs], ii) for fs, ii in log])
            rng = np.random.RandomState(1234 + k)
            ps, aa = variant_params(acc, rng)
            for q in ps:
                q.frame_begin -= first
                q.frame_split -= first
                q.frame_end -= first
            params.append(ps)
            augs.append(aa)
            acc._sync()                      # exact kept counts (unsynced: the upper bound n_in)
            n_keep.append(acc.cloud.resident_points())
        n = 0
        acc.sem_bev_generator.rng = np.random.RandomState(99 + k)
        for p in PRESENT_IDXS:
            bevs = acc.generate_bev(p, BEVS_PER_PRESENT, True)
            n += len(bevs)
            assert bevs[0]['rgb_full'].shape == (3, P, P)
        return n

    for k in range(N_DISTINCT):            # also the e2e warm-up (pageable -> sparse staging)
        assert e2e_scene(acc, scenes[k], k, record=True) == bevs_per_scene
    staging_pageable = acc.cloud.last_staging

    # ---- device-resident copies of the inputs (value path) ------------------------
    def to_dev(sc):
        out = []
        for o in sc:
            out.append(dict(
                pc=torch.from_numpy(o['pc']).to(dev), cam=torch.from_numpy(o['pc_cam_idx']).to(dev),
                rgb=[torch.from_numpy(np.ascontiguousarray(i)).to(dev) for i in o['images']],
                sem=[torch.from_numpy(c.astype(np.uint8)).to(dev) for c in o['_semseg']],
                T=None))
        T_gw = np.linalg.inv(sc[0]['ego_at_lidar_ts'])
        for o, d in zip(sc, out):
            d['T'] = T_gw @ o['ego_at_lidar_ts']
        return out

    dev_scenes = []
    for s in range(S):                      # S resident copies: the step's inputs exceed L2
        dev_scenes.append(to_dev(scenes[s % N_DISTINCT]))
    # scenes are independent: they alternate between n_streams accumulators, each on its own
    # CUDA stream, so one scene's launch gaps and low-occupancy tails overlap another's work
    n_str = max(1, args.streams)
    clouds = [DeviceCloud(n_in_scene + 8 * N_SWEEPS + 4096, N_SWEEPS + 8, local_rank)
              for _ in range(n_str)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_str)]
    outs = [torch.empty((bevs_per_scene, 3, 7, P, P), dtype=torch.float16, device=dev)
            for _ in range(n_str)]

    def device_step():
        main = torch.cuda.current_stream()
        for st in streams:
            st.wait_stream(main)
        for s in range(S):
            with torch.cuda.stream(streams[s % n_str]):
                scene_pass(s, clouds[s % n_str], outs[s % n_str])
        for st in streams:
            main.wait_stream(st)
        for c in clouds:
            c._keep.clear()

    # argument blocks that do not change from step to step are built once
    preps = [[c.prepare_records_batch(dev_scenes[s], synth.NUSC_FILTERS, 255.) if s % n_str == ci
              else None for s in range(S)] for ci, c in enumerate(clouds)]
    from pc_accumulation_lib_b200._lib import BevParams
    mark_rel = [(np.array([f for fs, _ in m for f in fs], dtype=np.int64),
                 np.ascontiguousarray(np.array([i for _, ii in m for i in ii], dtype=np.int32)))
                for m in marks]
    par_arr = [(BevParams * bevs_per_scene)(*ps) for ps in params]
    from pc_accumulation_lib_b200.device import BEV_DTYPE
    par_rec = [np.frombuffer(a, dtype=BEV_DTYPE) for a in par_arr]
    par_rel = [tuple(np.array([getattr(q, f) for q in ps], dtype=np.int64)
                     for f in ('frame_begin', 'frame_split', 'frame_end')) for ps in params]

    def scene_pass(s, cloud, out_planes):
        k = s % N_DISTINCT
        cloud.reset()
        # all 40 sweeps of the scene in one launch (pcacc_integrate_records_batch)
        first = cloud.integrate_prepared(preps[s % n_str][s])
        fr, ii = mark_rel[k]
        cloud.mark_dynamic_now(np.ascontiguousarray(fr + first), ii)
        arr = par_arr[k]
        rec, rel = par_rec[k], par_rel[k]          # numpy view of the same 32 parameter blocks
        rec['frame_begin'] = rel[0] + first
        rec['frame_split'] = rel[1] + first
        rec['frame_end'] = rel[2] + first
        cloud.rasterise(arr, P, out=out_planes)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
    # host time to ENQUEUE a step (no synchronisation inside): close to ms_per_step means the
    # device path is bound by the launching thread as much as by the GPU.  (Enqueuing from 2-8
    # host threads was measured: 2.28-2.62 ms per step against 2.07 ms from one thread.)
    t_h0 = time.perf_counter()
    for _ in range(3):
        device_step()
    host_enqueue_ms = (time.perf_counter() - t_h0) / 3 * 1e3
    barrier()
    if args.ncu_step:
        # profiling aid (never a bench value): one step between cudaProfilerStart/Stop, for
        #   ncu --profile-from-start off ... python bench.py --ncu-step --streams 1 --scenes-per-step 1
        torch.cuda.profiler.start()
        device_step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        if rank == 0:
            OUT.write(json.dumps({'ncu_step': True, 'scenes': S, 'streams': n_str}) + '\n')
            OUT.flush()
        return
    # Per-kernel device times.  (1) every kernel class timed ALONE: CUDA events around each launch
    # on ONE stream, untimed passes after the warm-up — this picks the dominant stage and gives its
    # own duration.  (2) inside the timed region only the dominant stage's launches carry events
    # (two cudaEventRecord per launch: timing every class there costs host time the throughput
    # measurement would pay for); with several streams that in-region duration also spans the other
    # streams' kernels sharing the SMs, so it is reported beside the alone figure, not instead.
    STAGES = {'bev_reduce': ('bev_reduce', 'bev_reduce_big'), 'bev_bin': ('bev_bin', 'bev_classify'),
              'integrate': ('integrate',), 'scan': ('scan',), 'bev_scatter': ('bev_scatter',),
              'mark_dynamic': ('mark_dynamic',)}

    def by_stage(prof):
        return {st: (sum(prof[k][0] for k in ks), prof[ks[0]][1]) for st, ks in STAGES.items()}

    torch.cuda.synchronize()
    clouds[0].profile(True)
    clouds[0].profile_read()
    own = [s for s in range(S) if s % n_str == 0]
    n_alone = 3
    for _ in range(n_alone):
        for s in own:
            scene_pass(s, clouds[0], outs[0])
        torch.cuda.synchronize()
        clouds[0]._keep.clear()
    alone_raw = clouds[0].profile_read()
    clouds[0].profile(False)
    alone = by_stage(alone_raw)
    dom = max(alone, key=lambda k: alone[k][0])
    for c in clouds:
        c.profile(STAGES[dom])
        c.profile_read()
    # rank 0 samples its GPU (its line is the one printed); eight ranks polling NVML every 5 ms
    # contend for the driver's locks with the kernel launches they are supposed to observe
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        device_step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    prof = {}
    for c in clouds:
        for kk, vv in c.profile_read().items():
            prof[kk] = (prof.get(kk, (0., 0))[0] + vv[0], prof.get(kk, (0, 0))[1] + vv[1])
        c.profile(False)
    # NVML answers in ~10-50 ms: a timed region of a few tens of ms yields one or two samples.
    # Keep the same load running (untimed, not counted) until there are at least 5.
    extended = 0
    while rank == 0 and len(clocks.sm) < 5 and extended < 200:
        device_step()
        torch.cuda.synchronize()
        extended += 1
    clk = clocks.stop()
    clk['untimed_steps_added_for_sampling'] = extended
    # the only collective of the run: summary statistics (sum of work, max of time)
    pts_step_rank = S * n_in_scene
    tot = parallel.reduce_stats(dist, {'points': pts_step_rank * args.steps,
                                       'bevs': S * bevs_per_scene * args.steps,
                                       'ms_max': ms_total}, device=dev)
    ms_max = tot['ms_max']
    value = tot['points'] / (ms_max * 1e-3)
    bevs_per_s = tot['bevs'] / (ms_max * 1e-3)

    # ---- roofline of the dominant kernel (CUDA events inside the timed region) ------
    n_res = float(np.mean([n_keep[s % N_DISTINCT] for s in range(S)]))
    n_vis = float(np.mean([sum(int((o['pc_cam_idx'] >= 0).sum()) for o in scenes[s % N_DISTINCT])
                           for s in range(S)])) / N_SWEEPS
    alg = {  # algorithmic bytes per launch, SURVEY.md §8(d) / DESIGN.md "Bytes"
        # one launch integrates the scene's 40 sweeps
        'integrate': 49.0 * n_in_scene + 7.0 * n_vis * N_SWEEPS + 37.0 * n_res,
        'bev_bin': 33.0 * n_res * bevs_per_scene,
        'bev_reduce': 42.0 * P * P * bevs_per_scene,
    }
    # stages made of two kernels are timed as one stage: reduce = small-cell pass + queued
    # large-cell pass; bin = streaming crop test + exact per-candidate pass
    region = by_stage(prof)
    n_passes = n_alone * len(own)
    peak = pk['hbm_gbs']
    us_alone = alone[dom][0] / n_passes * 1e3
    us_region = region[dom][0] / max(S * args.steps, 1) * 1e3
    ach = alg.get(dom, 0.0) / (us_alone * 1e-6) / 1e9 if us_alone > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get(dom)
    except Exception:
        pass
    kname = {'bev_bin': 'k_bev_classify+k_bev_bin', 'bev_reduce': 'k_bev_reduce_chunk+k_bev_reduce_big',
             'integrate': 'k_integrate_records_batch'}.get(dom, 'k_' + dom)
    alone_sum = max(sum(v[0] for v in alone.values()), 1e-9)
    roofline = {'bound': 'hbm', 'kernel': kname, 'achieved': ach, 'peak': peak, 'unit': 'GB/s',
                'frac': ach / peak, 'traffic': traffic, 'peak_source': pk_kind + ' (burst copy)',
                'launch_us': us_alone, 'launches_timed': n_passes,
                'algorithmic_bytes_per_launch': alg.get(dom, 0.0),
                'share_of_step': alone[dom][0] / alone_sum,
                'stages': {st: {'launch_us': round(v[0] / n_passes * 1e3, 2),
                                'algorithmic_bytes_per_launch': alg.get(st),
                                'frac': (alg[st] / (v[0] / n_passes * 1e-3) / 1e9 / peak) if st in alg and v[0] > 0 else None}
                           for st, v in alone.items()},
                'kernel_us_alone': {k: round(v[0] / n_passes * 1e3, 2) for k, v in alone_raw.items() if v[1]},
                'kernel_launches_per_step': {k: int(v[1] / n_passes * S) for k, v in alone_raw.items() if v[1]},
                'timed_region_overlapped': {
                    'launch_us': us_region, 'launches_timed': S * args.steps, 'streams': n_str,
                    'achieved': alg.get(dom, 0.0) / (us_region * 1e-6) / 1e9 if us_region > 0 else None,
                    'frac': alg.get(dom, 0.0) / (us_region * 1e-6) / 1e9 / peak if us_region > 0 else None},
                'timing': ('launch_us: CUDA events around the stage\'s launches on ONE stream, %d scene passes '
                           'right before the timed region; timed_region_overlapped: events around the same '
                           'launches inside the timed region, where %d streams share the SMs' % (n_passes, n_str))}
    b_step = S * (alg['integrate'] + alg['bev_bin'] + alg['bev_reduce'])
    path_ach = b_step * args.steps / (ms_total * 1e-3) / 1e9
    roofline_path = {'achieved': path_ach, 'peak': peak, 'unit': 'GB/s', 'frac': path_ach / peak,
                     'frac_of_nominal_8TBps': path_ach / 8000.0,
                     'bytes_per_step': b_step,
                     'resident_points_per_scene': n_res,
                     'note': 'all algorithmic bytes of the step (integrate + rasterise, SURVEY.md 8d: '
                             '49 N_in + 7 N_vis + 37 N_keep per sweep, 33 N_res + 42 P^2 per BEV) / step time'}
    gpu_launches = int(sum(v[1] for v in prof.values()))   # launch counts cover every class, timed or not

    # ---- e2e: public API, host buffers, copies inside the timed region ---------------
    # Scenes are independent (one accumulator per scene, run_nuscenes_bev_gen.py:165,203), so the
    # generation loop may keep several in flight: `--e2e-threads` Python threads, each with its
    # own accumulator and CUDA stream, take the step's scenes round-robin; while one waits for
    # its BEV planes to arrive in host memory another runs its integrate() calls.
    e2e_scenes = max(1, min(S, args.e2e_scenes))
    n_thr = max(1, min(args.e2e_threads, e2e_scenes))
    accs = [acc] + [new_acc() for _ in range(n_thr - 1)]
    e2e_streams = [torch.cuda.Stream(device=dev) for _ in range(n_thr)]

    def e2e_leg(src, steps):
        counts = [0] * n_thr
        errors = []

        def worker(t):
            try:
                torch.cuda.set_device(local_rank)
                with torch.cuda.stream(e2e_streams[t]):
                    for _ in range(steps):
                        for s_ in range(t, e2e_scenes, n_thr):
                            counts[t] += e2e_scene(accs[t], src[s_ % N_DISTINCT], s_ % N_DISTINCT)
                    e2e_streams[t].synchronize()
            except BaseException as e:      # surfaced by the caller
                errors.append(e)

        barrier()
        t0 = time.perf_counter()
        if n_thr == 1:
            worker(0)
        else:
            ths = [threading.Thread(target=worker, args=(t,)) for t in range(n_thr)]
            for th in ths:
                th.start()
            for th in ths:
                th.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if errors:
            raise errors[0]
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            sys.stderr.write('rank %d: e2e leg %.1f ms for %d scenes\n' % (rank, dt * 1e3, steps * e2e_scenes))
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()), sum(counts)

    if n_thr > 1:
        # a thread coming back from a blocking C call (event wait, launch) gets the GIL only when
        # the running thread reaches its switch interval: the default 5 ms is longer than a whole
        # generate_bev call
        sys.setswitchinterval(args.switch_interval)
    e2e_leg(scenes_pin, 1)                                   # warm-up of the pinned path and threads
    dt, n_b = e2e_leg(scenes_pin, args.e2e_steps)
    staging_pinned = acc.cloud.last_staging
    e2e_leg(scenes, 1)                                       # warm-up: every thread's staging slots exist
    dt_pg, n_b_pg = e2e_leg(scenes, max(1, args.e2e_steps // 2))
    sc0 = scenes[0]
    host_in = sum(o['pc'].nbytes + o['pc_cam_idx'].nbytes + sum(i.nbytes for i in o['images'])
                  + sum(c.nbytes for c in o['_semseg']) for o in sc0) * e2e_scenes
    n_vis_scene = sum(int((o['pc_cam_idx'] >= 0).sum()) for o in sc0)
    # bytes that cross the bus towards the device per step.  Pinned arrays are read in place:
    # cam_idx of every point (8 B), the (u, v) sector of every visible point (32 B), the
    # sectors holding x, y, z, intensity, inst of every kept point (64 B) and the class / rgb
    # sectors of the visible / kept pixels (32 B each) — counted in 32-byte sectors, the
    # granularity of a zero-copy read.
    h2d_direct = e2e_scenes * (8 * n_in_scene + 64 * n_vis_scene + 96 * int(np.mean(n_keep)))
    h2d_sparse = e2e_scenes * 64 * n_vis_scene            # packed rows + samples of the visible points
    d2h = e2e_scenes * bevs_per_scene * 21 * P * P * 2
    stage_name = {1: 'direct (pinned arrays read in place)', 2: 'sparse (visible rows + samples packed into a pinned slot)'}
    e2e = {'value': world * e2e_scenes * args.e2e_steps * n_in_scene / dt, 'unit': 'points/s',
           'h2d_bytes_per_step': int(h2d_direct), 'd2h_bytes_per_step': int(d2h),
           'host_input_bytes_per_step': int(host_in),
           'bevs_per_s': n_b * world / dt, 'scenes_per_step': e2e_scenes, 'steps': args.e2e_steps,
           'threads': n_thr, 'gil_switch_interval_s': args.switch_interval if n_thr > 1 else None, 'staging': stage_name.get(staging_pinned, str(staging_pinned)),
           'api': 'NuScenesOracleSemanticPointCloudAccumulator.integrate / generate_bev: numpy arrays in '
                  'page-locked host memory in, numpy float16 planes out (sync_each_integrate=False: error '
                  'flags checked at generate_bev)',
           'h2d_note': 'inputs live in pinned host memory and are fetched by the integrate kernel over the '
                       'bus inside the timed region (zero-copy): h2d_bytes_per_step counts the 32-byte sectors '
                       'it touches; host_input_bytes_per_step is the size of the arrays handed to the API',
           'pageable_inputs': {
               'value': world * e2e_scenes * max(1, args.e2e_steps // 2) * n_in_scene / dt_pg,
               'unit': 'points/s', 'bevs_per_s': n_b_pg * world / dt_pg,
               'h2d_bytes_per_step': int(h2d_sparse),
               'staging': stage_name.get(staging_pageable, str(staging_pageable)),
               'note': 'the same calls with ordinary (pageable) numpy arrays'}}

    # ---- CPU baseline + parity gate + extras (rank 0, N = 1) ---------------------------
    cpu = None
    parity = None
    extra = {}
    if rank == 0 and world == 1:
        if not args.no_cpu_baseline:
            # the CPU port runs bench scene 0 with the SAME 32 (present_idx, augmentation) pairs
            # as the device: its BEVs are the reference values of the parity gate
            cpu, cpu_bevs, cpu_acc = cpu_baseline_sample(scenes[0], augs[0])
            parity = parity_gate(torch, clouds[0], scene_pass, outs[0], par_arr[0], cpu_bevs, cpu_acc, augs[0])
        if not args.no_c3:
            del dev_scenes
            outs.clear()
            for c in clouds:
                c.close()
            torch.cuda.empty_cache()
            extra['kitti360_long_horizon'] = c3_extra(torch, DeviceCloud, pk)
            # configs[4]: 1024x1024, elevation max, 100 frames (12 M resident points)
            extra['highres_1024'] = c3_extra(torch, DeviceCloud, pk, F=100, P=1024, elevation_max=True)
            extra['input_side'] = input_side_extra(torch, DeviceCloud, pk)
            extra['kitti360_sequence'] = kitti_seq_extra(torch)
            extra['kitti360_prob_maps'] = prob_map_extra(torch, DeviceCloud, pk)
            extra['writer'] = writer_extra(torch)

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': 'points/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_max / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': bench_config(S, n_str),
            'host_enqueue_ms_per_step': host_enqueue_ms,
            'bevs_per_s': bevs_per_s, 'e2e': e2e, 'gpu_launches': gpu_launches,
            'roofline': roofline, 'roofline_path': roofline_path, 'clocks': clk,
            'parity': parity,
        }
        if cpu is not None:
            line['cpu_baseline'] = cpu
        if extra:
            line['extra'] = extra
        OUT.write(json.dumps(line) + '\n')
        OUT.flush()


def c3_extra(torch, DeviceCloud, pk, F=200, P=P, elevation_max=False):
    """BASELINE.json configs[2]: 200 all-points KITTI-360-shaped frames (24 M
    resident points, use_gt_sem path, lazy re-base), one 256x256 BEV at frame
    100 — the primary roofline configuration of SURVEY.md §8(d)."""
    from pc_accumulation_lib_b200.device import make_bev_params
    n_distinct = 8
    pcs = [torch.from_numpy(synth.kitti_lidar(synth.seed_for(3, f))).cuda() for f in range(n_distinct)]
    sgs = [torch.from_numpy(synth.kitti_sem_gt(synth.seed_for(3, f), pcs[0].shape[0])[:, 0].copy()).cuda()
           for f in range(n_distinct)]
    Ts = [synth.kitti_step_transform(synth.seed_for(3, f)) for f in range(F)]
    N = int(pcs[0].shape[0])
    cloud = DeviceCloud(F * N + 1024, F + 8)
    poses = []
    for f in range(F):
        if f:
            poses = [list(np.matmul(Ts[f], np.array([p + [1]]).T)[:, 0][:-1]) for p in poses]
        poses.append([0., 0., 0.])

    def integrate_all():
        cloud.reset()
        for f in range(F):
            if f:
                cloud.rebase(Ts[f], eager=False)
            cloud.integrate_gt(pcs[f % n_distinct], sgs[f % n_distinct], synth.KITTI_FILTERS)
        cloud._keep.clear()

    def timed(fn, n):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    ms_int = timed(integrate_all, 3)
    # the per-frame loop above is bound by its two host calls per frame (re-base + integrate, ~20 us):
    # the kernels' own time, from CUDA events around every launch
    cloud.profile(('integrate', 'rebase'))
    cloud.profile_read()
    integrate_all()
    kprof = cloud.profile_read()
    cloud.profile(False)
    cloud.sync()
    first, n_live = cloud.live_frames()
    p = F // 2
    origin = np.array(poses[p])
    d = np.array(poses[p - 1]) - np.array(poses[p - 2])
    rot = np.pi - (0.5 * np.pi + np.arctan2(d[1], d[0]))
    R = np.array([[np.cos(rot), -np.sin(rot), 0], [np.sin(rot), np.cos(rot), 0], [0, 0, 1]])
    bp = make_bev_params(first, first + p, first + n_live, origin, R, 0., 0., 80., None, 20., 20., .5,
                         0, synth.SEM_IDXS, elevation_max)
    out = torch.empty((1, 3, 7, P, P), dtype=torch.float16, device='cuda')
    n_res = cloud.resident_points()
    cloud.profile(True)
    cloud.profile_read()
    ms_ras = timed(lambda: cloud.rasterise([bp], P, out=out), 10)
    prof = cloud.profile_read()
    cloud.profile(False)
    st = cloud.raster_stats()
    b_int = 18.0 * N + 37.0 * N
    b_ras = 33.0 * n_res + 42.0 * P * P
    res = {
        'frames': F, 'resident_points': n_res, 'points_in_view': st['binned'],
        'exact_chain_replays': st['replays'],
        'integrate_ms_per_frame': ms_int / F, 'integrate_points_per_s': F * N / (ms_int * 1e-3),
        'integrate_alg_GBps': b_int * F / (ms_int * 1e-3) / 1e9,
        'integrate_kernel_us_per_frame': round(kprof['integrate'][0] / max(kprof['integrate'][1], 1) * 1e3, 2),
        'integrate_kernel_alg_GBps': b_int / (kprof['integrate'][0] / max(kprof['integrate'][1], 1) * 1e-3) / 1e9,
        'rebase_kernel_us_per_frame': round(kprof['rebase'][0] / max(kprof['rebase'][1], 1) * 1e3, 2),
        'rasterise_ms_per_bev': ms_ras, 'bevs_per_s': 1e3 / ms_ras,
        'rasterise_alg_GBps': b_ras / (ms_ras * 1e-3) / 1e9,
        'rasterise_frac_of_hbm_peak': b_ras / (ms_ras * 1e-3) / 1e9 / pk['hbm_gbs'],
        'kernel_us_per_bev': {k: round(v[0] / 11 * 1e3, 1) for k, v in prof.items() if v[1]},
        # against the nominal 8 TB/s as well, and the secondary figure that
        # credits the minimum traffic of the north star's radix sort (+61 B per cropped point
        # at 256^2) — reported beside the primary one, never instead of it
        'rasterise_frac_of_nominal_8TBps': b_ras / (ms_ras * 1e-3) / 1e9 / 8000.0,
        'rasterise_mandated_GBps': (b_ras + st['binned'] * (29 + 16 * int(np.ceil(2 * np.log2(P) / 8))))
                                   / (ms_ras * 1e-3) / 1e9,
        'P': P, 'elevation': 'max' if elevation_max else 'min',
        'note': 'inputs (%.1f GB ring) exceed L2; 33 B/resident point + 42 B/cell algorithmic' % (n_res * 37e-9),
    }
    cloud.close()
    return res


def kitti_seq_extra(torch, F=20, present_idx=10):
    """BASELINE.json configs[0]: the reference's own CPU-runnable case — 20 KITTI-360-shaped
    frames (120,000 points, 1408x376 camera, class map) through
    Kitti360SemanticPointCloudAccumulator.integrate (numpy in) and one 256x256 BEV
    (numpy out), with the ICP pose injected; next to the CPU port on one core."""
    from pc_accumulation_lib_b200 import Kitti360SemanticPointCloudAccumulator
    frames = []
    for f in range(F):
        seed = synth.seed_for(1, f)
        frames.append(dict(pc=synth.kitti_lidar(seed), T=synth.kitti_step_transform(seed),
                           rgb=synth.kitti_rgb(seed), cls=synth.kitti_class_map_fast(seed)))
    n_pts = sum(fr['pc'].shape[0] for fr in frames)
    from pc_accumulation_lib_b200 import pinned_like
    frames_pin = [dict(pc=pinned_like(fr['pc']), T=fr['T'], rgb=pinned_like(fr['rgb']), cls=pinned_like(fr['cls']))
                  for fr in frames]

    def run_gpu(frames=frames):
        acc = Kitti360SemanticPointCloudAccumulator(
            1e9, synth.kitti_calib(), 1.0, synth.FakeSemseg([fr['cls'] for fr in frames]),
            synth.KITTI_FILTERS, synth.SEM_IDXS, False, synth.kitti_bev_params(pixel_size=P),
            ring_capacity_pts=n_pts // 4 + 4096, ring_max_frames=F + 8)
        acc.sync_each_integrate = False
        t0 = time.perf_counter()
        for fr in frames:
            acc.integrate([(fr['rgb'], fr['pc'], None, fr['T'])])
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        bev = acc.generate_bev(present_idx, 1, True)[0]
        t2 = time.perf_counter()
        assert bev['rgb_full'].shape == (3, P, P)
        # the first call of a fresh accumulator allocates the rasteriser's workspace; a sequence calls
        # generate_bev once per present index, so the steady rate is the second call's
        acc.generate_bev(present_idx, 1, True)
        t3 = time.perf_counter()
        n_res = acc.cloud.resident_points()
        acc.cloud.close()
        return t1 - t0, t2 - t1, n_res, t3 - t2

    run_gpu()
    ti, tb, n_res, tb2 = min((run_gpu() for _ in range(3)), key=lambda r: r[0] + r[1])
    run_gpu(frames_pin)
    ti_pin, tb_pin, _, tb2_pin = min((run_gpu(frames_pin) for _ in range(3)), key=lambda r: r[0] + r[1])
    from oracle import oracle as orc          # CPU port, the baseline of this row only
    bp = synth.kitti_bev_params(pixel_size=P)
    gp = dict(sem_idxs=synth.SEM_IDXS, view_size=bp['view_size'], pixel_size=P,
              int_scaler=bp['int_scaler'], int_sep_scaler=bp['int_sep_scaler'],
              int_mid_threshold=bp['int_mid_threshold'], height_filter=bp['height_filter'], rgb_fill=0)
    cpu = orc.KittiOracle(1e9, synth.kitti_calib()['p_velo_frame'], synth.KITTI_FILTERS, gp)
    c0 = time.perf_counter()
    for fr in frames:
        cpu.integrate(fr['pc'], fr['rgb'], fr['cls'], fr['T'])
    c1 = time.perf_counter()
    cpu.generate_bev(present_idx)
    c2 = time.perf_counter()
    return {'frames': F, 'points': n_pts, 'resident_points': n_res,
            'integrate_ms_per_frame': ti / F * 1e3, 'integrate_points_per_s': n_pts / ti,
            'bev_ms': tb * 1e3, 'bev_ms_second_call': tb2 * 1e3,
            'pinned_inputs': {'integrate_ms_per_frame': ti_pin / F * 1e3, 'integrate_points_per_s': n_pts / ti_pin,
                              'bev_ms': tb_pin * 1e3, 'bev_ms_second_call': tb2_pin * 1e3,
                              'note': 'arrays in page-locked memory: no staging copy (camera maps read in place)'},
            'cpu_port_integrate_ms_per_frame': (c1 - c0) / F * 1e3, 'cpu_port_bev_ms': (c2 - c1) * 1e3,
            'cpu_port_cores': 1,
            'note': 'latency-bound by design (SURVEY.md 8d): one 120 k-point frame is 4 MB of traffic; '
                    'the literal reference takes 50-82 ms per frame and 14.6 s per BEV (BASELINE.md)'}


def prob_map_extra(torch, DeviceCloud, pk, n_frames=8):
    """North-star gather variant (sem_pc_accum.py:323-345 with K = 19): the semantic
    input is the (376,1408,19) float32 probability map, the class is its argmax at the projected
    pixel.  120,000-point frames, maps resident in HBM (8 distinct maps = 322 MB > L2).  Rows of 19
    floats (76 B, scalar loads), rows padded to 20 floats (80 B, 16-byte vector loads) and the
    class-index map (1 B) through the same kernel; algorithmic bytes 16 N + (3 + 4K) N_vis + 37 N_keep."""
    P_mat = synth.kitti_calib()['p_velo_frame']
    frames = []
    for f in range(n_frames):
        seed = synth.seed_for(1, f)
        prob = synth.kitti_prob_map(seed)
        frames.append(dict(
            pc=torch.from_numpy(synth.kitti_lidar(seed)).cuda(), rgb=torch.from_numpy(synth.kitti_rgb(seed)).cuda(),
            p19=torch.from_numpy(prob).cuda(),
            p20=torch.from_numpy(np.concatenate([prob, np.zeros(prob.shape[:2] + (1,), np.float32)], axis=2)).cuda(),
            cls=torch.from_numpy(np.argmax(prob, axis=2).astype(np.uint8)).cuda()))
    N = int(frames[0]['pc'].shape[0])
    cloud = DeviceCloud(n_frames * N + 1024, n_frames + 8)

    def run(key):
        cloud.reset()
        for fr in frames:
            cloud.integrate_frustum(fr['pc'], P_mat, fr['rgb'], fr[key], synth.KITTI_FILTERS)
        cloud._keep.clear()

    def timed(key, n=20):
        for _ in range(3):
            run(key)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            run(key)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (n * n_frames) * 1e3          # us per frame

    out = {'points_per_frame': N, 'frames_in_flight': n_frames}
    counts = {}
    for key, K in (('cls', 1), ('p19', 19), ('p20', 20)):
        us = timed(key)
        # the kernel's own duration: CUDA events around each launch (the loop above is bound by
        # the host's call rate, ~15 us per integrate call)
        cloud.profile(('integrate',))
        cloud.profile_read()
        for _ in range(5):
            run(key)
        pr = cloud.profile_read()['integrate']
        cloud.profile(False)
        k_us = pr[0] / max(pr[1], 1) * 1e3
        cloud.sync()
        n_keep = cloud.resident_points() / n_frames
        counts[key] = n_keep
        u, v, m = cloud.project(frames[0]['pc'], P_mat, synth.KITTI_IMG_H, synth.KITTI_IMG_W)
        n_vis = float(m.sum().item())
        gather = 3 + (1 if K == 1 else 4 * K)
        alg = 16.0 * N + gather * n_vis + 37.0 * n_keep
        out[key] = {'K': K, 'us_per_frame_host_call_rate': us, 'kernel_us': k_us,
                    'points_per_s': N / (us * 1e-6), 'n_vis': n_vis,
                    'n_keep': n_keep, 'algorithmic_bytes_per_frame': alg,
                    'kernel_alg_GBps': alg / (k_us * 1e-6) / 1e9,
                    'kernel_frac_of_hbm_peak': alg / (k_us * 1e-6) / 1e9 / pk['hbm_gbs']}
    assert counts['cls'] == counts['p19'] == counts['p20'], counts     # same classes either way
    out['note'] = ('frames back to back on one stream (launch-latency bound: a frame is ~4 MB of traffic); the '
                   'gather touches N_vis rows of 76-80 B of a 40 MB map, i.e. 3.7 % of it')
    cloud.close()
    return out


def writer_extra(torch, n_bevs=32, levels=(1, 6, 9)):
    """write_compressed_pickle (sem_pc_accum.py:280-294) at device rates — the
    threaded writer on this host's cores at gzip levels 1 / 6 / 9, on real BEV dicts of the bench
    workload (2.75 MB of float16 planes each).  BEVs/s next to the e2e generation rate."""
    import shutil
    import tempfile
    from pc_accumulation_lib_b200 import NuScenesOracleSemanticPointCloudAccumulator as Acc
    scene = make_scenes(0, 1)[0]
    acc = Acc(synth.SceneSemseg(), synth.NUSC_FILTERS, synth.SEM_IDXS, None, bev_setup(),
              ring_capacity_pts=sum(o['pc'].shape[0] for o in scene) + 4096, ring_max_frames=N_SWEEPS + 8)
    for o in scene:
        acc.semseg_model.register(o)
        acc.integrate([o])
    acc.sem_bev_generator.rng = np.random.RandomState(3)
    bevs = []
    for p in PRESENT_IDXS:
        bevs += acc.generate_bev(p, BEVS_PER_PRESENT, True)
    bevs = bevs[:n_bevs]
    acc.cloud.close()
    raw = sum(v.nbytes for v in bevs[0].values() if isinstance(v, np.ndarray))
    n_thr = max(1, min(32, os.cpu_count() or 1))
    res = {'threads': n_thr, 'bevs': len(bevs), 'raw_bytes_per_bev': raw, 'levels': {}}
    for lvl in levels:
        d = tempfile.mkdtemp(prefix='pcacc_writer_')
        try:
            t0 = time.perf_counter()
            with Acc.async_writer(n_threads=n_thr, compresslevel=lvl, max_pending=2 * n_thr) as wr:
                for k, b in enumerate(bevs):
                    wr.submit(b, f'bev_{k:04d}.pkl', d)
            dt = time.perf_counter() - t0
            res['levels'][str(lvl)] = {'bevs_per_s': len(bevs) / dt, 'raw_MB_per_s': raw * len(bevs) / dt / 1e6,
                                       'file_bytes_per_bev': wr.bytes_written / max(wr.n_written, 1)}
        finally:
            shutil.rmtree(d, ignore_errors=True)
    return res


def input_side_extra(torch, DeviceCloud, pk):
    """the dataloader's multi-camera projection and box -> point
    assignment for one nuScenes sample (10 sweeps x 34,688 points, 6 cameras, 64 boxes),
    device-resident, next to the same functions of the CPU port on one core."""
    c = synth.input_side_inputs(n=34688 * 10, n_boxes=64)
    cloud = DeviceCloud(1024, 4)
    pts = torch.from_numpy(c['pc']).cuda()
    pts32 = torch.from_numpy(c['pc_f32']).cuda()

    def timed(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    ms_p = timed(lambda: cloud.project_cameras(pts, c['glob_from_ego'], c['cams']))
    ms_b = timed(lambda: cloud.assign_boxes(pts32, c['boxes'], c['sizes'], c['tolerance']))
    n = int(pts.shape[0])
    from oracle import oracle as orc          # CPU port, timed as the baseline of this row only
    t0 = time.perf_counter()
    orc.project_to_cameras(c['pc'], c['glob_from_ego'], c['cams'])
    t1 = time.perf_counter()
    orc.assign_boxes(c['pc_f32'], c['boxes'], c['sizes'], c['tolerance'])
    t2 = time.perf_counter()
    cloud.close()
    return {'points': n, 'cameras': len(c['cams']), 'boxes': len(c['boxes']),
            'project_cameras_us': ms_p * 1e3, 'project_points_per_s': n / (ms_p * 1e-3),
            'project_alg_GBps': 48.0 * n / (ms_p * 1e-3) / 1e9,     # 24 B in, 16 B uv + 8 B index out
            'assign_boxes_us': ms_b * 1e3, 'assign_points_per_s': n / (ms_b * 1e-3),
            'assign_alg_GBps': 16.0 * n / (ms_b * 1e-3) / 1e9,      # 12 B in (f32 xyz), 4 B out
            'cpu_port_project_points_per_s': n / (t1 - t0), 'cpu_port_assign_points_per_s': n / (t2 - t1),
            'note': 'includes the per-call host work (numpy inverses, parameter block upload)'}


# ---------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on the host cores
# ---------------------------------------------------------------------------
_CPU_ACC = None
WINDOWS = ('present', 'future', 'full')
WORKLOAD = ('nuscenes-shaped oracle-pose scenes (BASELINE configs[1]) batched as configs[3]: 40 sweeps x '
            '34688 pts, CAM_FRONT 1600x900, 32 BEVs/scene (8 present idx x 4 aug), 256x256, '
            'present/future/full')


def bev_planes(bev):
    """(3 windows, 7 planes, P, P) float16 of one BEV dict, in libpcacc's plane order."""
    return np.stack([np.concatenate([bev[f'road_{w}'][None], bev[f'intensity_{w}'][None], bev[f'rgb_{w}'],
                                     bev[f'dynamic_{w}'][None], bev[f'elevation_{w}'][None]]) for w in WINDOWS])


def _cpu_bev(job):
    p, aug = job
    return bev_planes(_CPU_ACC.generate_bev(p, **aug))


def draw_jobs(rng, present_idxs):
    jobs = []
    for p in present_idxs:
        for _ in range(BEVS_PER_PRESENT):
            rot = 2 * np.pi * rng.random_sample()
            r, a = 5.0 * rng.random_sample(), 2 * np.pi * rng.random_sample()
            z = 1 + min(max(rng.normal(0, 0.1), -0.1), 0.1)
            jobs.append((p, dict(rot_ang=rot, trans_dx=r * np.cos(a), trans_dy=r * np.sin(a),
                                 zoom_scalar=z, do_warping=True)))
    return jobs


def cpu_scene(scene, n_sweeps, jobs, workers=1):
    """One bounded sample of the workload on the CPU port (oracle/oracle.py): returns
    (points integrated, BEV planes, seconds, the oracle accumulator).  integrate() is sequential
    by nature; the BEVs of the scene are independent and are spread over `workers` forked
    processes — the reference's own parallelism (Pool(bev_num), kitti360_sem_pc_accum.py:236-241).
    jobs: list of (present_idx, augmentation dict)."""
    global _CPU_ACC
    from oracle import oracle as orc
    bp = bev_setup()
    gp = dict(sem_idxs=synth.SEM_IDXS, view_size=bp['view_size'], pixel_size=P,
              int_scaler=bp['int_scaler'], int_sep_scaler=bp['int_sep_scaler'],
              int_mid_threshold=bp['int_mid_threshold'], height_filter=bp['height_filter'],
              rgb_fill=0)
    t0 = time.perf_counter()
    acc = orc.NuscOracle(synth.NUSC_FILTERS, gp)
    pts = 0
    for o in scene[:n_sweeps]:
        acc.integrate(o, o['_semseg'])
        pts += o['pc'].shape[0]
    _CPU_ACC = acc
    if workers > 1:
        import multiprocessing as mp
        with mp.get_context('fork').Pool(workers) as pool:
            bevs = pool.map(_cpu_bev, jobs)
    else:
        bevs = [_cpu_bev(j) for j in jobs]
    _CPU_ACC = None
    return pts, bevs, time.perf_counter() - t0, acc


def cpu_workers():
    return max(1, min(os.cpu_count() or 1, len(PRESENT_IDXS) * BEVS_PER_PRESENT))


def cpu_baseline_sample(scene, jobs):
    w = cpu_workers()
    jobs = [(p, dict(a, do_warping=True)) for p, a in jobs]
    pts, bevs, dt, acc = cpu_scene(scene, N_SWEEPS, jobs, w)
    n_b = len(bevs)
    lit = literal_reference_record()
    cpu = {'value': pts / dt, 'unit': 'points/s', 'cores': w, 'kind': 'port',
           'bevs_per_s': n_b / dt, 'seconds': dt,
           'sample': f'1 scene: {N_SWEEPS} sweeps ({pts} pts) + {n_b} BEVs through oracle/oracle.py '
                     '(vectorised numpy + C FMA-chain restatement of the reference; integrate serial, BEVs '
                     f'over {w} forked workers)',
           'host_cpus': os.cpu_count(), 'literal_reference': lit}
    return cpu, bevs, acc


def literal_reference_record():
    """The UNMODIFIED reference timed on this workload in the build container (it is pure Python
    and /root/reference does not travel to the GPU box): tools/time_literal_reference.py wrote
    profiles/literal_reference_cpu.json.  Reported beside the port, never as this run's number."""
    try:
        return json.load(open(os.path.join(ROOT, 'profiles', 'literal_reference_cpu.json')))
    except Exception:
        return None


def parity_gate(torch, cloud, scene_pass, out_planes, par0, cpu_bevs, cpu_acc, jobs):
    """In-run parity gate: bench scene 0 through the device path of the timed
    region (scene_pass) against the CPU port's output for the same 32 (present_idx,
    augmentation) pairs — kept-point counts `==`, exported records of three sweeps `==` (with
    the tracker's retro-active dynamic flags), cell indices of variant 0 `==` in the
    reference's order, 21 planes x 32 BEVs within 1 float16 ulp."""
    scene_pass(0, cloud, out_planes)
    flags = cloud.sync() & ~2                      # bit 1 is informational (float32 intensity)
    got = out_planes.cpu().numpy()
    first, n_live = cloud.live_frames()
    counts_ok = n_live == len(cpu_acc.sem_pcs) and all(
        cloud.frame_count(first + k) == cpu_acc.sem_pcs[k].shape[0] for k in range(n_live))
    rec_ok = all(np.array_equal(cloud.export_frame(first + k), cpu_acc.sem_pcs[k]) for k in (0, 17, 39))
    max_ulp, n_diff, n_vals = 0, 0, 0
    for v, ref in enumerate(cpu_bevs):
        d = np.abs(got[v].view(np.int16).astype(np.int32) - ref.view(np.int16).astype(np.int32))
        max_ulp = max(max_ulp, int(d.max()))
        n_diff += int((d != 0).sum())
        n_vals += d.size
    # cell indices of variant 0, in the reference's (frame, point) order
    _, _, cells = cloud.rasterise([par0[0]], P, want_cells=True)
    cloud.sync()
    p0, aug0 = jobs[0]
    dbg = cpu_acc.generate_bev(p0, return_f64=True, **dict(aug0, do_warping=True))['_debug']

    def window_cells(fids):
        out = []
        for f in fids:
            off, n = cloud.frame_offset(f), cloud.frame_count(f)
            c = cells[off:off + n].cpu().numpy()
            c = c[c >= 0]
            out.append(np.stack([c // P, c % P], axis=1))
        return np.concatenate(out)

    cells_ok = (np.array_equal(window_cells(range(first, first + p0)), dbg['cells_present'])
                and np.array_equal(window_cells(range(first + p0, first + n_live)), dbg['cells_future']))
    ok = bool(flags == 0 and counts_ok and rec_ok and cells_ok and max_ulp <= 1)
    return {'ok': ok, 'checked': f'bench scene 0: {n_live} sweeps, {len(cpu_bevs)} BEVs x 21 planes vs oracle/oracle.py',
            'kept_counts_equal': bool(counts_ok), 'records_equal_sweeps_0_17_39': bool(rec_ok),
            'cell_indices_equal_variant0': bool(cells_ok), 'planes_max_fp16_ulp': max_ulp,
            'planes_values_differing': n_diff, 'planes_values': n_vals, 'device_error_flags': int(flags)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    scene = make_scenes(0, 1)[0]
    rng = np.random.RandomState(5)
    w = cpu_workers()
    # the reference arm's step is ONE scene (a bounded sample of the S-scene step of the GPU arm:
    # a full 16-scene step would take minutes per step on the CPU)
    for _ in range(max(args.warmup, 0)):
        cpu_scene(scene, 8, draw_jobs(rng, PRESENT_IDXS[:1]), w)
    tot_pts, tot_b, tot_t = 0, 0, 0.0
    for _ in range(args.steps):
        pts, bevs, dt, _ = cpu_scene(scene, N_SWEEPS, draw_jobs(rng, PRESENT_IDXS), w)
        tot_pts, tot_b, tot_t = tot_pts + pts, tot_b + len(bevs), tot_t + dt
    v = tot_pts / tot_t
    bevs_per_scene = len(PRESENT_IDXS) * BEVS_PER_PRESENT
    sample = (f'each step = 1 scene of the workload: {N_SWEEPS} sweeps + {bevs_per_scene} BEVs '
              f'on the CPU port of the reference algorithm (oracle/oracle.py), integrate serial, BEVs over {w} '
              'forked workers; warm-up steps are 8-sweep / 4-BEV passes; the reference itself is pure Python '
              'and /root/reference is not present on the GPU box (its measured rate: cpu_baseline.literal_reference)')
    OUT.write(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'points/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 0), 'ms_per_step': tot_t / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'bevs_per_s': tot_b / tot_t,
        'config': bench_config(args.scenes_per_step, args.streams),
        'cpu_baseline': {'value': v, 'unit': 'points/s', 'cores': w, 'kind': 'port', 'sample': sample,
                         'host_cpus': os.cpu_count(), 'literal_reference': literal_reference_record()},
        'e2e': {'value': v, 'unit': 'points/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }) + '\n')
    OUT.flush()


def bench_config(S, n_str):
    """The `config` object of both arms (same keys, same values)."""
    n_in_scene = N_SWEEPS * 34688
    bevs_per_scene = len(PRESENT_IDXS) * BEVS_PER_PRESENT
    return {'workload': WORKLOAD, 'scenes_per_step_per_gpu': S, 'points_per_step_per_gpu': S * n_in_scene,
            'bevs_per_step_per_gpu': S * bevs_per_scene, 'sharding': 'scenes by rank',
            'streams_per_gpu': n_str,
            'l2': 'step inputs (%.1f GB resident copies) exceed the 126 MB L2' % (S * 0.32)}


def _quiet_stdout():
    """Libraries (NCCL's version banner, torchrun notices) write to fd 1; the contract is ONE
    JSON line on stdout.  Everything else is routed to stderr; the JSON goes to the saved fd."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, 'w')


def main():
    global OUT
    OUT = _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--scenes-per-step', type=int, default=16)
    ap.add_argument('--streams', type=int, default=4)
    ap.add_argument('--e2e-scenes', type=int, default=16)
    ap.add_argument('--e2e-steps', type=int, default=2)
    ap.add_argument('--e2e-threads', type=int, default=2)
    ap.add_argument('--switch-interval', type=float, default=1e-4)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-c3', action='store_true')
    ap.add_argument('--ncu-step', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    import torch
    # torchrun pins OMP_NUM_THREADS=1; the host-side staging copies of the e2e path are
    # parallel memcpys, so give every rank its share of the cores back
    torch.set_num_threads(max(1, min(16, (os.cpu_count() or 1) // max(world, 1))))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the product has no CPU path '
                         '(use --impl reference for the host baseline)')
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    try:
        run_ours(args, rank, world, local_rank, dist)
    finally:
        if dist is not None:
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
