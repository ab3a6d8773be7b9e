"""Times the UNMODIFIED reference (imported from /root/reference through oracle/ref_loader.py,
third-party modules stubbed) on one scene of bench.py's workload, next to the CPU port
(oracle/oracle.py) on the same machine, and writes profiles/literal_reference_cpu.json.

The reference is pure Python and /root/reference does not exist on the GPU box, so this runs in
the BUILD container only; bench.py reports the recorded figures beside the port it times live
(`cpu_baseline.literal_reference`).  Usage: python tools/time_literal_reference.py [n_bevs]
"""
import contextlib
import io
import json
import os
import platform
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                                               # noqa: E402

import bench                                                     # noqa: E402
from oracle import ref_loader                                    # noqa: E402
from pc_accumulation_lib_b200 import synth                       # noqa: E402


def main():
    n_bevs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    scene = bench.make_scenes(0, 1)[0]
    n_pts = sum(o['pc'].shape[0] for o in scene)
    ref = ref_loader.load()
    semseg = synth.SceneSemseg()
    for o in scene:
        semseg.register(o)
    acc = ref_loader.make_nusc_accum(ref, synth.NUSC_FILTERS, synth.SEM_IDXS, bench.bev_setup(), semseg)
    sink = io.StringIO()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(sink):
        for o in scene:
            acc.integrate([o])
    t_int = time.perf_counter() - t0
    present = bench.PRESENT_IDXS[:n_bevs]
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(sink):
        for p in present:
            acc.generate_bev(p, 1, True)          # one random-augmentation BEV, as run_nuscenes_bev_gen.py
    t_bev = (time.perf_counter() - t0) / len(present)
    bevs_per_scene = len(bench.PRESENT_IDXS) * bench.BEVS_PER_PRESENT
    t_scene = t_int + t_bev * bevs_per_scene

    # the port on the same machine, one core, same scene
    rng = np.random.RandomState(5)
    jobs = bench.draw_jobs(rng, present)[::bench.BEVS_PER_PRESENT]
    pts, bevs, dt, _ = bench.cpu_scene(scene, bench.N_SWEEPS, [], 1)
    t_int_port = dt
    t0 = time.perf_counter()
    bench.cpu_scene(scene, bench.N_SWEEPS, jobs, 1)
    t_bev_port = (time.perf_counter() - t0 - t_int_port) / len(jobs)
    t_scene_port = t_int_port + t_bev_port * bevs_per_scene
    out = {
        'what': 'UNMODIFIED reference (robin-karlsson0/pc-accumulation-lib, imported with third-party stubs) '
                'on bench scene 0: 40 sweeps x 34688 pts through NuScenesOracleSemanticPointCloudAccumulator.integrate, '
                f'{len(present)} BEVs through generate_bev(p, 1, True), scaled to the scene\'s {bevs_per_scene} BEVs',
        'where': 'build container (no GPU), 1 core: the reference is single-threaded Python',
        'cpu': platform.processor() or platform.machine(), 'host_cpus': os.cpu_count(),
        'numpy': np.__version__,
        'integrate_s_per_scene': t_int, 'integrate_points_per_s': n_pts / t_int,
        'bev_s': t_bev, 'bevs_timed': len(present), 'scene_s_extrapolated': t_scene,
        'points_per_s': n_pts / t_scene, 'bevs_per_s': bevs_per_scene / t_scene,
        'port_same_machine_1core': {'integrate_s_per_scene': t_int_port, 'bev_s': t_bev_port,
                                    'points_per_s': n_pts / t_scene_port,
                                    'bevs_per_s': bevs_per_scene / t_scene_port},
        'port_over_literal': t_scene / t_scene_port,
    }
    # BASELINE.json configs[0]: 20 KITTI-360-shaped frames (120,000 points, 1408x376) + one 256^2 BEV
    F, p_idx = 20, 10
    frames = []
    for f in range(F):
        seed = synth.seed_for(1, f)
        frames.append(dict(pc=synth.kitti_lidar(seed), T=synth.kitti_step_transform(seed),
                           rgb=synth.kitti_rgb(seed), cls=synth.kitti_class_map_fast(seed)))
    kacc = ref_loader.make_kitti_accum(ref, 1e9, synth.kitti_calib(), synth.KITTI_FILTERS, synth.SEM_IDXS,
                                       synth.kitti_bev_params(pixel_size=256),
                                       synth.FakeSemseg([fr['cls'] for fr in frames]), use_gt_sem=False)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(sink):
        for fr in frames:
            ref_loader.ICP_QUEUE.append(fr['T'])
            kacc.integrate([(fr['rgb'], fr['pc'], None)])
    t_kint = time.perf_counter() - t0
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(sink):
        kacc.generate_bev(p_idx, 1, True)
    t_kbev = time.perf_counter() - t0
    out['kitti360_sequence'] = {
        'what': 'configs[0]: 20 frames x 120000 pts through Kitti360SemanticPointCloudAccumulator.integrate '
                '(ICP scripted, semseg stand-in) + generate_bev(10, 1, True)',
        'integrate_ms_per_frame': t_kint / F * 1e3, 'integrate_points_per_s': F * 120000 / t_kint,
        'bev_s': t_kbev, 'resident_points': int(sum(s.shape[0] for s in kacc.sem_pcs))}
    path = os.path.join(ROOT, 'profiles', 'literal_reference_cpu.json')
    json.dump(out, open(path, 'w'), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
