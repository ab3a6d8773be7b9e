"""cProfile of the KITTI-360-shaped sequence (BASELINE configs[0]) through the reference-facing API:
20 x 120 000-point frames through integrate(), one generate_bev per present index.  Scratch tool."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pc_accumulation_lib_b200 import Kitti360SemanticPointCloudAccumulator, pinned_like, synth

F, P = 20, 256
frames = []
for f in range(F):
    seed = synth.seed_for(1, f)
    frames.append(dict(pc=pinned_like(synth.kitti_lidar(seed)), T=synth.kitti_step_transform(seed),
                       rgb=pinned_like(synth.kitti_rgb(seed)), cls=pinned_like(synth.kitti_class_map_fast(seed))))
n_pts = sum(fr['pc'].shape[0] for fr in frames)
pr = cProfile.Profile()


def run(profile=False):
    acc = Kitti360SemanticPointCloudAccumulator(
        1e9, synth.kitti_calib(), 1.0, synth.FakeSemseg([fr['cls'] for fr in frames]), synth.KITTI_FILTERS,
        synth.SEM_IDXS, False, synth.kitti_bev_params(pixel_size=P), ring_capacity_pts=n_pts // 4 + 4096,
        ring_max_frames=F + 8)
    acc.sync_each_integrate = False
    acc.generate_bev  # noqa
    if profile:
        pr.enable()
    t0 = time.perf_counter()
    for fr in frames:
        acc.integrate([(fr['rgb'], fr['pc'], None, fr['T'])])
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for p in (5, 10, 15):
        acc.generate_bev(p, 1, True)
    t2 = time.perf_counter()
    if profile:
        pr.disable()
    acc.cloud.close()
    return (t1 - t0) / F, (t2 - t1) / 3


for _ in range(4):
    ti, tb = run()
    print(f'integrate {ti * 1e6:.1f} us/frame, generate_bev {tb * 1e6:.1f} us/call')
for _ in range(10):
    run(True)
pstats.Stats(pr).sort_stats('tottime').print_stats(22)
