"""cProfile of the reference-facing API path (host buffers in, numpy out) for one
nuScenes-shaped scene: where does the end-to-end time go?
E2E_PINNED=1 (default): observations in page-locked memory; 0: pageable (sparse staging)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pc_accumulation_lib_b200 import NuScenesOracleSemanticPointCloudAccumulator, pin_observation, synth

PINNED = bool(int(os.environ.get('E2E_PINNED', '1')))
SYNC = bool(int(os.environ.get('E2E_SYNC', '0')))
scene = bench.make_scenes(0, 1)[0]
if PINNED:
    scene = [pin_observation(o) for o in scene]
semseg = bench.HostSemseg()
for o in scene:
    for img, cls in zip(o['images'], o['_semseg']):
        semseg.by_id[id(img)] = cls
n_in = sum(o['pc'].shape[0] for o in scene)
acc = NuScenesOracleSemanticPointCloudAccumulator(
    semseg, synth.NUSC_FILTERS, synth.SEM_IDXS, None, bench.bev_setup(),
    ring_capacity_pts=n_in + 4096, ring_max_frames=bench.N_SWEEPS + 8)
acc.sync_each_integrate = SYNC


def run():
    acc.reset()
    t0 = time.perf_counter()
    for o in scene:
        acc.integrate([o])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t1s = time.perf_counter()
    acc.sem_bev_generator.rng = np.random.RandomState(1)
    n = 0
    for p in bench.PRESENT_IDXS:
        n += len(acc.generate_bev(p, bench.BEVS_PER_PRESENT, True))
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return t1 - t0, t1s - t1, t2 - t1s, n


run()
for _ in range(3):
    ti, tw, tb, n = run()
    print(f'pinned={PINNED} staging={acc.cloud.last_staging}: integrate 40 sweeps host {ti*1e3:.2f} ms '
          f'({ti/40*1e3:.3f} ms/sweep) + drain {tw*1e3:.2f} ms; {n} BEVs: {tb*1e3:.2f} ms '
          f'({tb/8*1e3:.3f} ms/call) -> {n_in/(ti+tw+tb)/1e6:.1f} M points/s')
pr = cProfile.Profile()
pr.enable()
for _ in range(20):          # 20 scenes: pstats prints milliseconds, per-scene = value / 20
    run()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
pstats.Stats(pr).sort_stats('tottime').print_stats(25)
