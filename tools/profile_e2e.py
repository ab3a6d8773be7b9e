"""cProfile of the reference-facing API path (host buffers in, numpy out) for one
nuScenes-shaped scene: where does the end-to-end time go?"""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pc_accumulation_lib_b200 import NuScenesOracleSemanticPointCloudAccumulator, synth

scene = bench.make_scenes(0, 1)[0]
semseg = bench.HostSemseg()
for o in scene:
    for img, cls in zip(o['images'], o['_semseg']):
        semseg.by_id[id(img)] = cls
n_in = sum(o['pc'].shape[0] for o in scene)

def run():
    acc = NuScenesOracleSemanticPointCloudAccumulator(
        semseg, synth.NUSC_FILTERS, synth.SEM_IDXS, None, bench.bev_setup(),
        ring_capacity_pts=n_in + 4096, ring_max_frames=bench.N_SWEEPS + 8)
    acc.sync_each_integrate = SYNC
    t0 = time.perf_counter()
    for o in scene:
        acc.integrate([o])
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    acc.sem_bev_generator.rng = np.random.RandomState(1)
    n = 0
    for p in bench.PRESENT_IDXS:
        n += len(acc.generate_bev(p, bench.BEVS_PER_PRESENT, True))
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, n

import os as _os
SYNC = bool(int(_os.environ.get('E2E_SYNC', '0')))
run()
ti, tb, n = run()
print(f'integrate 40 sweeps: {ti*1e3:.1f} ms ({ti/40*1e3:.2f} ms/sweep); {n} BEVs: {tb*1e3:.1f} ms ({tb/n*1e3:.2f} ms/BEV)')
pr = cProfile.Profile()
pr.enable()
run()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
pstats.Stats(pr).sort_stats('tottime').print_stats(22)
