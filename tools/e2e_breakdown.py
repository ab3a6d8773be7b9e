"""Wall-clock breakdown of generate_bev / integrate of the reference-facing API (no profiler:
perf_counter around the named functions).  Scratch tool."""
import sys, os, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pc_accumulation_lib_b200 import NuScenesOracleSemanticPointCloudAccumulator as Acc, pin_observation, synth
from pc_accumulation_lib_b200 import device, sem_pc_accum
from pc_accumulation_lib_b200.bev_generator import bev_generator as bg

T = collections.defaultdict(float)
N = collections.Counter()


def wrap(obj, name, label=None):
    f = getattr(obj, name)
    label = label or name

    def g(*a, **k):
        t0 = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            T[label] += time.perf_counter() - t0
            N[label] += 1
    setattr(obj, name, g)


for name in ('integrate', 'obs2sem_vec_space', 'get_split_dyn_obj_trajs', 'generate_bev', '_boxes_to_world', 'get_dyn_obj_trajs'):
    wrap(Acc, name)
for name in ('_window_inputs', '_generate', '_sync'):
    wrap(sem_pc_accum.SemanticPointCloudAccumulator, name)
for name in ('generate_batch', '_rasterise_windows_begin', '_rasterise_device', 'preprocess_trajs_batch', 'rand_aug_params',
             'draw_warp', 'warp_trajs', '_assemble', 'get_random_warp_params'):
    wrap(bg.BEVGenerator, name)
for name in ('rasterise', 'planes_to_host_begin', 'planes_to_host_finish', 'integrate_records_host', 'mark_dynamic', 'flush_marks', 'warp_planes'):
    wrap(device.DeviceCloud, name)
wrap(device, 'make_bev_params_batch')
from pc_accumulation_lib_b200.bev_generator import sem_bev
wrap(sem_bev.SemBEVGenerator, '_assemble')
wrap(sem_pc_accum._LazyTrajs, '_materialise')
wrap(device.DeviceCloud, 'sync')
wrap(device.DeviceCloud, 'refresh')
bg.make_bev_params_batch = device.make_bev_params_batch

scene = [pin_observation(o) for o in bench.make_scenes(0, 1)[0]]
semseg = bench.HostSemseg()
for o in scene:
    for img, cls in zip(o['images'], o['_semseg']):
        semseg.by_id[id(img)] = cls
n_in = sum(o['pc'].shape[0] for o in scene)
acc = Acc(semseg, synth.NUSC_FILTERS, synth.SEM_IDXS, None, bench.bev_setup(),
          ring_capacity_pts=n_in + 4096, ring_max_frames=bench.N_SWEEPS + 8)
acc.sync_each_integrate = False


def run():
    acc.reset()
    for o in scene:
        acc.integrate([o])
    acc.sem_bev_generator.rng = np.random.RandomState(1)
    for p in bench.PRESENT_IDXS:
        acc.generate_bev(p, bench.BEVS_PER_PRESENT, True)


run(); run()
T.clear(); N.clear()
R = 20
t0 = time.perf_counter()
for _ in range(R):
    run()
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print(f'{tot / R * 1e3:.2f} ms per scene ({n_in * R / tot / 1e6:.1f} M points/s incl. instrumentation)')
for k, v in sorted(T.items(), key=lambda kv: -kv[1]):
    print(f'{k:28s} {v / R * 1e3:7.3f} ms/scene  {v / N[k] * 1e6:8.1f} us/call  x{N[k] // R}')
