import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from pc_accumulation_lib_b200 import device, synth, _lib
scene = bench.make_scenes(0, 1)[0]
n_in = sum(o['pc'].shape[0] for o in scene)
cloud = device.DeviceCloud(n_in + 4096, 64)
T = np.eye(4)
def run():
    cloud.reset()
    t0 = time.perf_counter()
    for o in scene:
        cloud.integrate_records_host(o['pc'], o['pc_cam_idx'], o['images'], o['_semseg'], T, synth.NUSC_FILTERS, 255., _lib.STAGE_SPARSE)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / len(scene) * 1e6
run(); run()
print(os.environ.get('TAG'), 'sparse integrate host us/sweep:', [round(run(), 1) for _ in range(5)])
