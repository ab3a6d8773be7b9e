"""Quick device-resident timing of the long-horizon window (BASELINE config 3):
F all-points frames integrated with the use_gt_sem path, one BEV at F/2.
Prints per-stage CUDA-event times. Scratch tool for kernel work, not the
contract bench (bench.py)."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pc_accumulation_lib_b200 import device as dev, synth

ap = argparse.ArgumentParser()
ap.add_argument('--frames', type=int, default=200)
ap.add_argument('--P', type=int, default=256)
ap.add_argument('--iters', type=int, default=5)
ap.add_argument('--eager', type=int, default=0)
ap.add_argument('--variants', type=int, default=1)
a = ap.parse_args()

F = a.frames
n_distinct = 8
pcs = [torch.from_numpy(synth.kitti_lidar(synth.seed_for(3, f))).cuda() for f in range(n_distinct)]
sgs = [torch.from_numpy(synth.kitti_sem_gt(synth.seed_for(3, f), 120000)[:, 0].copy()).cuda()
       for f in range(n_distinct)]
Ts = [synth.kitti_step_transform(synth.seed_for(3, f)) for f in range(F)]
N = pcs[0].shape[0]
cloud = dev.DeviceCloud(capacity_pts=F * N + 1024, max_frames=F + 8)

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

def integrate_all():
    cloud.reset()
    poses = []
    for f in range(F):
        if f > 0:
            cloud.rebase(Ts[f], eager=bool(a.eager))
            poses = [list((Ts[f] @ np.array(p + [1.]))[:3]) for p in poses]
        cloud.integrate_gt(pcs[f % n_distinct], sgs[f % n_distinct], synth.KITTI_FILTERS)
        poses.append([0., 0., 0.])
    return poses

for it in range(a.iters):
    e0 = ev()
    poses = integrate_all()
    e1 = ev()
    torch.cuda.synchronize()
    print(f'integrate {F} frames: {e0.elapsed_time(e1):.3f} ms  ({F*N/e0.elapsed_time(e1)/1e6:.2f} Gpts/s)')
cloud.sync()
first, n_live = cloud.live_frames()
p = F // 2
origin = np.array(poses[p])
d = np.array(poses[p - 1]) - np.array(poses[p - 2])
rot = np.pi - (0.5 * np.pi + np.arctan2(d[1], d[0]))
R = np.array([[np.cos(rot), -np.sin(rot), 0], [np.sin(rot), np.cos(rot), 0], [0, 0, 1]])
bps = [dev.make_bev_params(first, first + p, first + n_live, origin, R, 0.1 * v, 0., 80., None, 20., 20., .5, 0,
                           synth.SEM_IDXS) for v in range(a.variants)]
out = torch.empty((a.variants, 3, 7, a.P, a.P), dtype=torch.float16, device='cuda')
n_res = cloud.resident_points()
cloud.profile(True)
cloud.profile_read()
for it in range(a.iters):
    e0 = ev()
    cloud.rasterise(bps, a.P, out=out)
    e1 = ev()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    alg = (33 * n_res + 42 * a.P * a.P) * a.variants
    print(f'rasterise {n_res} pts x{a.variants}: {ms:.3f} ms  alg {alg/ms/1e6:.1f} GB/s', cloud.raster_stats())

prof = cloud.profile_read()
print({k: round(v[0] / a.iters * 1e3, 1) for k, v in prof.items() if v[1]})
# for `ncu --profile-from-start off`: one more rasterise between cudaProfilerStart/Stop
cloud.profile(False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
cloud.rasterise(bps, a.P, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
if os.environ.get('C3_CELL_STATS'):
    _, _, cells = cloud.rasterise(bps[:1], a.P, want_cells=True)
    c = cells[cells >= 0].long()
    cnt = torch.bincount(c, minlength=a.P * a.P)
    s = torch.sort(cnt, descending=True).values
    print('cells non-empty', int((cnt > 0).sum()), 'points', int(cnt.sum()), 'largest', s[:12].tolist(),
          'quantiles 50/90/99/99.9 %', [int(torch.quantile(cnt[cnt > 0].float(), q)) for q in (0.5, 0.9, 0.99, 0.999)],
          '> 15:', int((cnt > 15).sum()), '> 32:', int((cnt > 32).sum()), '> 1024:', int((cnt > 1024).sum()))
