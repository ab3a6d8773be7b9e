"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  It is
used by `tests/golden/make_golden.py` to generate the committed golden vectors
and by `oracle/validate_against_reference.py` to pin the numpy/C restatement in
`oracle/oracle.py` against the literal reference.  Nothing on the product path,
in `-m gpu` tests, `smoke()` or `bench.py` imports this module.

Recipe: third-party modules that are absent here
(open3d, onnxruntime, matplotlib, nuscenes-devkit, pyquaternion) are stubbed in
`sys.modules`; Open3D's ICP is replaced by a scripted list of 4x4 transforms
(ICP is out of scope: the pose is an input); the ONNX model by an object with
the same `.pred()` contract.
"""
from __future__ import annotations

import importlib.machinery as _im
import os
import sys
import types

REF_ROOT = '/root/reference'
ICP_QUEUE: list = []          # push one T_new_prev per KITTI integrate() call
_loaded = {}


def available() -> bool:
    return os.path.isdir(REF_ROOT)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = _im.ModuleSpec(name, None)
    m.__path__ = []
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load():
    """Returns a namespace with the reference's classes."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError('/root/reference is not present on this machine')

    o3d = _stub('open3d')
    _stub('onnxruntime')
    mpl = _stub('matplotlib')
    plt = _stub('matplotlib.pyplot')
    mpl.pyplot = plt
    plt.style = types.SimpleNamespace(use=lambda *a, **k: None)
    _stub('nuscenes')
    _stub('nuscenes.nuscenes', NuScenes=object)
    _stub('nuscenes.utils')
    _stub('nuscenes.utils.data_classes', LidarPointCloud=object)
    _stub('nuscenes.utils.geometry_utils', transform_matrix=None,
          view_points=None)
    _stub('nuscenes.map_expansion')
    _stub('nuscenes.map_expansion.map_api', NuScenesMap=object)
    _stub('pyquaternion', Quaternion=object)

    def _icp(target, source, thr, init, est):
        return types.SimpleNamespace(transformation=ICP_QUEUE.pop(0))

    o3d.pipelines = types.SimpleNamespace(registration=types.SimpleNamespace(
        registration_icp=_icp,
        TransformationEstimationPointToPlane=lambda: None))

    saved = {k: sys.modules.get(k) for k in ('datasets', 'utils')}
    sys.path.insert(0, REF_ROOT)
    # the reference's datasets/ and utils/ are namespace packages which an
    # installed regular package of the same name would shadow
    _stub('datasets').__path__ = [os.path.join(REF_ROOT, 'datasets')]
    _stub('utils').__path__ = [os.path.join(REF_ROOT, 'utils')]
    try:
        import sem_pc_accum as ref_sem_pc_accum
        import kitti360_sem_pc_accum as ref_kitti
        import nuscenes_oracle_sem_pc_accum as ref_nusc
        from bev_generator import bev_generator as ref_bevgen
        from bev_generator import sem_bev as ref_sem_bev
        from datasets import nuscenes_utils as ref_nusc_utils
    finally:
        sys.path.remove(REF_ROOT)

    ref_kitti.Kitti360SemanticPointCloudAccumulator.pc2pcd = staticmethod(
        lambda pc: None)

    _loaded.update(
        sem_pc_accum=ref_sem_pc_accum,
        KittiAccum=ref_kitti.Kitti360SemanticPointCloudAccumulator,
        NuscOracleAccum=ref_nusc.NuScenesOracleSemanticPointCloudAccumulator,
        SemBEVGenerator=ref_sem_bev.SemBEVGenerator,
        BEVGenerator=ref_bevgen.BEVGenerator,
        homo_transform=ref_nusc_utils.homo_transform,
        pts_feat_from_img=ref_nusc_utils.pts_feat_from_img,
        nusc_utils=ref_nusc_utils,
        base_module=ref_sem_pc_accum,
    )
    # keep the stubs for `datasets`/`utils` registered: the reference modules
    # hold references to them; restore anything we displaced that the caller
    # may need afterwards
    for k, v in saved.items():
        if v is not None:
            _loaded['_displaced_' + k] = v
    return types.SimpleNamespace(**_loaded)


def make_kitti_accum(ref, horizon_dist, calib, filters, sem_idxs, bev_params,
                     semseg_model=None, use_gt_sem=False):
    """Constructs the reference KITTI-360 accumulator without an ONNX session
    (sem_pc_accum.py:79-81 only builds one when use_gt_sem is False)."""
    acc = ref.KittiAccum(horizon_dist, calib, 1.0, None, filters, sem_idxs,
                         True, bev_params)
    acc.use_gt_sem = use_gt_sem
    acc.semseg_model = semseg_model
    return acc


def make_nusc_accum(ref, filters, sem_idxs, bev_params, semseg_model):
    """The oracle accumulator forbids use_gt_sem
    (nuscenes_oracle_sem_pc_accum.py:59-60), so the SemSegONNX name is swapped
    for a factory returning the stand-in before construction."""
    mod = ref.base_module
    orig = mod.SemSegONNX
    mod.SemSegONNX = lambda path: semseg_model
    try:
        acc = ref.NuscOracleAccum(None, filters, sem_idxs, False, bev_params)
    finally:
        mod.SemSegONNX = orig
    return acc
