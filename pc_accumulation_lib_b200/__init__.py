"""pc_accumulation_lib_b200 — B200-native drop-in for the per-frame semantic
fusion and BEV rasterisation path of robin-karlsson0/pc-accumulation-lib.

Importing the package does not touch CUDA; constructing an accumulator or a
DeviceCloud does, and fails loudly (ImportError / PcaccError) when libpcacc.so
is not built or no CUDA device is present — there is no CPU path.
"""
__all__ = ['Kitti360SemanticPointCloudAccumulator', 'NuScenesOracleSemanticPointCloudAccumulator',
           'SemanticPointCloudAccumulator', 'SemBEVGenerator', 'RGBBEVGenerator', 'BEVGenerator',
           'DeviceCloud', 'AsyncBevWriter', 'pinned_empty', 'pinned_like', 'pin_observation']


def __getattr__(name):
    if name in ('Kitti360SemanticPointCloudAccumulator',):
        from .kitti360_sem_pc_accum import Kitti360SemanticPointCloudAccumulator as c
        return c
    if name in ('NuScenesOracleSemanticPointCloudAccumulator',):
        from .nuscenes_oracle_sem_pc_accum import NuScenesOracleSemanticPointCloudAccumulator as c
        return c
    if name == 'SemanticPointCloudAccumulator':
        from .sem_pc_accum import SemanticPointCloudAccumulator as c
        return c
    if name in ('SemBEVGenerator', 'RGBBEVGenerator', 'BEVGenerator'):
        from . import bev_generator as m
        return getattr(m, name)
    if name == 'AsyncBevWriter':
        from .sem_pc_accum import AsyncBevWriter as c
        return c
    if name in ('DeviceCloud', 'pinned_empty', 'pinned_like', 'pin_observation'):
        from . import device as m
        return getattr(m, name)
    raise AttributeError(name)
