"""ctypes binding of libpcacc.so (include/pcacc.h).

The shared library is the product: there is no Python or CPU fallback.  If it
has not been built (`python -c "import __graft_entry__ as g; g.build()"` or
`make -C pc_accumulation_lib_b200/csrc`) loading raises ImportError, and every
compute entry point raises PcaccError when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libpcacc.so')

OK = 0
ERR_ARG, ERR_CUDA, ERR_CAPACITY, ERR_NOMEM, ERR_STATE = -1, -2, -3, -4, -5
FLAG_UV_OUT_OF_IMAGE, FLAG_INTENSITY_F32, FLAG_ATTR_RANGE, FLAG_CELL_OVERFLOW = 1, 2, 4, 8
SEM_U8, SEM_I32, SEM_I64, SEM_F32_PROB, SEM_I16 = 0, 1, 2, 3, 4
IMG_F32, IMG_F64 = 6, 7
BEV_PLANES, BEV_WINDOWS = 7, 3
STAGE_AUTO, STAGE_DIRECT, STAGE_SPARSE = 0, 1, 2
OPT_REDUCE_STRIPS, OPT_CLASSIFY_SINGLE = 1, 2
ABI_VERSION = 1


class PcaccError(RuntimeError):
    def __init__(self, status, detail):
        super().__init__(f'libpcacc status {status}: {detail}')
        self.status = status


class BevParams(C.Structure):
    """Mirror of pcacc_bev_params (include/pcacc.h)."""
    _fields_ = [
        ('frame_begin', C.c_int64), ('frame_split', C.c_int64),
        ('frame_end', C.c_int64),
        ('origin', C.c_double * 3), ('R', C.c_double * 9),
        ('trans_dx', C.c_double), ('trans_dy', C.c_double),
        ('view', C.c_double), ('height_filter', C.c_double),
        ('int_scaler', C.c_double), ('int_sep_scaler', C.c_double),
        ('int_mid_threshold', C.c_double), ('rgb_fill', C.c_double),
        ('road_cls', C.c_int32), ('veh_cls', C.c_int32 * 4),
        ('elevation_max', C.c_int32),
    ]


_vp, _i64, _i32, _dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double

# name -> (restype, argtypes); every symbol include/pcacc.h declares
SIGNATURES = {
    'pcacc_strerror': (C.c_char_p, [_i32]),
    'pcacc_last_error': (C.c_char_p, [_vp]),
    'pcacc_abi_version': (_i32, []),
    'pcacc_create': (_i32, [_i32, _i64, _i32, C.POINTER(_vp)]),
    'pcacc_destroy': (_i32, [_vp]),
    'pcacc_reset': (_i32, [_vp, _vp]),
    'pcacc_project': (_i32, [_vp, _i64, _i32, _vp, _i32, _i32, _dbl, _vp, _vp,
                             _vp, _vp]),
    'pcacc_velo2img': (_i32, [_vp, _vp, _i64, _i32, _vp, _i32, _i32, _dbl, _vp, _vp, _vp]),
    'pcacc_gen_semantic_pc': (_i32, [_vp, _vp, _i64, _vp, _vp, _i32, _i32,
                                     _i32, _i32, _vp, _vp, _vp]),
    'pcacc_integrate_frustum': (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _i32,
                                       _i32, _i32, _i32, _dbl, _vp, _i32,
                                       C.POINTER(_i64), _vp]),
    'pcacc_integrate_gt': (_i32, [_vp, _vp, _i64, _vp, _vp, _i32,
                                  C.POINTER(_i64), _vp]),
    'pcacc_integrate_records': (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _i32,
                                       _i32, _i32, _i32, _vp, _dbl, _vp, _i32,
                                       C.POINTER(_i64), _vp]),
    'pcacc_integrate_records_batch': (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32,
                                             _i32, _vp, _dbl, _vp, _i32, C.POINTER(_i64), _vp]),
    'pcacc_integrate_records_host': (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _i32, _i32, _i32, _i32,
                                            _vp, _dbl, _vp, _i32, _i32, C.POINTER(_i32),
                                            C.POINTER(_i64), _vp]),
    'pcacc_host_is_pinned': (_i32, [_vp]),
    'pcacc_memcpy_d2h_async': (_i32, [_vp, _vp, C.c_size_t, _vp]),
    'pcacc_integrate_cloud': (_i32, [_vp, _vp, _i64, C.POINTER(_i64), _vp]),
    'pcacc_rebase': (_i32, [_vp, _vp, _i32, _vp]),
    'pcacc_evict': (_i32, [_vp, _i32]),
    'pcacc_set_option': (_i32, [_vp, _i32, _i32]),
    'pcacc_mark_dynamic': (_i32, [_vp, _vp, _vp, _i32, _vp]),
    'pcacc_sync': (_i32, [_vp, C.POINTER(C.c_uint32), _vp]),
    'pcacc_num_frames': (_i32, [_vp, C.POINTER(_i64), C.POINTER(_i32)]),
    'pcacc_frame_count': (_i32, [_vp, _i64, C.POINTER(_i64)]),
    'pcacc_resident_points': (_i64, [_vp]),
    'pcacc_export_frame': (_i32, [_vp, _i64, _vp, _vp]),
    'pcacc_rasterise': (_i32, [_vp, C.POINTER(BevParams), _i32, _i32, _vp, _vp,
                               _vp, _vp]),
    'pcacc_warp_planes': (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    'pcacc_frame_offset': (_i32, [_vp, _i64, C.POINTER(_i64)]),
    'pcacc_raster_stats': (_i32, [_vp, C.POINTER(_i64 * 3), _vp]),
    'pcacc_crop_trajectory': (_i32, [_vp, _i32, _dbl, _dbl, _vp, C.POINTER(_i32)]),
    'pcacc_preprocess_trajectories': (_i32, [_vp, _vp, _i32, _vp, _i32, _i32, _dbl, _vp, _vp]),
    'pcacc_event_create': (_i32, [C.POINTER(_vp)]),
    'pcacc_event_record': (_i32, [_vp, _vp]),
    'pcacc_event_sync': (_i32, [_vp]),
    'pcacc_event_destroy': (_i32, [_vp]),
    'pcacc_assign_boxes': (_i32, [_vp, _vp, _i32, _i64, _i64, _vp, _vp, _i32, _dbl, _vp, _vp, _vp]),
    'pcacc_project_cameras': (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i32, _dbl, _vp, _vp, _vp]),
    'pcacc_pts_feat_from_img': (_i32, [_vp, _vp, _i64, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    'pcacc_static_obj_partitioning': (_i32, [_vp, _vp, _i64, _i32, _dbl, _vp, _vp, _vp, _vp]),
    'pcacc_elevation_map': (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp]),
    'pcacc_velo2frame': (_i32, [_vp, _vp, _i32, _i64, _i32, _vp, _vp, _vp]),
    'pcacc_preprocess_pc': (_i32, [_vp, _vp, _i64, _i32, _vp, _dbl, _dbl, _dbl, _dbl, _dbl, _i32,
                                   _vp, _vp, _vp]),
    'pcacc_cell_stats': (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _vp,
                                _vp, _vp, _vp]),
    'pcacc_partition_semantic_pc': (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp]),
    'pcacc_dirichlet_expectation': (_i32, [_vp, _vp, _i32, _i64, _dbl, _vp]),
    'pcacc_road_marking': (_i32, [_vp, _vp, _i64, _dbl, _dbl, _dbl, _i32, _vp, _vp]),
    'pcacc_profile': (_i32, [_vp, _i32]),
    'pcacc_profile_read': (_i32, [_vp, C.POINTER(_dbl * 10), C.POINTER(_i64 * 10)]),
}
KERNEL_CLASSES = ('integrate', 'rebase', 'mark_dynamic', 'bev_bin', 'scan', 'bev_scatter',
                  'bev_reduce', 'export', 'bev_reduce_big', 'bev_classify')

_lib = None


def load():
    """Loads libpcacc.so and types every exported symbol. Raises ImportError
    when the library is missing — there is nothing to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f'{LIB_PATH} not built: run `make -C {os.path.join(_HERE, "csrc")}`'
            ' (or __graft_entry__.build()); this package has no CPU fallback')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the ABI drifted
        fn.restype = res
        fn.argtypes = args
    if lib.pcacc_abi_version() != ABI_VERSION:
        raise ImportError('libpcacc ABI version mismatch')
    _lib = lib
    return lib


def check(status, handle=None):
    if status == OK:
        return
    lib = load()
    detail = lib.pcacc_last_error(handle).decode(errors='replace')
    if not detail:
        detail = lib.pcacc_strerror(status).decode()
    raise PcaccError(status, detail)
