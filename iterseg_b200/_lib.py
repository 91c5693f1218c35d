"""ctypes binding of libiterseg_b200.so (the C-ABI in include/iterseg_b200.h).

There is no CPU fallback: if the library is missing the import of any compute
entry point raises, and every compute call requires a CUDA device.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, 'libiterseg_b200.so')

_c = ctypes
_vp, _i64, _i32, _sz, _f32 = _c.c_void_p, _c.c_int64, _c.c_int, _c.c_size_t, _c.c_float


class PostParams(_c.Structure):
    _fields_ = [('aff_ch', _i32 * 3), ('mask_ch', _i32), ('cent_ch', _i32), ('r1', _i32),
                ('r2', _i32), ('peak_thresh', _f32), ('use_absolute_thresh', _i32),
                ('absolute_thresh', _f32), ('min_area', _i64), ('max_area', _i64),
                ('scale', _f32 * 3), ('use_aff_div', _i32), ('aff_div', _f32 * 3), ('own_z0', _i32),
                ('own_z1', _i32), ('open_faces', _i32), ('seed_keys_out', _vp)]


DOG_MAX_SIGMAS, GAUSS_MAX_RADIUS = 9, 27


class DogParams(_c.Structure):
    _fields_ = [('weights', (_c.c_double * 12) * 4), ('radius', _i32 * 4), ('threshold', _f32),
                ('scale_factor', _f32), ('prune_d2', _i32), ('prune_radius', _i32),
                ('n_layers', _i32), ('overlap', _c.c_double), ('layer_sigma', _c.c_double * DOG_MAX_SIGMAS),
                ('layer_radius', _i32 * DOG_MAX_SIGMAS),
                ('layer_weights', (_c.c_double * (GAUSS_MAX_RADIUS + 1)) * DOG_MAX_SIGMAS),
                ('mask_radius', _i32 * 2), ('mask_weights', (_c.c_double * (GAUSS_MAX_RADIUS + 1)) * 2)]


# name -> (restype, argtypes); this table is also what the symbol-export test checks
SIGNATURES = {
    'isg_version': (_i32, []),
    'isg_last_error': (_c.c_char_p, []),
    'isg_device_check': (_i32, []),
    'isg_launch_count': (_c.c_uint64, []),
    'isg_set_post_sm_reservation': (_i32, [_i32]),
    'isg_frame_minmax': (_i32, [_vp, _i64, _vp, _vp, _sz, _vp]),
    'isg_frame_divide_by_max': (_i32, [_vp, _i64, _vp, _vp]),
    'isg_flood_workspace_bytes': (_sz, [_i64, _i64, _i64, _i64]),
    'isg_affinity_flood': (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i64,
                                  _vp, _vp, _sz, _vp]),
    'isg_post_workspace_bytes': (_sz, [_i64, _i64, _i64, _i64]),
    'isg_post_workspace_bytes_capped': (_sz, [_i64, _i64, _i64, _i64, _i64]),
    'isg_segment_features': (_i32, [_vp, _i32, _i64, _i64, _i64, _c.POINTER(PostParams), _vp, _vp,
                                    _vp, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    'isg_slab_stats': (_i32, [_vp, _i32, _i64, _i64, _i64, _c.POINTER(PostParams), _vp, _i32, _vp, _vp, _vp,
                              _vp, _sz, _vp]),
    'isg_otsu_from_hist': (_i32, [_vp, _vp, _vp, _vp, _sz, _vp]),
    'isg_sort_tmp_bytes': (_sz, [_i64]),
    'isg_sort_keys_u64': (_i32, [_vp, _i64, _vp, _sz, _vp]),
    'isg_relabel_by_keys': (_i32, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _vp]),
    'isg_dog_workspace_bytes': (_sz, [_i64, _i64, _i64, _i64]),
    'isg_dog_workspace_bytes_layers': (_sz, [_i64, _i64, _i64, _i64, _i32]),
    'isg_dog_blob_segment': (_i32, [_vp, _i64, _i64, _i64, _c.POINTER(DogParams), _vp, _vp, _vp, _i64, _vp,
                                    _vp, _sz, _vp]),
    'isg_metrics_workspace_bytes': (_sz, [_i64, _i64]),
    'isg_label_metrics': (_i32, [_vp, _vp, _i64, _i64, _c.c_double, _vp, _vp, _sz, _vp]),
    'isg_unet_packed_weight_bytes': (_sz, []),
    'isg_unet_weights_pack': (_i32, [_vp, _i32, _vp, _vp]),
    'isg_unet_workspace_bytes': (_sz, [_i32, _i32, _i32, _i32]),
    'isg_unet_plan_create': (_vp, [_vp, _i32, _i32, _i32, _i32, _i64, _i64, _i64, _vp, _vp, _vp,
                                   _vp, _sz]),
    'isg_unet_plan_set_chunks': (_i32, [_vp, _vp, _vp, _vp]),
    'isg_unet_plan_destroy': (None, [_vp]),
    'isg_debug_flat_tiling': (_i32, [_i32, _i32, _i32, _i64, _vp]),
    'isg_unet_plan_overflowed': (_i32, [_vp]),
    'isg_unet_plan_clear_overflow': (_i32, [_vp, _vp]),
    'isg_unet_forward_chunks': (_i32, [_vp, _vp, _vp, _vp]),
    'isg_unet_forward_chunks_norm': (_i32, [_vp, _vp, _vp, _vp, _vp]),
    'isg_unet_debug_activation': (_i32, [_vp, _vp, _c.c_char_p, _i32, _vp, _i64, _vp]),
    'isg_unet_plan_flops': (_c.c_double, [_vp]),
    'isg_unet_plan_profile': (_i32, [_vp, _i32]),
    'isg_unet_plan_profile_read': (_i32, [_vp, _vp]),
    'isg_unet_plan_profile_launches': (_i32, [_vp, _vp, _vp, _i32]),
    'isg_unet_plan_profile_timeline': (_i32, [_vp, _vp, _i32]),
    'isg_add_label_offset': (_i32, [_vp, _i64, _c.c_uint32, _vp]),
    'isg_crop_labels': (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _vp]),
}

_lib = None


class IsgError(RuntimeError):
    status = None            # the C-ABI status code (ISG_ERR_*) when the error came from the library


ISG_ERR_CUDA, ISG_ERR_ARG, ISG_ERR_WORKSPACE, ISG_ERR_OVERFLOW, ISG_ERR_DEVICE = 1, 2, 3, 4, 5


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IsgError(
            f'{LIB_PATH} is missing: build it with `python -m iterseg_b200._build` '
            '(there is no CPU fallback for this path)')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != 0:
        msg = load().isg_last_error().decode('utf-8', 'replace')
        err = IsgError(f'{what} failed (status {rc}): {msg}')
        err.status = int(rc)
        raise err


def require_device():
    """Raise unless a compute-capability-10.x CUDA device is usable."""
    import torch
    if not torch.cuda.is_available():
        raise IsgError('iterseg_b200 needs a CUDA (sm_100a) device; there is no CPU fallback')
    check(load().isg_device_check(), 'isg_device_check')


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
