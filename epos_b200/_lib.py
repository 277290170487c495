"""ctypes binding of the C ABI in include/epos_b200.h.

There is no CPU fallback: if the CUDA library is missing or a call fails, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('EPOS_B200_LIB') or os.path.join(_HERE, 'csrc', 'libepos_b200.so')   # override: developer A/B builds
_lib = None

vp, i32, i64, u64, f32, f64, sz = C.c_void_p, C.c_int, C.c_longlong, C.c_ulonglong, C.c_float, C.c_double, C.c_size_t


class FitParams(C.Structure):
    """epos_fit_params (include/epos_b200.h)."""
    _fields_ = [('threshold', f64), ('spatial_coherence_weight', f64), ('neighborhood_ball_radius', f64),
                ('scaling_from_millimeters', f64), ('min_triangle_area', f64), ('min_coverage', f64),
                ('max_iters', C.c_int32), ('min_iters', C.c_int32), ('min_iters_before_lo', C.c_int32),
                ('max_lo_trials', C.c_int32), ('max_graph_cuts', C.c_int32), ('max_lsq_iters', C.c_int32),
                ('max_unsuccessful', C.c_int32), ('max_neighbors', C.c_int32),
                ('apply_numerical_optimization', C.c_int32), ('reserved', C.c_int32)]


class MultiParams(C.Structure):
    """epos_multi_params (include/epos_b200.h)."""
    _fields_ = [('max_model_number_for_pearl', C.c_int32), ('min_point_number', C.c_int32), ('confidence', f64),
                ('max_tanimoto_similarity', f64)]


_SIGS = {
    'epos_last_error': (C.c_char_p, []),
    'epos_version': (i32, []),
    'epos_compiled_arch': (i32, []),
    'epos_launch_count': (u64, []),
    'epos_conv3x3_rgb_s2': (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    'epos_maxpool3x3_s2': (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
    'epos_subsample_f32': (i32, [vp, i32, vp, i32, i32, i32, i32, i32, vp]),
    'epos_conv3x3_gemm': (i32, [vp, i32, sz, vp, i32, vp, vp, i32, vp, i32, vp, i32, sz, i32, i32, i32, i32, i32, i32, i32, vp]),
    'epos_conv3x3_dense': (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    'epos_dwconv3x3': (i32, [vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    'epos_pwconv_gemm': (i32, [vp, i32, sz, vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, sz, i32, i32, i32, i32, vp]),
    'epos_gemm_pieces': (i32, [i32, i32, i32, i32, i32, vp, i32]),
    'epos_pwconv_simt': (i32, [vp, i32, vp, vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, vp]),
    'epos_split_bf16': (i32, [vp, i32, vp, i32, sz, i32, i32, i32, i32, i32, i32, vp]),
    'epos_global_mean': (i32, [vp, vp, i32, i32, i32, vp]),
    'epos_resize_bilinear': (i32, [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]),
    'epos_softmax_rows': (i32, [vp, vp, sz, i32, vp]),
    'epos_softmax_rows_masked': (i32, [vp, vp, sz, i32, i32, f32, vp]),
    'epos_preprocess_u8': (i32, [vp, i32, i32, sz, i32, i32, i32, i32, i32, vp, vp, vp, vp]),
    'epos_corresp': (i32, [vp, vp, vp, i32, i32, i32, i32, i32, vp, i32, vp, vp, f64, f32, f32, i32, i32,
                           vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    'epos_corresp_workspace_bytes': (sz, [i32, i32, i32, i32]),
    'epos_corresp_lazy_loc': (i32, [vp, vp, vp, i32, sz, i32, vp, vp, i32, i32, i32, i32, i32, vp, i32, vp, vp, f64, f32,
                                    f32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    'epos_fit_max_points': (i32, []),
    'epos_fit_debug_state': (i32, [vp, i32, vp]),
    'epos_fit_debug_trace': (i32, [vp, i32, i32, vp]),
    'epos_fit_enable_timing': (i32, [i32]),
    'epos_fit_last_kernel_ms': (i32, [C.POINTER(f32), C.POINTER(f32)]),
    'epos_fit_params_default': (None, [C.POINTER(FitParams)]),
    'epos_fit_poses': (i32, [vp, vp, vp, vp, i32, vp, vp, C.POINTER(FitParams), vp, vp, vp, sz, vp]),
    'epos_fit_workspace_bytes': (sz, [i32, i32, C.POINTER(FitParams)]),
    'epos_fit_poses_multi': (i32, [vp, vp, vp, vp, i32, vp, vp, C.POINTER(FitParams), C.POINTER(MultiParams), vp, vp, vp,
                                   vp, vp, vp, vp, sz, vp]),
    'epos_fit_multi_workspace_bytes': (sz, [i32]),
    'epos_fit_max_instances': (i32, []),
    'epos_fit_multi_debug_state': (i32, [vp, i32, vp]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('epos_b200: %s not built (run `python -c "import __graft_entry__ as g; g.build()"`); '
                               'there is no CPU fallback' % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)     # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def exported_symbols():
    return sorted(_SIGS.keys())


def check(rc, what):
    if rc != 0:
        raise RuntimeError('%s failed (%d): %s' % (what, rc, lib().epos_last_error().decode()))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
