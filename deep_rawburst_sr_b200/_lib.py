"""ctypes binding of libdbsr_b200.so (the C ABI declared in include/dbsr_b200.h).

There is no CPU fallback: if the shared object is missing or the device is not an sm_100 part, loading
raises and every op of the package is unusable.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libdbsr_b200.so')

DBSR_F32, DBSR_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2
CORR_AUTO, CORR_CUDA_CORES = 0, 1


class NhwcView(ctypes.Structure):
    """struct dbsr_nhwc (include/dbsr_b200.h)."""
    _fields_ = [('data', ctypes.c_void_p), ('n', ctypes.c_int32), ('h', ctypes.c_int32), ('w', ctypes.c_int32),
                ('c', ctypes.c_int32), ('c_off', ctypes.c_int32), ('c_pitch', ctypes.c_int32),
                ('dtype', ctypes.c_int32)]


class ConvDesc(ctypes.Structure):
    """struct dbsr_conv (include/dbsr_b200.h)."""
    _fields_ = [('x', NhwcView), ('y', NhwcView), ('residual', NhwcView), ('w', ctypes.c_void_p),
                ('bias', ctypes.c_void_p), ('ksize', ctypes.c_int32), ('stride', ctypes.c_int32),
                ('dilation', ctypes.c_int32), ('act', ctypes.c_int32), ('shuffle_r', ctypes.c_int32),
                ('grid_limit', ctypes.c_int32), ('residual_group', ctypes.c_int32), ('flags', ctypes.c_int32)]


class ResBlockDesc(ctypes.Structure):
    """struct dbsr_resblock (include/dbsr_b200.h)."""
    _fields_ = [('x', NhwcView), ('y', NhwcView), ('w1', ctypes.c_void_p), ('b1', ctypes.c_void_p), ('w2', ctypes.c_void_p),
                ('b2', ctypes.c_void_p), ('pred_w', ctypes.c_void_p), ('pred_b', ctypes.c_void_p), ('pred', ctypes.c_void_p),
                ('pred_c', ctypes.c_int32), ('pred_q14', ctypes.c_int32), ('grid_limit', ctypes.c_int32),
                ('flags', ctypes.c_int32)]


_VP = ctypes.c_void_p
_I = ctypes.c_int32
_F = ctypes.c_float
_PV = ctypes.POINTER(NhwcView)
_PC = ctypes.POINTER(ConvDesc)
_PR = ctypes.POINTER(ResBlockDesc)

# name -> (restype, argtypes); must list every symbol of include/dbsr_b200.h (tests check this)
PROTOTYPES = {
    'dbsr_version': (_I, []),
    'dbsr_last_error': (ctypes.c_char_p, []),
    'dbsr_device_check': (_I, [_I]),
    'dbsr_nchw_to_nhwc': (_I, [_VP, _PV, _VP]),
    'dbsr_nhwc_to_nchw': (_I, [_PV, _VP, _VP]),
    'dbsr_copy_channels': (_I, [_PV, _PV, _I, _I, _I, _VP]),
    'dbsr_prep_burst': (_I, [_VP, _I, _I, _I, _PV, _PV, _VP]),
    'dbsr_prep_burst_s2d': (_I, [_VP, _I, _I, _I, _I, _I, _PV, _PV, _VP]),
    'dbsr_conv2d_direct': (_I, [_PC, _VP]),
    'dbsr_conv2d_tc': (_I, [_PC, _VP]),
    'dbsr_conv2d_tc_supported': (_I, [_PC]),
    'dbsr_conv2d_tc_predictor': (_I, [_PC, _VP, _VP, _I, _VP, _I, _VP]),
    'dbsr_resblock32_tc': (_I, [_PR, _VP]),
    'dbsr_resblock32_tc_supported': (_I, [_PR]),
    'dbsr_quantize_q14': (_I, [_VP, _VP, ctypes.c_int64, _VP]),
    'dbsr_conv2d_tc_geometry': (_I, [_I, _I, ctypes.POINTER(_I), ctypes.POINTER(_I), ctypes.POINTER(_I),
                                     ctypes.POINTER(_I)]),
    'dbsr_space_to_depth2': (_I, [_PV, _PV, _VP]),
    'dbsr_deconv4x4s2': (_I, [_PV, _VP, _VP, _PV, _PV, _VP]),
    'dbsr_deconv_col2im': (_I, [_PV, _VP, _PV, _PV, _VP, _VP, _PV, _PV, _VP]),
    'dbsr_deconv_col2im_ftaps': (_I, [_PV, _VP, _PV, _PV, _PV, _VP, _VP, _VP, _PV, _PV, _VP]),
    'dbsr_flow_from_taps': (_I, [_PV, _VP, _PV, _VP]),
    'dbsr_corr81': (_I, [_PV, _PV, _PV, _F, _PV, _I, _I, _I, _I, _VP]),
    'dbsr_corr81_copy': (_I, [_PV, _PV, _PV, _F, _PV, _PV, _I, _I, _I, _I, _VP]),
    'dbsr_flow_head': (_I, [_PV, _VP, _I, _I, _I, _I, _VP]),
    'dbsr_warp': (_I, [_PV, _VP, _PV, _I, _VP]),
    'dbsr_offsets_mod': (_I, [_VP, _PV, _I, _I, _F, _VP]),
    'dbsr_build_wp_input': (_I, [_PV, _PV, _I, _VP]),
    'dbsr_warp_proj': (_I, [_PV, _VP, _VP, _PV, _I, _VP]),
    'dbsr_warp_proj_split': (_I, [_PV, _VP, _VP, _PV, _PV, _I, _VP]),
    'dbsr_softmax_wsum': (_I, [_PV, _PV, _VP, _PV, _VP, _I, _VP]),
    'dbsr_blur3x3': (_I, [_PV, _PV, ctypes.POINTER(ctypes.c_float), _VP]),
    'dbsr_predictor': (_I, [_PV, _VP, _VP, _I, _VP, _VP]),
    'dbsr_ssim_workspace_floats': (_I, [_I, _I, _I, _I, _I, _I]),
    'dbsr_ssim': (_I, [_VP, _VP, _I, _I, _I, _I, _I, ctypes.POINTER(ctypes.c_float), _I, _F, _VP, _VP, _VP, _VP, _VP]),
    'dbsr_avgpool2_pair': (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _VP]),
    'dbsr_unprocess_rgb': (_I, [_VP, _VP, _I, _I, _I, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), _I, _I, _VP]),
    'dbsr_mosaic_noise': (_I, [_VP, _VP, _VP, _I, _I, _I, _F, _F, _VP]),
    'dbsr_single2lrburst': (_I, [_VP, _I, _I, _VP, _VP, _I, _I, _I, _I, _VP, _VP, _VP]),
    'dbsr_unprocess_rgb_batch': (_I, [_VP, _VP, _I, _I, _I, _VP, _I, _I, _VP]),
    'dbsr_single2lrburst_batch': (_I, [_VP, _I, _I, _I, _VP, _VP, _I, _I, _I, _I, _VP, _VP, _VP]),
    'dbsr_mosaic_noise_batch': (_I, [_VP, _VP, _VP, _I, _I, _I, _VP, _I, _VP]),
    'dbsr_mse_workspace_floats': (_I, [_I]),
    'dbsr_mse_per_image': (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _I, _VP, _VP, _VP]),
}

_lock = threading.Lock()
_lib = None


class DbsrB200Error(RuntimeError):
    pass


def load_library(path: str | None = None) -> ctypes.CDLL:
    """dlopen the C-ABI library and attach prototypes.  No GPU is needed to load it."""
    global _lib
    with _lock:
        if _lib is not None and path is None:
            return _lib
        p = path or os.environ.get('DBSR_B200_LIB') or LIB_PATH   # DBSR_B200_LIB: A/B an alternative build of the same ABI
        if not os.path.exists(p):
            raise DbsrB200Error(
                f'{p} not found: build it with `python -m deep_rawburst_sr_b200.build` (nvcc, sm_100a). '
                'deep_rawburst_sr_b200 has no CPU / PyTorch fallback path.')
        lib = ctypes.CDLL(p)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        if path is None:
            _lib = lib
        return lib


def last_error() -> str:
    return load_library().dbsr_last_error().decode('utf-8', 'replace')


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise DbsrB200Error(f'{what} failed (rc={rc}): {last_error()}')
