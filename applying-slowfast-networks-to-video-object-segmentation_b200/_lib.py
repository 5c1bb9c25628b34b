"""ctypes binding of libsfvos.so (include/sfvos.h).  No torch types cross this boundary: only raw device
pointers, sizes and the CUDA stream handle.  There is no CPU fallback: if the library is missing or the device
is not sm_100 every op raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsfvos.so")

F32, BF16, F64 = 0, 1, 2
i64, i32, f32, f64, vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_double, ctypes.c_void_p


class ConvParams(ctypes.Structure):
    _fields_ = [("x", vp), ("B", i64), ("T", i64), ("H", i64), ("W", i64), ("C", i64), ("x_cstride", i64),
                ("x_hstride", i64), ("x_tstride", i64), ("x_bstride", i64),
                ("w", vp), ("Cp", i64), ("N", i64), ("kt", i64), ("kh", i64), ("kw", i64),
                ("pad_t", i64), ("pad_h", i64), ("pad_w", i64), ("To", i64),
                ("y", vp), ("y_dtype", i32), ("relu", i32), ("y_cstride", i64),
                ("scale", vp), ("shift", vp), ("sum", vp), ("sumsq", vp),
                ("accumulate", i32), ("reserved", i32),
                ("OH", i64), ("OW", i64), ("oy_mul", i64), ("oy_off", i64), ("ox_mul", i64), ("ox_off", i64),
                ("relu_mask", vp), ("relu_mask_cstride", i64),
                ("addend", vp), ("addend_dtype", i32), ("reserved2", i32), ("addend_cstride", i64)]


class WgradParams(ctypes.Structure):
    _fields_ = [("x", vp), ("B", i64), ("T", i64), ("H", i64), ("W", i64), ("C", i64), ("x_cstride", i64),
                ("x_hstride", i64), ("x_tstride", i64), ("x_bstride", i64),
                ("dy", vp), ("To", i64), ("N", i64), ("dy_cstride", i64),
                ("dy_hstride", i64), ("dy_tstride", i64), ("dy_bstride", i64),
                ("kt", i64), ("kh", i64), ("kw", i64), ("pad_t", i64), ("pad_h", i64), ("pad_w", i64),
                ("dw", vp), ("workspace", vp), ("workspace_bytes", i64)]


class RoiParams(ctypes.Structure):
    _fields_ = [("feat", vp * 4), ("dfeat", vp * 4), ("H", i64 * 4), ("W", i64 * 4), ("scale", f32 * 4),
                ("n_levels", i32), ("feat_dtype", i32), ("N", i64), ("C", i64), ("cstride", i64),
                ("rois", vp), ("levels", vp), ("K", i64), ("P", i32), ("sampling_ratio", i32),
                ("out", vp), ("out_dtype", i32), ("out_nchw", i32)]


BN_MAX_CALLS = 8


class BnRunningParams(ctypes.Structure):
    _fields_ = [("sum", vp * BN_MAX_CALLS), ("sumsq", vp * BN_MAX_CALLS), ("count", f64 * BN_MAX_CALLS),
                ("n_calls", i32), ("stats_dtype", i32), ("conv_bias", vp), ("running_mean", vp), ("running_var", vp),
                ("num_batches_tracked", vp), ("momentum", f64), ("C", i64)]


_SIGS = {
    "sfvos_version": [],
    "sfvos_device_check": [],
    "sfvos_tma_overlap_supported": [vp],
    "sfvos_probe_l2": [i32, vp, i64, i32, vp, vp],
    "sfvos_conv_umma": [ctypes.POINTER(ConvParams), vp],
    "sfvos_conv_simt": [ctypes.POINTER(ConvParams), vp],
    "sfvos_wgrad_umma": [ctypes.POINTER(WgradParams), vp],
    "sfvos_wgrad_simt": [ctypes.POINTER(WgradParams), vp],
    "sfvos_pack_weights": [vp, vp, i32, i32, i64, i64, i64, i64, i64, i64, i64, i64, vp],
    "sfvos_unpack_wgrad": [vp, vp, i32, i64, i64, i64, i64, i64, i64, i64, vp],
    "sfvos_channel_stats": [vp, i64, i64, i64, vp, vp, i64, vp],
    "sfvos_bn_finalize": [vp, vp, i32, f64, vp, vp, vp, vp, vp, vp, f64, f64, vp, vp, vp, vp, i64, vp],
    "sfvos_bn_running_update": [ctypes.POINTER(BnRunningParams), vp],
    "sfvos_bn_fold_eval": [vp, vp, vp, vp, vp, f64, vp, vp, vp, vp, i64, vp],
    "sfvos_affine_act": [vp, i32, i64, vp, i32, i64, vp, vp, i32, i64, i64, vp],
    "sfvos_bn_bwd_reduce": [vp, i32, i64, vp, i64, vp, vp, vp, vp, i32, i64, i64, vp, vp, i64, vp],
    "sfvos_bn_bwd_apply": [vp, i32, i64, vp, i64, vp, vp, vp, vp, vp, i32, i64, i64, vp, vp, i32, i64, vp, vp, i32, vp, vp],
    "sfvos_relu_bwd": [vp, i32, i64, vp, i32, i64, vp, i32, i64, vp, i64, i64, vp],
    "sfvos_nchw_to_nhwc": [vp, i32, i64, vp, i32, i64, i64, i64, i64, vp],
    "sfvos_nhwc_to_nchw": [vp, i32, i64, vp, i64, i64, i64, vp],
    "sfvos_roi_levels": [vp, i64, i32, i32, vp, vp],
    "sfvos_roi_align_fwd": [ctypes.POINTER(RoiParams), vp],
    "sfvos_roi_align_bwd": [ctypes.POINTER(RoiParams), vp],
    "sfvos_mask_targets": [vp, i64, i64, i64, vp, i64, i32, vp, vp],
    "sfvos_mask_logits_fwd": [vp, i32, vp, vp, vp, i64, i64, i64, i32, i32, vp],
    "sfvos_mask_bce_fwd": [vp, vp, vp, vp, i64, i64, i32, vp],
    "sfvos_mask_logits_bwd": [vp, i32, vp, vp, vp, i32, vp, vp, i64, i64, i64, i32, i32, vp],
    "sfvos_mask_logits_relu_bwd": [vp, i32, vp, vp, vp, i32, vp, vp, vp, i64, i64, i64, i32, i32, vp],
    "sfvos_mask_bce_bwd": [vp, vp, vp, vp, vp, i64, i64, i32, vp],
    "sfvos_mask_probs": [vp, vp, vp, i64, i64, i32, vp],
    "sfvos_fastrcnn_loss_fwd": [vp, i64, vp, i64, vp, vp, i64, i32, f32, vp, vp],
    "sfvos_fastrcnn_loss_bwd": [vp, i64, vp, i64, vp, vp, vp, i64, i32, f32, vp, i64, vp, i64, vp],
    "sfvos_paste_masks": [vp, vp, i64, i32, i32, i64, i64, vp, vp],
    "sfvos_upsample_add": [vp, i64, i64, vp, vp, i64, i64, i64, i64, vp],
    "sfvos_im2col": [vp, vp, i32, i64, i64, i64, i64, i64, i64, i64, i64, i64, vp],
    "sfvos_maxpool3x3s2": [vp, vp, i32, i64, i64, i64, i64, vp],
    "sfvos_add_relu": [vp, vp, vp, vp, i64, vp],
    "sfvos_axpby": [vp, vp, f32, f32, i64, vp],
}

# size queries: host-only, return int64 byte counts
_SIZE_SIGS = {
    "sfvos_reduce_workspace_bytes": [i64, i64],
    "sfvos_wgrad_simt_workspace_bytes": [ctypes.POINTER(WgradParams)],
}

EXPORTED = sorted(list(_SIGS) + list(_SIZE_SIGS) + ["sfvos_last_error", "sfvos_last_kernel"])
_lib = None


def load():
    """Load libsfvos.so (built in-tree by build.py).  Raises if it is missing -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python {os.path.join(HERE, 'build.py')}` "
                           "(the CUDA extension is mandatory; there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = ctypes.c_int
    for name, args in _SIZE_SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = ctypes.c_int64
    lib.sfvos_last_error.argtypes = []
    lib.sfvos_last_error.restype = ctypes.c_char_p
    lib.sfvos_last_kernel.argtypes = []
    lib.sfvos_last_kernel.restype = ctypes.c_char_p
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError(f"libsfvos error {rc}: {load().sfvos_last_error().decode()}")


LAUNCHES = 0   # number of kernel-launching C-ABI calls made by this process (bench.py reports it)


def last_kernel():
    return load().sfvos_last_kernel().decode()


def call(name, *args):
    global LAUNCHES
    LAUNCHES += 1
    check(getattr(load(), name)(*args))
