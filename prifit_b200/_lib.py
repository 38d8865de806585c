"""ctypes binding of libprifit_b200.so -- the C ABI declared in include/prifit_b200.h.

There is no CPU fallback: if the library is missing, or a call fails, this raises.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libprifit_b200.so")

_p = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_sz = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/prifit_b200.h one to one
SIGNATURES = {
    "prifit_version": (_i, []),
    "prifit_set_gram_engine": (_i, [_i]),
    "prifit_last_error_string": (ctypes.c_char_p, []),
    "prifit_device_ok": (_i, []),
    "prifit_normalize_fwd": (_i, [_p, _i64, _i, _p, _p]),
    "prifit_normalize_bwd": (_i, [_p, _p, _i64, _i, _p, _p]),
    "prifit_normalize_fwd_cf": (_i, [_p, _i, _i, _i, _p, _p]),
    "prifit_normalize_bwd_cf": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "prifit_normalize_bwd_scaled": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "prifit_bandwidth_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "prifit_bandwidth_fwd": (_i, [_p, _i, _i, _i, _p, _i, _p, _p, _p, _sz, _p]),
    "prifit_meanshift_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "prifit_meanshift_fwd": (_i, [_p, _p, _i, _i, _i, _i, _p, _i, _p, _sz, _p]),
    "prifit_nms_workspace_bytes": (_sz, [_i, _i, _i]),
    "prifit_nms_fwd": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _sz, _p]),
    "prifit_nms_labels": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _sz, _p]),
    "prifit_meanshift_rows_prepare": (_i, [_p, _i, _i, _i, _i, _p, _sz, _p]),
    "prifit_meanshift_rows_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "prifit_meanshift_rows_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _i, _p, _sz, _p]),
    "prifit_meanshift_rows_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _i, _p, _sz, _p]),
    "prifit_membership_workspace_bytes": (_sz, [_i, _i, _i]),
    "prifit_membership_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "prifit_membership_bwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "prifit_membership_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "prifit_fit_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "prifit_fit_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "prifit_sdf_workspace_bytes": (_sz, [_i, _i]),
    "prifit_sdf_loss_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "prifit_sdf_loss_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p]),
    "prifit_intersect_workspace_bytes": (_sz, [_i, _i]),
    "prifit_intersect_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _sz, _p]),
    "prifit_intersect_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "prifit_noise_scatter": (_i, [_p, _p, _i, _i, _p, _p, _p]),
    "prifit_noise_scatter_range": (_i, [_p, _p, _i, _i, _i, _p, _p, _p]),
    "prifit_pack_counts": (_i, [_p, _p, _i, _p, _p, _p]),
    "prifit_spin_until_ge": (_i, [_p, _i, _p]),
    "prifit_masked_mean_fwd": (_i, [_p, _p, _i, _i, _p, _p, _p]),
    "prifit_masked_mean_bwd": (_i, [_p, _p, _p, _p, _i, _p, _p]),
    "prifit_sample_counts": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "prifit_sample_surface": (_i, [_p, _p, _i, _i, _i, ctypes.c_uint64, _p, _p, _p, _p]),
    "prifit_surface_points_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p]),
    "prifit_surface_points_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "prifit_nn_workspace_bytes": (_sz, [_i, _i]),
    "prifit_nn_loss_fwd": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "prifit_nn_loss_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "prifit_entropy_workspace_bytes": (_sz, [_i, _i]),
    "prifit_entropy_fwd": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "prifit_entropy_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "prifit_fps": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "prifit_ball_query": (_i, [_p, _p, _i, _i, _i, ctypes.c_float, _i, _p, _p]),
    "prifit_three_nn": (_i, [_p, _p, _i, _i, _i, _p, _p, _p]),
    "prifit_interpolate_fwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "prifit_interpolate_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "prifit_debug_tc_probe": (_i, [_p, _p, _i, _i, _i, _p, _p, _p]),
    "prifit_debug_tc_gram": (_i, [_p, _i, _i, _p, _p, _p]),
}

FIT_CTX = 48
MS_F16_TCGEN05 = 0
MS_FP32_SIMT = 1
ROWS_SPLIT_TCGEN05 = 0
ROWS_FP32_SIMT = 1
ROWS_WS_HOLDS_SPLIT = 0x100
ROWS_WIDE = 0x200
ROWS_NARROW = 0x400

_lib = None


class PrifitError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and declare every prototype.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PrifitError(
            "libprifit_b200.so not found at %s -- build it with `python -m prifit_b200.build` "
            "(there is no CPU or PyTorch fallback for this path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().prifit_last_error_string().decode("utf-8", "replace")
        raise PrifitError("%s failed with code %d: %s" % (what, rc, msg))


# kernels (and memset nodes) each entry point enqueues; bench.py reports the sum as `gpu_launches`
LAUNCHES = {
    "prifit_normalize_fwd": 1, "prifit_normalize_bwd": 1, "prifit_normalize_fwd_cf": 1, "prifit_normalize_bwd_cf": 1, "prifit_normalize_bwd_scaled": 1, "prifit_bandwidth_fwd": 6, "prifit_meanshift_fwd": 3,
    "prifit_nms_fwd": 11, "prifit_nms_labels": 2, "prifit_meanshift_rows_prepare": 1, "prifit_meanshift_rows_fwd": 2, "prifit_meanshift_rows_bwd": 2,
    "prifit_membership_fwd": 2, "prifit_membership_bwd": 2, "prifit_fit_fwd": 1, "prifit_fit_bwd": 1,
    "prifit_sdf_loss_fwd": 2, "prifit_sdf_loss_bwd": 1,
    "prifit_masked_mean_fwd": 1, "prifit_masked_mean_bwd": 1, "prifit_noise_scatter": 1, "prifit_noise_scatter_range": 1, "prifit_pack_counts": 1, "prifit_spin_until_ge": 1,
    "prifit_intersect_fwd": 2, "prifit_intersect_bwd": 1, "prifit_entropy_fwd": 2, "prifit_entropy_bwd": 1, "prifit_nn_loss_fwd": 2, "prifit_nn_loss_bwd": 1,
}
_launches = 0


def reset_launch_count():
    global _launches
    _launches = 0


def launch_count():
    return _launches


# bench.py sets this to a list to collect (entry point, start event, end event) around every C-ABI call of the eager path
# (CUDA events on the launching stream -- the stream handle is the last argument of every entry point)
TIMING = None


def call(name, *args, launches=None):
    """launches: kernels this particular call enqueues when that differs from the entry point's usual count."""
    global _launches
    if TIMING is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(getattr(load(), name)(*args), name)
        e1.record()
        TIMING.append((name, e0, e1))
    else:
        check(getattr(load(), name)(*args), name)
    _launches += LAUNCHES.get(name, 0) if launches is None else launches
