"""
ctypes binding of ``libnemo_b200.so`` (C ABI declared in include/nemo_b200.h).

There is no CPU fallback: if the shared library is missing, or a compute entry point is called without a
CUDA device, this module raises.  PyTorch is used only for device memory (tensor handles) and streams.
"""

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get("NEMO_B200_LIB") or os.path.join(HERE, "libnemo_b200.so")   # the override is for kernel experiments

PREC_F32 = 0
PREC_F64 = 1

c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_dbl = ctypes.c_double
c_ptr = ctypes.c_void_p

# name -> argtypes (return type is int unless listed in _RESTYPES); must match include/nemo_b200.h
_PROTOTYPES = {
    "nb_last_error": [],
    "nb_version": [],
    "nb_device_info": [ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_int)],
    "nb_fft_size_supported": [c_int],
    "nb_launch_count": [],
    "nb_launch_count_add": [c_ll],
    "nb_prof_enable": [c_int],
    "nb_prof_report": [ctypes.c_char_p, c_ll],
    "nb_rfft2": [c_int, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_dbl, c_ptr],
    "nb_irfft2": [c_int, c_ptr, c_ptr, c_int, c_int, c_int, c_dbl, c_ptr, c_ptr, c_int, c_ptr],
    "nb_apply_filter": [c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_int, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr],
    "nb_window_deconv": [c_int, c_ptr, c_ptr, c_int, c_int, c_int, c_ptr, c_ptr, c_int, c_ptr],
    "nb_xpower_smooth": [c_int, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_dbl, c_ptr],
    "nb_build_filter": [c_int, c_ptr, c_ptr, ctypes.POINTER(c_dbl), c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_int, c_ptr],
    "nb_build_filter_fused": [c_int, c_ptr, c_ptr, ctypes.POINTER(c_dbl), c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_int, c_dbl, c_ptr],
    "nb_fold_filter": [c_int, c_ptr, c_ptr, c_int, c_int, c_int, c_ptr],
    "nb_template_norm": [c_int, c_ptr, c_int, c_int, c_int, c_int, c_dbl, c_dbl, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    "nb_plane_values_at": [c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr],
    "nb_widen_f32": [c_ptr, c_ptr, c_ll, c_ptr],
    "nb_rms_area_table": [c_int, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_ptr],
    "nb_paint_objects": [c_int, c_ptr, c_ptr, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_ptr],
    "nb_add_maps": [c_int, c_ptr, c_ptr, c_ptr, c_ll, c_ll, c_ptr],
    "nb_xmatch_keep_best": [c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_dbl, c_ptr, c_ptr, c_ptr],
    "nb_xmatch_scratch_ints": [c_int, c_dbl],
    "nb_smart_rms": [c_int, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    "nb_median_weights": [c_int, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_ptr],
    "nb_cell_rms": [c_int, c_ptr, c_ptr, c_int, c_int, c_int, c_int, ctypes.POINTER(c_int), c_int, ctypes.POINTER(c_int), c_int,
                    c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr],
    "nb_cell_rms_est": [c_int, c_ptr, c_ptr, c_int, c_int, c_int, c_int, ctypes.POINTER(c_int), c_int, ctypes.POINTER(c_int), c_int,
                        c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr],
    "nb_snr_mask": [c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    "nb_snr_mask_cells": [c_int, c_ptr, c_ptr, c_ptr, c_ptr, ctypes.POINTER(c_int), c_int, ctypes.POINTER(c_int), c_int, c_int, c_int, c_ptr, c_ptr,
                          c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    "nb_snr_from_rms": [c_int, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_ptr],
    "nb_label4": [c_int, c_ptr, c_dbl, c_dbl, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_ptr],
    "nb_label4_scratch_ints": [c_int, c_int, c_int],
    "nb_label_moments": [c_int, c_ptr, c_dbl, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_ptr],
    "nb_label_shapes": [c_int, c_ptr, c_dbl, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_ptr],
    "nb_detect": [c_int, c_ptr, c_dbl, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_int, c_int, c_int, c_int, c_ptr],
    "nb_detect_scratch_ints": [c_int, c_int, c_int],
    "nb_find_objects": [c_int, c_ptr, c_dbl, c_dbl, c_int, c_dbl, c_dbl, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_int, c_int,
                        c_int, c_int, c_ptr],
    "nb_find_objects_scratch_ints": [c_int, c_int, c_int],
    "nb_ring_flags": [c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_dbl, c_int, c_ptr, c_ptr],
    "nb_spline_eval": [c_int, c_ptr, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_int, c_dbl, c_ptr],
    "nb_gauss_filter": [c_int, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_dbl, c_dbl, c_int, c_ptr],
    "nb_conv2d": [c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_int, c_dbl, c_ptr, c_ptr, c_ptr],
    "nb_conv2d_batch": [c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_int, c_dbl, c_ptr, c_ptr,
                        c_int, c_int, c_int, c_int, c_ptr],
    "nb_mean_over_rows": [c_int, c_ptr, c_ptr, c_int, c_int, c_int, c_ptr],
    "nb_cast": [c_int, c_ptr, c_int, c_ptr, c_ll, c_ptr],
    "nb_mask_to_u8": [c_int, c_ptr, c_ptr, c_ll, c_ptr],
}
_RESTYPES = {"nb_last_error": ctypes.c_char_p, "nb_xmatch_scratch_ints": c_ll, "nb_label4_scratch_ints": c_ll, "nb_detect_scratch_ints": c_ll, "nb_find_objects_scratch_ints": c_ll, "nb_prof_report": c_ll, "nb_launch_count": c_ll, "nb_launch_count_add": c_ll}

EXPORTED_SYMBOLS = sorted(_PROTOTYPES.keys())

_lib = None


class NemoB200Error(Exception):
    pass


def load():
    """Loads the shared library (once) and returns the ctypes handle.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise NemoB200Error("libnemo_b200.so is missing (%s): build it with `python -m nemo_b200._build` "
                            "(or __graft_entry__.build()); there is no CPU fallback." % LIBPATH)
    lib = ctypes.CDLL(LIBPATH)
    for name, argtypes in _PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = lib
    return lib


def call(name, *args):
    """Calls an int-returning entry point and raises NemoB200Error with nb_last_error() on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.nb_last_error()
        raise NemoB200Error("%s failed (%d): %s" % (name, rc, msg.decode() if msg else "?"))
    return rc


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise NemoB200Error("nemo_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    return torch


def launchCount():
    """Kernels launched by libnemo_b200.so in this process so far."""
    return int(load().nb_launch_count())


_PROF_ON = [False]


def profEnable(on=True):
    """Starts (and clears) or stops the library's per-kernel CUDA-event timers."""
    load().nb_prof_enable(1 if on else 0)
    _PROF_ON[0] = bool(on)


def profEnabled():
    return _PROF_ON[0]


def profReport():
    """{kernel name: (launches, total ms)} since the last profEnable(True); synchronises the device."""
    lib = load()
    cap = 1 << 16
    buf = ctypes.create_string_buffer(cap)
    n = lib.nb_prof_report(buf, cap)
    if n < 0:
        raise NemoB200Error("nb_prof_report failed (%d)" % n)
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split(',')
        out[name] = (int(cnt), float(ms))
    return out
