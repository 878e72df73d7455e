"""ctypes binding of include/ganslate_b200.h.

The CUDA library is the only compute path: importing this module fails loudly when the shared object is
missing (run `python -c "import __graft_entry__ as g; g.build()"`), there is no CPU or eager fallback.
"""
import ctypes as C
from pathlib import Path

GB_MAX_TAPS = 128
GB_MAX_CLASSES = 8
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_TANH, ACT_PRELU = 0, 1, 2, 3, 4

LIB_PATH = Path(__file__).resolve().parent / "_lib" / "libganslate_b200.so"


class View(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("sn", C.c_int64), ("sz", C.c_int64), ("sy", C.c_int64), ("sx", C.c_int64),
                ("N", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
                ("pad", C.c_int32)]


class ConvClass(C.Structure):
    _fields_ = [("off", C.c_int32 * 3), ("ntaps", C.c_int32), ("tap_begin", C.c_int32), ("kpad", C.c_int32),
                ("w_offset", C.c_int64)]


class ConvParams(C.Structure):
    _fields_ = [("inp", View), ("out", View), ("wpacked", C.c_void_p), ("bias", C.c_void_p), ("ncols", C.c_int32),
                ("npad", C.c_int32), ("in_mul", C.c_int32 * 3), ("out_mul", C.c_int32 * 3), ("nclass", C.c_int32),
                ("cls", ConvClass * GB_MAX_CLASSES), ("taps", (C.c_int8 * 4) * GB_MAX_TAPS), ("act", C.c_int32),
                ("act_slope", C.c_float), ("out_fp32", C.c_int32), ("accumulate", C.c_int32), ("stats", C.c_void_p),
                ("in_c_valid", C.c_int32)]


class WgradParams(C.Structure):
    _fields_ = [("plain", View), ("gathered", View), ("dw", C.c_void_p), ("rows", C.c_int32), ("kpad", C.c_int32),
                ("ntaps", C.c_int32), ("mul", C.c_int32 * 3), ("taps", (C.c_int8 * 4) * GB_MAX_TAPS),
                ("splits", C.c_int32), ("gathered_c_valid", C.c_int32)]


class PackParams(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("sn", C.c_int64), ("sc", C.c_int64), ("st", C.c_int64),
                ("rows", C.c_int32), ("rows_pad", C.c_int32), ("chans", C.c_int32), ("chans_pad", C.c_int32),
                ("nclass", C.c_int32), ("ntaps", C.c_int32 * GB_MAX_CLASSES), ("kpad", C.c_int32 * GB_MAX_CLASSES),
                ("w_offset", C.c_int64 * GB_MAX_CLASSES), ("tap_begin", C.c_int32 * GB_MAX_CLASSES),
                ("tap_id", C.c_int32 * GB_MAX_TAPS)]


GB_UNPACK_BATCH = 56


class UnpackItem(C.Structure):
    _fields_ = [("dw", C.c_void_p), ("dst", C.c_void_p), ("dsr", C.c_int64), ("dsc", C.c_int64), ("dst_t", C.c_int64),
                ("rows", C.c_int32), ("chans", C.c_int32), ("chans_pad", C.c_int32), ("ntaps", C.c_int32),
                ("kpad", C.c_int32), ("accumulate", C.c_int32)]


class UnpackBatch(C.Structure):
    _fields_ = [("count", C.c_int32), ("pad_", C.c_int32), ("item", UnpackItem * GB_UNPACK_BATCH)]


GB_ADAM_BATCH = 96


class AdamItem(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_int32),
                ("vec4", C.c_int32)]


class AdamBatch(C.Structure):
    _fields_ = [("lr", C.c_void_p), ("step", C.c_void_p), ("beta1", C.c_float), ("beta2", C.c_float),
                ("eps", C.c_float), ("count", C.c_int32), ("item", AdamItem * GB_ADAM_BATCH)]


class InFwdParams(C.Structure):
    _fields_ = [("x", View), ("y", View), ("res", View), ("stats", C.c_void_p), ("prelu", C.c_void_p),
                ("eps", C.c_float), ("act", C.c_int32), ("act_slope", C.c_float), ("res_before_act", C.c_int32),
                ("out_scale", C.c_float)]


class InBwdParams(C.Structure):
    _fields_ = [("x", View), ("y", View), ("dy_a", View), ("dy_b", View), ("dy_sum", View), ("dx", View),
                ("stats", C.c_void_p), ("bstats", C.c_void_p), ("prelu", C.c_void_p), ("dprelu", C.c_void_p),
                ("dbias", C.c_void_p), ("eps", C.c_float), ("act", C.c_int32), ("act_slope", C.c_float),
                ("res", View), ("res_before_act", C.c_int32), ("dy_sum_acc", C.c_int32), ("dx_fp32_acc", C.c_int32),
                ("out_scale", C.c_float)]


_SIGNATURES = {
    "gb_conv_data": [C.POINTER(ConvParams), C.c_void_p],
    "gb_conv_wgrad": [C.POINTER(WgradParams), C.c_void_p],
    "gb_pack_weights": [C.POINTER(PackParams), C.c_void_p],
    "gb_pack_weights_multi": [C.c_void_p, C.c_int, C.c_int64, C.c_void_p],
    "gb_unpack_wgrad_multi": [C.POINTER(UnpackBatch), C.c_void_p],
    "gb_unpack_wgrad": [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                        C.c_int, C.c_void_p],
    "gb_adam_multi": [C.POINTER(AdamBatch), C.c_void_p],
    "gb_colsum": [C.POINTER(View), C.c_void_p, C.c_void_p],
    "gb_in_stats": [C.POINTER(View), C.c_void_p, C.c_void_p],
    "gb_in_fwd": [C.POINTER(InFwdParams), C.c_void_p],
    "gb_in_bwd": [C.POINTER(InBwdParams), C.c_void_p],
    "gb_nchw_to_cl": [C.c_void_p, C.c_int, C.POINTER(View), C.POINTER(View), C.c_int, C.c_void_p],
    "gb_cl_to_nchw": [C.POINTER(View), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p],
    "gb_replicate_pad_fwd": [C.POINTER(View), C.POINTER(View), C.c_int, C.c_int, C.c_int, C.c_void_p],
    "gb_replicate_pad_bwd": [C.POINTER(View), C.POINTER(View), C.c_int, C.c_int, C.c_int, C.c_void_p],
    "gb_mse_const": [C.c_void_p, C.c_float, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p],
    "gb_l1": [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p],
    "gb_ssim_fwd": [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p,
                    C.c_void_p],
    "gb_ssim_bwd": [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p,
                    C.c_void_p, C.c_void_p],
    "gb_patchnce_fwd": [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p],
    "gb_patchnce_bwd": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p],
    "gb_patch_mlp_fwd": [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                         C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "gb_patch_mlp_bwd": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int,
                         C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                         C.c_void_p, C.c_void_p, C.c_void_p],
    "gb_version": [],
    "gb_tma_window_supported": [],
    "gb_debug_knob": [C.c_int, C.c_int],
    "gb_debug_timeline": [C.c_void_p, C.c_longlong],
    "gb_debug_cg2_watchdog": [C.c_void_p],
    "gb_workspace_bytes": [C.c_int, C.c_void_p, C.POINTER(C.c_int64)],
    "gb_debug_cg2_plan": [C.POINTER(ConvParams), C.c_int, C.c_void_p, C.c_void_p, C.c_int64],
}

_lib = None


def exported_symbols():
    """Every entry point include/ganslate_b200.h declares (checked by the CPU test-suite)."""
    return list(_SIGNATURES) + ["gb_last_error", "gb_launch_count"]


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: the CUDA extension is the only compute path of ganslate_b200. "
                               "Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
        L = C.CDLL(str(LIB_PATH))
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        L.gb_last_error.argtypes = []
        L.gb_last_error.restype = C.c_char_p
        L.gb_launch_count.argtypes = []
        L.gb_launch_count.restype = C.c_ulonglong
        _lib = L
        # bring-up knobs: GB_KNOBS="4=1,1=128" -> gb_debug_knob(4, 1); gb_debug_knob(1, 128)
        import os
        for item in filter(None, os.environ.get("GB_KNOBS", "").split(",")):
            k, v = item.split("=")
            L.gb_debug_knob(int(k), int(v))
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {lib().gb_last_error().decode()}")
