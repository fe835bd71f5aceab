"""ctypes binding of libpf_b200.so (the C ABI declared in include/pf_b200.h).

There is no fallback: if the shared library is missing or a call fails, this raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpf_b200.so")

_lib = None


class ConvInfo(C.Structure):
    _fields_ = [("cin", C.c_int), ("cout", C.c_int), ("ksize", C.c_int), ("stride", C.c_int),
                ("name", C.c_char * 64)]


class PFError(RuntimeError):
    pass


_vp, _i, _sz, _f = C.c_void_p, C.c_int, C.c_size_t, C.c_float

SIGNATURES = {
    "pf_version": (C.c_int, []),
    "pf_last_error": (C.c_char_p, []),
    "pf_zsplat_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "pf_zsplat_forward": (_i, [_vp] * 8 + [_i] * 5 + [_vp] * 5 + [_sz, _vp]),
    "pf_zsplat_forward_frames": (_i, [_vp] * 8 + [_i] * 5 + [_vp] * 5 + [_sz, _vp]),
    "pf_zsplat_forward_frames_hop": (_i, [_vp] * 8 + [_i] * 4 + [_vp] * 4 + [_f, _f, _vp, _sz, _vp]),
    "pf_zsplat_forward_frames_hop_packed": (_i, [_vp] * 9 + [_i] * 4 + [_vp] * 4 + [_f, _f, _vp, _sz, _vp]),
    "pf_zsplat_forward_host": (_i, [_vp] * 8 + [_i] * 5 + [_vp] * 3),
    "pf_zsplat_launches_per_forward": (_i, []),
    "pf_zsplat_launches_for": (_i, [_i, _i, _i, _i]),
    "pf_depth_disk_hop": (_i, [_vp, _vp, _vp, _sz, _f, _f, _vp]),
    "pf_bgnet_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i]),
    "pf_bgnet_destroy": (None, [_vp]),
    "pf_bgnet_num_convs": (_i, [_vp]),
    "pf_bgnet_conv_info": (_i, [_vp, _i, C.POINTER(ConvInfo)]),
    "pf_bgnet_load_conv": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _f]),
    "pf_bgnet_load_final": (_i, [_vp, _vp, _vp]),
    "pf_bgnet_set_depth_norm": (_i, [_vp, _f, _f]),
    "pf_bgnet_workspace_bytes": (_sz, [_vp, _i, _i, _i]),
    "pf_bgnet_forward": (_i, [_vp] * 4 + [_i] * 5 + [_vp] * 5 + [_sz, _vp]),
    "pf_bgnet_forward_dense": (_i, [_vp] * 4 + [_i] * 5 + [_vp] * 5 + [_sz, _vp]),
    "pf_bgnet_launches_per_forward": (_i, [_vp]),
    "pf_bgnet_set_profiling": (_i, [_vp, _i]),
    "pf_bgnet_num_steps": (_i, [_vp]),
    "pf_bgnet_step_info": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_i)]),
    "pf_bgnet_read_profile": (_i, [_vp, C.POINTER(_f), _i]),
    "pf_bgnet_debug_conv": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp]),
    "pf_upsample_argmax": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "pf_panoptic_paint_order": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "pf_panoptic_merge": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
}


def lib():
    """Loads the library once; raises if it has not been built (python -m panoptic_forecasting_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PFError("libpf_b200.so not built: run `python -m panoptic_forecasting_b200.build` "
                          "(there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().pf_last_error().decode("utf-8", "replace")
        raise PFError("%s failed (rc=%d): %s" % (what, rc, msg))


def ptr(t):
    """data_ptr of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()
