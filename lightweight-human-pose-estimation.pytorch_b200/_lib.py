"""ctypes binding of the C-ABI library (include/lwpose_b200.h).  There is no fallback: if the shared
library is missing or a call fails, an exception is raised."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblwpose_b200.so")

_c_int = ctypes.c_int
_c_void_p = ctypes.c_void_p
_c_size_t = ctypes.c_size_t
_c_double = ctypes.c_double

# name -> (restype, argtypes); must list every symbol include/lwpose_b200.h declares
SIGNATURES = {
    "lwp_version": (_c_int, []),
    "lwp_last_error": (ctypes.c_char_p, []),
    "lwp_check_device": (_c_int, [_c_int]),
    "lwp_upsample_cubic": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_int, _c_int,
                                    _c_double, _c_double, _c_void_p]),
    "lwp_extract_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "lwp_extract_keypoints": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p,
                                       _c_void_p, _c_int, _c_int, _c_void_p, _c_size_t, _c_void_p, _c_void_p]),
    "lwp_extract_keypoints_fused": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                             _c_double, _c_double, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int,
                                             _c_void_p, _c_size_t, _c_void_p, _c_void_p]),
    "lwp_group_keypoints_fused": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int,
                                           _c_int, _c_int, _c_int, _c_double, _c_double, _c_int, _c_double, _c_void_p,
                                           _c_void_p, _c_int, _c_int, _c_void_p, _c_size_t, _c_void_p, _c_void_p]),
    "lwp_group_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int, _c_int]),
    "lwp_group_keypoints": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int,
                                     _c_int, _c_int, _c_double, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p,
                                     _c_size_t, _c_void_p, _c_void_p]),
    "lwp_plan_create": (_c_int, [_c_int, ctypes.POINTER(_c_void_p)]),
    "lwp_plan_destroy": (None, [_c_void_p]),
    "lwp_plan_num_ops": (_c_int, [_c_void_p]),
    "lwp_plan_add_stem": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int]),
    "lwp_plan_add_stem_u8": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int,
                                      ctypes.POINTER(_c_double), _c_double]),
    "lwp_plan_add_depthwise": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int,
                                        _c_int, _c_int, _c_int, _c_int, _c_int, _c_int]),
    "lwp_plan_add_heads_fused": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p,
                                          _c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int]),
    "lwp_plan_add_conv_gemm": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                        _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int,
                                        _c_int, _c_int, _c_int, _c_int]),
    "lwp_plan_add_dwpw": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p,
                                   _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int,
                                   _c_int, _c_int, _c_int]),
    "lwp_plan_add_nhwc_to_nchw": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_int,
                                           _c_int, _c_int]),
    "lwp_plan_run": (_c_int, [_c_void_p, _c_void_p, _c_void_p]),
    "lwp_plan_run_range": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_void_p]),
    "lwp_plan_num_launches": (_c_int, [_c_void_p]),
    "lwp_plan_error_flag": (_c_int, [_c_void_p]),
}

_lib = None


class LwpError(RuntimeError):
    pass


def load():
    """Load liblwpose_b200.so (built in-tree by __graft_entry__.build()).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LwpError("%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` first; "
                           "lwpose_b200 has no CPU or PyTorch fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().lwp_last_error()
        raise LwpError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else ""))


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise LwpError("lwpose_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    check(load().lwp_check_device(torch.cuda.current_device()), "lwp_check_device")
