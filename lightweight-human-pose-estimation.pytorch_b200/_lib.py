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
    "lwp_timing_experiments": (_c_int, []),
    "lwp_check_device": (_c_int, [_c_int]),
    "lwp_resize_pad_u8": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                   _c_double, _c_double, _c_int, _c_int, _c_int, _c_void_p]),
    "lwp_upsample_cubic_ex": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, ctypes.c_longlong, ctypes.c_longlong,
                                       _c_void_p, _c_int, _c_int, _c_double, _c_double, ctypes.c_float, _c_void_p]),
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
    "lwp_paf_pack_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "lwp_group_keypoints": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int,
                                     _c_int, _c_int, _c_double, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p,
                                     _c_size_t, _c_void_p, _c_void_p]),
    "lwp_pose_convert": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_int, _c_int, _c_double, _c_double,
                                  _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "lwp_copy_flagged": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_size_t, _c_void_p]),
    "lwp_net_load": (_c_int, [_c_void_p, _c_size_t, ctypes.POINTER(_c_void_p)]),
    "lwp_net_destroy": (None, [_c_void_p]),
    "lwp_net_info": (_c_int, [_c_void_p] + [ctypes.POINTER(_c_int)] * 5),
    "lwp_net_forward": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_void_p]),
    "lwp_net_heads": (_c_int, [_c_void_p, _c_int, ctypes.POINTER(_c_void_p), ctypes.POINTER(_c_int)]),
    "lwp_net_output_nchw": (_c_int, [_c_void_p, _c_int, ctypes.POINTER(_c_void_p)]),
    "lwp_postprocess_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int, _c_int, _c_int]),
    "lwp_postprocess": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_double, _c_void_p, _c_void_p,
                                 _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_size_t,
                                 _c_void_p, _c_void_p]),
    "lwp_plan_create": (_c_int, [_c_int, ctypes.POINTER(_c_void_p)]),
    "lwp_plan_destroy": (None, [_c_void_p]),
    "lwp_plan_num_ops": (_c_int, [_c_void_p]),
    "lwp_plan_add_stem": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int]),
    "lwp_plan_add_stem_u8": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int,
                                      ctypes.POINTER(_c_double), _c_double]),
    "lwp_plan_add_frontend": (_c_int, [_c_void_p] + [_c_void_p] * 13 + [_c_int, _c_int, _c_int, _c_int, ctypes.POINTER(_c_double),
                                                              _c_double]),
    "lwp_plan_add_conv3x3_pw": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int,
                                         _c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_int,
                                         _c_int, _c_int]),
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
    "lwp_plan_add_sepconv": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p,
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
        refuse_debug_env(L)
        _lib = L
    return _lib


def refuse_debug_env(L=None):
    """The LWP_DEBUG_* switches make kernels skip loads / MMAs / epilogues (timing experiments: results are wrong by
    design).  A release library ignores them, but a run that has them set is refused anyway, so that no timed or
    product run can ever be taken with one of them in the environment; only a library built with
    -DLWP_TIMING_EXPERIMENTS (lwp_timing_experiments() == 1) together with LWP_ALLOW_TIMING_EXPERIMENTS=1 accepts them."""
    bad = sorted(k for k, v in os.environ.items() if k.startswith("LWP_DEBUG_") and v not in ("", "0"))
    if not bad:
        return
    L = L if L is not None else _lib
    experiments = L is not None and L.lwp_timing_experiments() == 1 and os.environ.get("LWP_ALLOW_TIMING_EXPERIMENTS") == "1"
    if not experiments:
        raise LwpError("refusing to run with %s set: these switches skip work inside the kernels (timing experiments "
                       "only; needs a -DLWP_TIMING_EXPERIMENTS build and LWP_ALLOW_TIMING_EXPERIMENTS=1)" % ", ".join(bad))


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
