"""The network through the three-call C entry points (include/lwpose_b200.h: lwp_net_load / lwp_net_forward /
lwp_net_heads) -- what a non-Python host uses.  The blob is produced once per (checkpoint, precision, batch, height,
width) by `export()` and cached on disk keyed by a hash of the state_dict: it is the persistent pre-folded,
pre-packed form of the weights (reference modules/load_state.py:4-15 loads a checkpoint into the module; here the
module's state_dict is turned into the blob, and `CNet` never touches the module again).

    blob = cnet.export(net, "bf16", 64, 368, 656)         # or cnet.cached_blob(net, ...) -> path
    c = cnet.CNet(blob)
    c.forward(x_cuda)                                       # NCHW float32 [64,3,368,656]
    heads = c.heads()                                       # float32 [64,46,82,64] view of the library's buffer
"""
import ctypes
import hashlib
import os

from . import _lib


def state_hash(net, extra=""):
    """sha256 over the module's state_dict (names, shapes, bytes) + `extra`."""
    import torch
    h = hashlib.sha256()
    for k, v in net.state_dict().items():
        h.update(k.encode())
        h.update(str(tuple(v.shape)).encode())
        h.update(v.detach().to("cpu", torch.float32).contiguous().numpy().tobytes())
    h.update(extra.encode())
    return h.hexdigest()


def export(net, precision, n, H, W, input_u8=None):
    """Blob bytes for lwp_net_load: BatchNorm folded, weights packed, layer list recorded (needs the module on a GPU)."""
    plan = net.engine().new_plan(precision, n, H, W, input_u8=input_u8)
    return plan.export_blob()


def cached_blob(net, precision, n, H, W, input_u8=None, cache_dir=None):
    """Path of the blob for this (checkpoint, precision, shape), written on first use."""
    cache_dir = cache_dir or os.environ.get("LWP_CACHE_DIR") or os.path.join(os.path.expanduser("~"), ".cache", "lwpose_b200")
    key = state_hash(net, "%s:%d:%d:%d:%r:v%d" % (precision, n, H, W, input_u8, _lib.load().lwp_version()))
    path = os.path.join(cache_dir, "%s.lwpb" % key[:32])
    if not os.path.exists(path):
        os.makedirs(cache_dir, exist_ok=True)
        tmp = path + ".tmp%d" % os.getpid()
        with open(tmp, "wb") as f:
            f.write(export(net, precision, n, H, W, input_u8=input_u8))
        os.replace(tmp, path)
    return path


class _DevMem:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class CNet:
    """ctypes handle of a lwp_net (the library owns all device memory)."""

    def __init__(self, blob):
        self.L = _lib.load()
        _lib.require_cuda()
        if isinstance(blob, str):
            with open(blob, "rb") as f:
                blob = f.read()
        self._h = _lib._c_void_p()
        buf = ctypes.create_string_buffer(blob, len(blob))
        _lib.check(self.L.lwp_net_load(ctypes.cast(buf, ctypes.c_void_p), len(blob), self._h), "lwp_net_load")
        vals = [ctypes.c_int() for _ in range(5)]
        _lib.check(self.L.lwp_net_info(self._h, *vals), "lwp_net_info")
        self.dtype, self.n, self.H, self.W, self.n_stages = [v.value for v in vals]

    def close(self):
        if getattr(self, "_h", None):
            self.L.lwp_net_destroy(self._h)
            self._h = None

    __del__ = close

    def forward(self, x, with_nchw=False):
        assert x.is_cuda and x.is_contiguous()
        _lib.check(self.L.lwp_net_forward(self._h, x.data_ptr(), int(bool(with_nchw)), _lib.current_stream()), "lwp_net_forward")
        return self

    def heads(self, stage=-1):
        """float32 [n, H/8, W/8, 64] tensor aliasing the library's head buffer of `stage` (valid until the next forward)."""
        import torch
        ptr, ld = _lib._c_void_p(), ctypes.c_int()
        _lib.check(self.L.lwp_net_heads(self._h, stage, ptr, ld), "lwp_net_heads")
        return torch.as_tensor(_DevMem(ptr.value, (self.n, self.H // 8, self.W // 8, ld.value), "<f4"), device="cuda")

    def output_nchw(self, index, channels):
        import torch
        ptr = _lib._c_void_p()
        _lib.check(self.L.lwp_net_output_nchw(self._h, index, ptr), "lwp_net_output_nchw")
        return torch.as_tensor(_DevMem(ptr.value, (self.n, channels, self.H // 8, self.W // 8), "<f4"), device="cuda")
