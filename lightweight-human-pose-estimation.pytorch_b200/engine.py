"""Host side of the network forward: folds BatchNorm, packs weights for the tcgen05 kernels and records
one launch plan per (precision, batch, height, width) through the C ABI (include/lwpose_b200.h).

It walks the module tree of lwpose_b200.models.with_mobilenet.PoseEstimationWithMobileNet -- the same
parameters, in the same containers, as the reference's models/with_mobilenet.py:89-123 -- and replaces
its forward (:114-123):

  model (stem + 11 depthwise-separable blocks)  -> stem kernel, depthwise kernels, 1x1 implicit GEMMs
  cpm (:18-21)                                  -> 1x1 GEMM, 3 x (depthwise+ELU, 1x1 GEMM+ELU), residual in
                                                   the last epilogue, 3x3 GEMM
  initial_stage (:41-45)                        -> 3 x 3x3 GEMM, both heads merged into two GEMMs
  refinement_stages (:57-60,82-86,119-121)      -> the 185-channel concat is never materialised by a cat:
                                                   producers write straight into a 192-wide buffer
PyTorch is used for device memory and one-off weight re-layout only.
"""
import collections
import os

import torch
from torch import nn

from . import _lib

DTYPE_BF16, DTYPE_TF32 = 0, 1
ACT_NONE, ACT_RELU, ACT_ELU = 0, 1, 2
_PREC = {"bf16": (DTYPE_BF16, torch.bfloat16), "tf32": (DTYPE_TF32, torch.float32)}
HEAD_LD = 64      # float32 head buffer: 19 heat-map + 38 PAF channels + 7 zero pad
CONCAT_LD = 192   # 128 backbone + 19 + 38 + 7 zero pad


class _Unit:
    """One Conv2d with its BatchNorm / bias folded to per-channel scale & shift and its activation."""

    def __init__(self, conv_m, bn_m, act_m):
        w = conv_m.weight.detach().float()
        cout = w.shape[0]
        dev = w.device
        scale = torch.ones(cout, device=dev)
        shift = torch.zeros(cout, device=dev)
        if conv_m.bias is not None:
            shift = conv_m.bias.detach().float().clone()
        if bn_m is not None:  # eval-mode BN: y = (x - mean) / sqrt(var + eps) * gamma + beta
            inv = bn_m.weight.detach().float() / torch.sqrt(bn_m.running_var.detach().float() + bn_m.eps)
            shift = (shift - bn_m.running_mean.detach().float()) * inv + bn_m.bias.detach().float()
            scale = inv
        self.weight, self.scale, self.shift = w, scale, shift
        self.act = ACT_RELU if isinstance(act_m, nn.ReLU) else ACT_ELU if isinstance(act_m, nn.ELU) else ACT_NONE
        self.stride, self.dilation = conv_m.stride[0], conv_m.dilation[0]
        self.depthwise = conv_m.groups > 1
        self.ksize = conv_m.kernel_size[0]
        if self.ksize == 3 and conv_m.padding[0] != self.dilation:
            raise ValueError("3x3 conv with padding != dilation is not part of the hot path")


def _units(seq):
    """Split a Sequential of Conv2d [BatchNorm2d] [ReLU|ELU] ... into folded units."""
    mods = list(seq.children())
    out, i = [], 0
    while i < len(mods):
        assert isinstance(mods[i], nn.Conv2d), type(mods[i])
        conv_m, bn_m, act_m = mods[i], None, None
        i += 1
        if i < len(mods) and isinstance(mods[i], nn.BatchNorm2d):
            bn_m = mods[i]
            i += 1
        if i < len(mods) and isinstance(mods[i], (nn.ReLU, nn.ELU)):
            act_m = mods[i]
            i += 1
        out.append(_Unit(conv_m, bn_m, act_m))
    return out


def _pad_rows(t, rows):
    if t.shape[0] == rows:
        return t
    pad = torch.zeros((rows - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    return torch.cat([t, pad], 0)


class _GemmW:
    """Weights of one implicit-GEMM conv: [Cout_pad][taps*Cin] K-major in the plan dtype + fp32 scale/shift."""

    def __init__(self, weight, scale, shift, act, tdtype, dilation=1, cin_pad=None):
        cout, cin, kh, kw = weight.shape
        self.taps = kh * kw
        w = weight.permute(0, 2, 3, 1).reshape(cout, self.taps, cin)
        self.cin_real = cin
        if cin_pad is not None and cin_pad != cin:
            w = torch.cat([w, torch.zeros(cout, self.taps, cin_pad - cin, device=w.device)], 2)
            cin = cin_pad
        self.cin, self.cout = cin, cout
        self.cout_pad = (cout + 63) // 64 * 64
        self.w = _pad_rows(w.reshape(cout, self.taps * cin), self.cout_pad).to(tdtype).contiguous()
        self.scale = _pad_rows(scale.reshape(-1), self.cout_pad).float().contiguous()
        self.shift = _pad_rows(shift.reshape(-1), self.cout_pad).float().contiguous()
        self.act, self.dilation = act, dilation

    @staticmethod
    def from_unit(u, tdtype, cin_pad=None):
        return _GemmW(u.weight, u.scale, u.shift, u.act, tdtype, u.dilation, cin_pad)


class _DwW:
    def __init__(self, u):
        c = u.weight.shape[0]
        self.w = u.weight.reshape(c, 9).t().contiguous()  # [9][C] tap-major
        self.scale, self.shift = u.scale.contiguous(), u.shift.contiguous()
        self.act, self.stride, self.dilation, self.c = u.act, u.stride, u.dilation, c


def _merge_heads(hm_seq, paf_seq, tdtype):
    """Both heads read the same trunk features: first layers stacked along Cout, second layers as one
    block-diagonal GEMM writing [19 heat-maps | 38 PAFs | zero pad]."""
    (h1, h2), (p1, p2) = _units(hm_seq[0]) + _units(hm_seq[1]), _units(paf_seq[0]) + _units(paf_seq[1])
    first = _GemmW(torch.cat([h1.weight, p1.weight], 0), torch.cat([h1.scale, p1.scale]),
                   torch.cat([h1.shift, p1.shift]), h1.act, tdtype)
    mid_h, mid_p = h2.weight.shape[1], p2.weight.shape[1]
    nh, np_ = h2.weight.shape[0], p2.weight.shape[0]
    w2 = torch.zeros(nh + np_, mid_h + mid_p, 1, 1, device=h2.weight.device)
    w2[:nh, :mid_h] = h2.weight
    w2[nh:, mid_h:] = p2.weight
    second = _GemmW(w2, torch.cat([h2.scale, p2.scale]), torch.cat([h2.shift, p2.shift]), h2.act, tdtype)
    second.flops_per_px = 2.0 * (mid_h * nh + mid_p * np_)  # algorithmic: the zero blocks do not count
    return first, second


class _Packed:
    """All weights of the network re-laid-out for one precision."""

    def __init__(self, net, tdtype):
        blocks = list(net.model.children())
        stem = _units(blocks[0])[0]
        self.stem = (stem.weight.contiguous(), stem.scale.contiguous(), stem.shift.contiguous())
        self.backbone = []
        for b in blocks[1:]:
            dw, pw = _units(b)
            self.backbone.append((_DwW(dw), _GemmW.from_unit(pw, tdtype)))
        self.cpm_align = _GemmW.from_unit(_units(net.cpm.align)[0], tdtype)
        self.cpm_trunk = []
        for b in net.cpm.trunk.children():
            dw, pw = _units(b)
            self.cpm_trunk.append((_DwW(dw), _GemmW.from_unit(pw, tdtype)))
        self.cpm_conv = _GemmW.from_unit(_units(net.cpm.conv)[0], tdtype)
        self.init_trunk = [_GemmW.from_unit(_units(c)[0], tdtype) for c in net.initial_stage.trunk.children()]
        self.init_heads = _merge_heads(net.initial_stage.heatmaps, net.initial_stage.pafs, tdtype)
        self.refine = []
        for st in net.refinement_stages:
            blks = []
            for i, blk in enumerate(st.trunk.children()):
                ini = _GemmW.from_unit(_units(blk.initial)[0], tdtype, cin_pad=CONCAT_LD if i == 0 else None)
                t0, t1 = [_GemmW.from_unit(_units(c)[0], tdtype) for c in blk.trunk.children()]
                blks.append((ini, t0, t1))
            self.refine.append((blks, _merge_heads(st.heatmaps, st.pafs, tdtype)))


class _RecordingLib:
    """Proxy of the ctypes library that remembers every successful lwp_plan_add_* call (name + arguments after the
    plan handle): Plan.export_blob() serialises that list, so a non-Python host can rebuild the same plan with
    lwp_net_load() without re-implementing this file's layer walk."""

    def __init__(self, lib, log):
        self._lib, self._log = lib, log

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.startswith("lwp_plan_add_"):
            return fn

        def call(handle, *args):
            rc = fn(handle, *args)
            if rc == 0:
                self._log.append((name, args))
            return rc
        return call


# function ids of the blob format (csrc/net_blob.cu must agree)
_BLOB_FUNCS = {"lwp_plan_add_stem": 0, "lwp_plan_add_stem_u8": 1, "lwp_plan_add_depthwise": 2, "lwp_plan_add_conv_gemm": 3,
               "lwp_plan_add_dwpw": 4, "lwp_plan_add_sepconv": 5, "lwp_plan_add_heads_fused": 6, "lwp_plan_add_nhwc_to_nchw": 7,
               "lwp_plan_add_frontend": 8, "lwp_plan_add_conv3x3_pw": 9}


def _all_tensors(obj, out, seen):
    """Every torch tensor reachable from the packed-weights object tree."""
    if id(obj) in seen:
        return
    seen.add(id(obj))
    if isinstance(obj, torch.Tensor):
        out.append(obj)
    elif isinstance(obj, (list, tuple)):
        for o in obj:
            _all_tensors(o, out, seen)
    elif hasattr(obj, "__dict__"):
        for o in vars(obj).values():
            _all_tensors(o, out, seen)


class Plan:
    """One recorded launch list + its buffers for a fixed (precision, N, H, W)."""

    def __init__(self, packed, precision, n, H, W, n_stages_out, num_heatmaps, num_pafs, device, input_u8=None,
                 double_heads=False):
        if H % 8 or W % 8:
            raise ValueError("input height/width must be multiples of 8 (got %dx%d)" % (H, W))
        self.call_log = []
        self.lib = _RecordingLib(_lib.load(), self.call_log)
        self.n, self.H, self.W = n, H, W
        self.h, self.w = H // 8, W // 8
        self.precision = precision
        code, tdtype = _PREC[precision]
        self.tdtype = tdtype
        self.device = device
        self.packed = packed  # keeps the packed weights alive
        # input_u8 = (mean3, scale): the stem reads uint8 [n,H,W,3] frames and applies val.normalize on the fly
        self.input_u8 = input_u8
        # double_heads: the last stage's heads op is recorded a second time (at the END of the op list) writing to a second
        # float32 buffer; run_compute(x, alt=1) runs that one instead.  PosePipeline alternates between the two so the
        # post-processing of batch i can still read its heads while the network of batch i + 1 runs -- without a copy.
        self.double_heads = bool(double_heads)
        self.heads_alt = None
        handle = _lib._c_void_p()
        _lib.check(self.lib.lwp_plan_create(code, handle), "lwp_plan_create")
        self.handle = handle
        # depthwise -> 1x1 fusion policy: the two-kernel form (TMA depthwise at ~75 % of HBM peak + tcgen05 GEMM at ~50 %
        # tensor pipe) is faster than the fused block on every layer of this network, so fusion is opt-in:
        # LWP_DWPW_FUSION=1 fuses every block the fused kernel can take, or a comma-separated list of op names
        fuse = os.environ.get("LWP_DWPW_FUSION", "0")
        self.fuse_dwpw = fuse not in ("", "0")
        self.fuse_only = None if fuse in ("", "0", "1", "all") else set(fuse.split(","))
        # weight-resident fused block for the thin layers (sepconv_gemm.cu).  Measured (round 2, 64 x 368x656 bf16): it takes
        # the depthwise intermediate out of HBM (ncu: 451 MB instead of 741 MB of DRAM traffic on model.3) but is
        # instruction-issue-bound on the CUDA cores -- 208 us vs 178 us for TMA depthwise + tcgen05 GEMM on model.3, 95 vs 93 us
        # on the Cpm trunk blocks -- so it is OPT-IN: LWP_SEPCONV=1 uses it wherever it fits, a comma-separated list of op
        # names restricts it (DESIGN.md section 3.2d has the ncu evidence)
        sep = os.environ.get("LWP_SEPCONV", "0")
        self.use_sepconv = sep not in ("", "0")
        self.sepconv_only = None if sep in ("", "0", "1", "all") else set(sep.split(","))
        self.bufs = []
        self.buf_zero = []
        self.op_names = []
        self.op_meta = []  # per op: kind, algorithmic flops and bytes (real channel counts, no padding)
        self._build(packed, n_stages_out, num_heatmaps, num_pafs)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.lwp_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # -- helpers ------------------------------------------------------------------------------
    def _buf(self, *shape, dtype=None, zero=False):
        t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype or self.tdtype, device=self.device)
        self.bufs.append(t)
        self.buf_zero.append(bool(zero))
        return t

    def _gemm(self, name, src, src_ld, g, n, H, W, out=None, out_ld=0, out_ptr_off=0, residual=None, res_ld=0,
              out_f32=None, out_f32_ld=0):
        es = 2 if self.tdtype == torch.bfloat16 else 4
        out_ptr = (out.data_ptr() + out_ptr_off * es) if out is not None else None
        _lib.check(self.lib.lwp_plan_add_conv_gemm(
            self.handle, src.data_ptr(), src_ld, g.w.data_ptr(), g.scale.data_ptr(), g.shift.data_ptr(),
            residual.data_ptr() if residual is not None else None, res_ld, out_ptr, out_ld,
            out_f32.data_ptr() if out_f32 is not None else None, out_f32_ld, n, H, W, g.cin, g.cout, g.taps,
            g.dilation, g.act), "lwp_plan_add_conv_gemm(%s)" % name)
        self.op_names.append(name)
        cin_real = getattr(g, "cin_real", g.cin)
        px = n * H * W
        nbytes = px * (cin_real + g.cout) * es + g.taps * cin_real * g.cout * es
        if residual is not None:
            nbytes += px * g.cout * es
        if out_f32 is not None:
            nbytes += px * g.cout * 4
        self.op_meta.append(dict(kind="gemm3x3" if g.taps == 9 else "gemm1x1", flops=px * getattr(g, "flops_per_px", 2.0 * cin_real * g.cout * g.taps),
                                 bytes=float(nbytes)))

    def _dw(self, name, src, dst, d, n, H, W):
        _lib.check(self.lib.lwp_plan_add_depthwise(self.handle, src.data_ptr(), dst.data_ptr(), d.w.data_ptr(),
                                                   d.scale.data_ptr(), d.shift.data_ptr(), n, H, W, d.c, d.stride,
                                                   d.dilation, d.act), "lwp_plan_add_depthwise(%s)" % name)
        self.op_names.append(name)
        es = 2 if self.tdtype == torch.bfloat16 else 4
        ho, wo = (H - 1) // d.stride + 1, (W - 1) // d.stride + 1
        self.op_meta.append(dict(kind="depthwise", flops=2.0 * 9 * n * ho * wo * d.c,
                                 bytes=float((n * H * W * d.c + n * ho * wo * d.c) * es + 9 * d.c * 4)))

    def _fusable(self, d, g, name=None):
        """Depthwise (stride 1) + pointwise pair that the fused tcgen05 kernel can take (and the policy asks for)."""
        kb_ch = 64 if self.tdtype == torch.bfloat16 else 32
        if self.fuse_only is not None and name not in self.fuse_only:
            return False
        return (self.fuse_dwpw and d.stride == 1 and d.dilation in (1, 2) and g.taps == 1 and d.c == g.cin
                and (d.c % kb_ch == 0 or d.c < kb_ch) and d.c % 8 == 0 and g.cout_pad <= 512
                and 512 % g.cout_pad == 0)

    def _dwpw(self, name, src, d, g, n, H, W, out, out_ld, residual=None, res_ld=0):
        """Record the fused block; returns False (nothing recorded) when its tiles do not fit in shared memory."""
        es = 2 if self.tdtype == torch.bfloat16 else 4
        rc = self.lib.lwp_plan_add_dwpw(
            self.handle, src.data_ptr(), d.w.data_ptr(), d.scale.data_ptr(), d.shift.data_ptr(), d.act, d.dilation,
            g.w.data_ptr(), g.scale.data_ptr(), g.shift.data_ptr(), g.act,
            residual.data_ptr() if residual is not None else None, res_ld, out.data_ptr(), out_ld, n, H, W, d.c,
            g.cout)
        if rc == 3:  # LWP_ECAP: fall back to the two-kernel form of this block
            return False
        _lib.check(rc, "lwp_plan_add_dwpw(%s)" % name)
        self.op_names.append(name)
        px = n * H * W
        nbytes = px * (d.c + g.cout) * es + d.c * g.cout * es + 9 * d.c * 4
        if residual is not None:
            nbytes += px * g.cout * es
        self.op_meta.append(dict(kind="dwpw", flops=2.0 * px * d.c * (g.cout + 9), bytes=float(nbytes)))
        return True

    def _sepconv(self, name, src, d, g, n, H, W, out, out_ld, residual=None, res_ld=0):
        """Record the weight-resident fused depthwise + 1x1 block (thin layers); returns False (nothing recorded) when
        the layer is not eligible or its weights + rings do not fit in shared memory."""
        kb_ch = 64 if self.tdtype == torch.bfloat16 else 32
        if not self.use_sepconv or (self.sepconv_only is not None and name not in self.sepconv_only):
            return False
        if not (d.dilation == 1 and d.stride in (1, 2) and g.taps == 1 and d.c == g.cin and d.c % kb_ch == 0
                and g.cout_pad <= 256):
            return False
        es = 2 if self.tdtype == torch.bfloat16 else 4
        rc = self.lib.lwp_plan_add_sepconv(
            self.handle, src.data_ptr(), d.w.data_ptr(), d.scale.data_ptr(), d.shift.data_ptr(), d.act, d.stride,
            g.w.data_ptr(), g.scale.data_ptr(), g.shift.data_ptr(), g.act,
            residual.data_ptr() if residual is not None else None, res_ld, out.data_ptr(), out_ld, n, H, W, d.c, g.cout)
        if rc == 3:  # LWP_ECAP: two-kernel form
            return False
        _lib.check(rc, "lwp_plan_add_sepconv(%s)" % name)
        self.op_names.append(name)
        ho, wo = (H - 1) // d.stride + 1, (W - 1) // d.stride + 1
        px = n * ho * wo
        nbytes = (n * H * W * d.c + px * g.cout) * es + d.c * g.cout * es + 9 * d.c * 4
        if residual is not None:
            nbytes += px * g.cout * es
        self.op_meta.append(dict(kind="sepconv", flops=2.0 * px * d.c * (g.cout + 9), bytes=float(nbytes)))
        return True

    def _heads(self, name, src, src_ld, heads, n, h, w, big, out, out_ptr_off, out_f32):
        """The two (merged) 1x1 layers of a stage's heads: one back-to-back GEMM kernel on bf16 plans (the intermediate
        never leaves the SM; LWP_HEADS_FUSION=0 keeps the two-kernel form), two conv GEMMs otherwise."""
        g0, g1 = heads
        px = n * h * w
        fuse = (self.tdtype == torch.bfloat16 and os.environ.get("LWP_HEADS_FUSION", "1") != "0" and g0.cin % 64 == 0
                and g0.cout_pad % 64 == 0 and g1.cout_pad == 64 and g1.cin == g0.cout_pad)
        if fuse:
            out_ptr = (out.data_ptr() + out_ptr_off * 2) if out is not None else None
            rc = self.lib.lwp_plan_add_heads_fused(
                self.handle, src.data_ptr(), src_ld, g0.w.data_ptr(), g0.scale.data_ptr(), g0.shift.data_ptr(), g0.cout_pad,
                g1.w.data_ptr(), g1.scale.data_ptr(), g1.shift.data_ptr(), out_ptr, CONCAT_LD, out_f32.data_ptr(), HEAD_LD,
                px, g0.cin)
            if rc != 3:  # LWP_ECAP: tiles too large for shared memory -> two kernels
                _lib.check(rc, "lwp_plan_add_heads_fused(%s)" % name)
                self.op_names.append(name + ".fused")
                flops = px * (2.0 * g0.cin_real * g0.cout + getattr(g1, "flops_per_px", 2.0 * g1.cin_real * g1.cout))
                nbytes = px * (g0.cin_real * 2 + HEAD_LD * 4 + (64 * 2 if out is not None else 0)) + (g0.w.numel() + g1.w.numel()) * 2
                self.op_meta.append(dict(kind="gemm1x1", flops=flops, bytes=float(nbytes)))
                return
        mid = g0.cout_pad
        hb = big[: px * mid]
        self._gemm(name + ".0", src, src_ld, g0, n, h, w, out=hb, out_ld=mid)
        self._gemm(name + ".1", hb, mid, g1, n, h, w, out=out, out_ld=CONCAT_LD, out_ptr_off=out_ptr_off,
                   out_f32=out_f32, out_f32_ld=HEAD_LD)

    # -- the layer walk -----------------------------------------------------------------------
    def _build(self, P, n_stages_out, num_heatmaps, num_pafs):
        n, H, W = self.n, self.H, self.W
        # backbone ping-pong buffers, both sized for the largest activation of the backbone
        biggest = n * (H // 2) * (W // 2) * 32
        hh, ww, c = H // 2, W // 2, 32
        for dw, pw in P.backbone:
            ho, wo = (hh - 1) // dw.stride + 1, (ww - 1) // dw.stride + 1
            biggest = max(biggest, n * ho * wo * c, n * ho * wo * pw.cout_pad)
            hh, ww, c = ho, wo, pw.cout
        pp = [self._buf(biggest), self._buf(biggest)]

        w_, s_, b_ = P.stem
        es_ = 2 if self.tdtype == torch.bfloat16 else 4
        # fused front end (frontend_fused.cu): stem + model.1 (dw + pw) + the depthwise half of model.2 as one kernel on
        # bf16 plans -- the 32- / 64-channel maps at H/2 x W/2 never reach HBM (ncu: 276 MB of DRAM traffic instead of
        # 2.1 GB) and the result is bit-identical.  OPT-IN (LWP_FRONTEND_FUSION=1): measured on 64 x 368x656 it takes 546 us
        # (float frames) / 587 us (uint8) against 471 / 503 us for the four HBM-bound kernels -- it is bound by the CUDA
        # cores' instruction issue (31 k warp instructions per 8x16 output tile at 46 % issue utilisation with the one
        # 512-thread CTA an SM can hold; DESIGN.md section 3.2e, profiles/r02k_*).
        dw1, pw1 = P.backbone[0]
        dw2, _pw2 = P.backbone[1]
        self.frontend_fused = (self.tdtype == torch.bfloat16 and os.environ.get("LWP_FRONTEND_FUSION", "0") not in ("", "0")
                               and not self.fuse_dwpw and not self.use_sepconv
                               and dw1.c == 32 and dw1.stride == 1 and dw1.dilation == 1 and dw1.act == ACT_RELU
                               and pw1.cin == 32 and pw1.cout == 64 and pw1.taps == 1 and pw1.act == ACT_RELU
                               and dw2.c == 64 and dw2.stride == 2 and dw2.dilation == 1 and dw2.act == ACT_RELU)
        if self.frontend_fused:
            mean3, img_scale = self.input_u8 if self.input_u8 is not None else ((0.0, 0.0, 0.0), 1.0)
            mean_arr = (_lib._c_double * 3)(*[float(m) for m in mean3])
            _lib.check(self.lib.lwp_plan_add_frontend(
                self.handle, w_.data_ptr(), s_.data_ptr(), b_.data_ptr(), dw1.w.data_ptr(), dw1.scale.data_ptr(),
                dw1.shift.data_ptr(), pw1.w.data_ptr(), pw1.scale.data_ptr(), pw1.shift.data_ptr(), dw2.w.data_ptr(),
                dw2.scale.data_ptr(), dw2.shift.data_ptr(), pp[0].data_ptr(), n, H, W, 0 if self.input_u8 is None else 1,
                mean_arr, float(img_scale)), "lwp_plan_add_frontend")
            self.op_names.append("model.0-2.frontend")
            px2, px4 = n * (H // 2) * (W // 2), n * (H // 4) * (W // 4)
            in_bytes = n * 3 * H * W * (4 if self.input_u8 is None else 1)
            self.op_meta.append(dict(kind="frontend", flops=2.0 * (27 * 32 + 9 * 32 + 32 * 64) * px2 + 2.0 * 9 * 64 * px4,
                                     bytes=float(in_bytes + px4 * 64 * es_)))
        elif self.input_u8 is None:
            _lib.check(self.lib.lwp_plan_add_stem(self.handle, w_.data_ptr(), s_.data_ptr(), b_.data_ptr(),
                                                  pp[0].data_ptr(), n, H, W), "lwp_plan_add_stem")
        else:
            mean3, img_scale = self.input_u8
            mean_arr = (_lib._c_double * 3)(*[float(m) for m in mean3])
            _lib.check(self.lib.lwp_plan_add_stem_u8(self.handle, w_.data_ptr(), s_.data_ptr(), b_.data_ptr(),
                                                     pp[0].data_ptr(), n, H, W, mean_arr, float(img_scale)),
                       "lwp_plan_add_stem_u8")
        if not self.frontend_fused:
            self.op_names.append("model.0")
            self.op_meta.append(dict(kind="stem", flops=2.0 * 27 * 32 * n * (H // 2) * (W // 2),
                                     bytes=float(n * 3 * H * W * 4 + n * (H // 2) * (W // 2) * 32 * es_)))
        hh, ww, c, cur, cur_off = H // 2, W // 2, 32, 0, 0
        pair_chunks = int(os.environ.get("LWP_PAIR_CHUNKS", "0") or 0)
        pair_from = int(os.environ.get("LWP_PAIR_CHUNK_FROM", "2") or 2)
        for i, (dw, pw) in enumerate(P.backbone):
            ho, wo = (hh - 1) // dw.stride + 1, (ww - 1) // dw.stride + 1
            if self.frontend_fused and i == 0:     # model.1 is inside the front-end kernel ...
                hh, ww, c = ho, wo, pw.cout
                continue
            if self.frontend_fused and i == 1:     # ... and so is model.2's depthwise conv: pp[0] holds its output
                self._gemm("model.2.pw", pp[0], c, pw, n, ho, wo, out=pp[1], out_ld=pw.cout_pad)
                cur = 1
                hh, ww, c = ho, wo, pw.cout
                continue
            # fused: the depthwise result goes straight into the GEMM's smem A operand
            if self._fusable(dw, pw, "model.%d.dwpw" % (i + 1)) and self._dwpw("model.%d.dwpw" % (i + 1), pp[cur], dw, pw, n, hh, ww, pp[cur ^ 1],
                                                    pw.cout_pad):
                cur ^= 1
            elif self._sepconv("model.%d.sep" % (i + 1), pp[cur], dw, pw, n, hh, ww, pp[cur ^ 1], pw.cout_pad):
                cur ^= 1
            elif pair_chunks > 1 and i >= pair_from and n >= pair_chunks:
                # L2-resident sub-batches (experiment, LWP_PAIR_CHUNKS=S): the pair runs S times on n / S frames each, the
                # depthwise output of a sub-batch goes to one of two small scratch slots (re-written every other sub-batch,
                # so its dirty lines can stay in the 126 MB L2) and is read back by the 1x1 conv of the same sub-batch.
                # Input: pp[cur][cur_off:]; scratch and output both live in the other buffer.
                k = (n + pair_chunks - 1) // pair_chunks
                scratch = k * ho * wo * c
                out_off = 2 * scratch
                assert out_off + n * ho * wo * pw.cout_pad <= pp[cur ^ 1].numel()
                for sb in range(pair_chunks):
                    lo = sb * k
                    cnt = min(k, n - lo)
                    if cnt <= 0:
                        break
                    src = pp[cur][cur_off + lo * hh * ww * c:]
                    mid = pp[cur ^ 1][(sb & 1) * scratch:]
                    dst = pp[cur ^ 1][out_off + lo * ho * wo * pw.cout_pad:]
                    self._dw("model.%d.dw#%d" % (i + 1, sb), src, mid, dw, cnt, hh, ww)
                    self._gemm("model.%d.pw#%d" % (i + 1, sb), mid, c, pw, cnt, ho, wo, out=dst, out_ld=pw.cout_pad)
                cur ^= 1
                cur_off = out_off
            else:
                src = pp[cur][cur_off:]
                self._dw("model.%d.dw" % (i + 1), src, pp[cur ^ 1], dw, n, hh, ww)
                self._gemm("model.%d.pw" % (i + 1), pp[cur ^ 1], c, pw, n, ho, wo, out=pp[cur], out_ld=pw.cout_pad)
                cur_off = 0
            hh, ww, c = ho, wo, pw.cout
        assert (hh, ww) == (self.h, self.w) and c == 512
        feat = pp[cur][cur_off:]
        h, w = self.h, self.w
        px = n * h * w
        nc = P.cpm_align.cout  # 128
        A, t0, t1, i0 = (self._buf(px * nc) for _ in range(4))
        r = [self._buf(px * nc), self._buf(px * nc)]
        concat = self._buf(px * CONCAT_LD, zero=True)
        self.concat = concat
        mid0 = P.init_heads[0].cout_pad
        big = self._buf(px * mid0)
        self.heads_f32 = [self._buf(px * HEAD_LD, dtype=torch.float32, zero=True) for _ in range(n_stages_out)]

        self._gemm("cpm.align", feat, 512, P.cpm_align, n, h, w, out=A, out_ld=nc)
        src = A
        for i, (dw, pw) in enumerate(P.cpm_trunk):
            last = i == len(P.cpm_trunk) - 1
            res = A if last else None  # x + trunk(x) fused into the last epilogue
            dst = t1 if src is t0 else t0
            if self._fusable(dw, pw, "cpm.trunk.%d.dwpw" % i) and self._dwpw("cpm.trunk.%d.dwpw" % i, src, dw, pw, n, h, w, dst, nc,
                                                    residual=res, res_ld=nc):
                pass
            elif self._sepconv("cpm.trunk.%d.sep" % i, src, dw, pw, n, h, w, dst, nc, residual=res, res_ld=nc):
                pass
            else:
                mid = t1 if src is t0 else t0
                dst = t0 if mid is t1 else t1
                self._dw("cpm.trunk.%d.dw" % i, src, mid, dw, n, h, w)
                self._gemm("cpm.trunk.%d.pw" % i, mid, nc, pw, n, h, w, out=dst, out_ld=nc, residual=res, res_ld=nc)
            src = dst
        t1, t0 = (src, t1 if src is t0 else t0)  # t1 := cpm trunk output, t0 := the free scratch buffer
        self._gemm("cpm.conv", t1, nc, P.cpm_conv, n, h, w, out=concat, out_ld=CONCAT_LD)
        # initial stage
        srcs = [(concat, CONCAT_LD), (t0, nc), (t1, nc)]
        dsts = [t0, t1, t0]
        for i, g in enumerate(P.init_trunk):
            self._gemm("initial_stage.trunk.%d" % i, srcs[i][0], srcs[i][1], g, n, h, w, out=dsts[i], out_ld=nc)
        more = len(P.refine) > 0
        last_heads = ("initial_stage.heads", t0, nc, P.init_heads, n, h, w, big, concat if more else None, nc)
        i_last = len(self.op_names)
        self._heads(*last_heads, self.heads_f32[0])
        # refinement stages
        # bf16 plans: the second 3x3 of block k (+ residual) and the `initial` 1x1 of block k + 1 run as ONE kernel
        # (lwp_plan_add_conv3x3_pw: the block output never leaves tensor memory); LWP_CONV3_PW=0 keeps the two ops
        fuse_pw = self.tdtype == torch.bfloat16 and os.environ.get("LWP_CONV3_PW", "1") != "0"
        i0b = self._buf(px * nc) if (fuse_pw and len(P.refine) > 0) else None
        es = 2
        for s, (blks, heads) in enumerate(P.refine):
            src, src_ld = concat, CONCAT_LD
            ibuf = [i0, i0b]
            have_initial = False     # True when the previous block's fused kernel already produced this block's initial features
            for k, (ini, c0, c1) in enumerate(blks):
                cur_i = ibuf[k & 1] if fuse_pw else i0
                if not have_initial:
                    self._gemm("refinement_stages.%d.trunk.%d.initial" % (s, k), src, src_ld, ini, n, h, w, out=cur_i,
                               out_ld=nc)
                self._gemm("refinement_stages.%d.trunk.%d.trunk.0" % (s, k), cur_i, nc, c0, n, h, w, out=t0, out_ld=nc)
                have_initial = False
                if fuse_pw and k + 1 < len(blks):
                    nxt = blks[k + 1][0]
                    nxt_i = ibuf[(k + 1) & 1]
                    rc = self.lib.lwp_plan_add_conv3x3_pw(
                        self.handle, t0.data_ptr(), nc, c1.w.data_ptr(), c1.scale.data_ptr(), c1.shift.data_ptr(),
                        cur_i.data_ptr(), nc, c1.act, nxt.w.data_ptr(), nxt.scale.data_ptr(), nxt.shift.data_ptr(), nxt.act,
                        nxt_i.data_ptr(), nc, n, h, w, c1.cin, c1.dilation) if (
                            c1.taps == 9 and c1.cout == 128 and nxt.taps == 1 and nxt.cin == 128 and nxt.cout == 128) else 3
                    if rc != 3:   # LWP_ECAP: two ops
                        _lib.check(rc, "lwp_plan_add_conv3x3_pw")
                        self.op_names.append("refinement_stages.%d.trunk.%d.trunk.1+%d.initial" % (s, k, k + 1))
                        pxs = n * h * w
                        self.op_meta.append(dict(kind="gemm3x3", flops=pxs * (2.0 * 9 * c1.cin * c1.cout + 2.0 * nxt.cin * nxt.cout),
                                                 bytes=float(pxs * (c1.cin + 2 * c1.cout) * es + (9 * c1.cin * c1.cout + nxt.cin * nxt.cout) * es)))
                        have_initial = True
                        continue
                dst = r[k & 1]
                self._gemm("refinement_stages.%d.trunk.%d.trunk.1" % (s, k), t0, nc, c1, n, h, w, out=dst, out_ld=nc,
                           residual=cur_i, res_ld=nc)  # initial_features + trunk_features
                src, src_ld = dst, nc
            more = s + 1 < len(P.refine)
            last_heads = ("refinement_stages.%d.heads" % s, src, nc, heads, n, h, w, big, concat if more else None, nc)
            i_last = len(self.op_names)
            self._heads(*last_heads, self.heads_f32[s + 1])
        self.num_compute_ops = len(self.op_names)
        self._last_heads, self._last_heads_span = last_heads, (i_last, self.num_compute_ops)
        # NCHW float32 tensors handed back by forward()
        self.outputs = []
        for s in range(n_stages_out):
            hm = self._buf(n, num_heatmaps, h, w, dtype=torch.float32)
            paf = self._buf(n, num_pafs, h, w, dtype=torch.float32)
            for t, c0_, cc in ((hm, 0, num_heatmaps), (paf, num_heatmaps, num_pafs)):
                _lib.check(self.lib.lwp_plan_add_nhwc_to_nchw(self.handle, self.heads_f32[s].data_ptr(), HEAD_LD, 1,
                                                              c0_, cc, t.data_ptr(), n, h, w),
                           "lwp_plan_add_nhwc_to_nchw")
                self.op_names.append("to_nchw.%d" % s)
                self.op_meta.append(dict(kind="layout", flops=0.0, bytes=float(2 * n * cc * h * w * 4)))
            self.outputs += [hm, paf]
        self._alt_span = None
        if self.double_heads:
            self.heads_alt = self._buf(px * HEAD_LD, dtype=torch.float32, zero=True)
            a0 = len(self.op_names)
            self._heads(*self._last_heads, self.heads_alt)
            self._alt_span = (a0, len(self.op_names))

    # -- export -------------------------------------------------------------------------------
    def export_blob(self):
        """Serialise this plan -- the op list with every argument, the pre-folded / pre-packed constants (data) and the
        activation buffers (sizes only) -- into the self-contained blob lwp_net_load() (include/lwpose_b200.h) rebuilds
        a runnable network from, without Python.  Format: csrc/net_blob.cu."""
        import ctypes
        import struct
        torch.cuda.synchronize(self.device)
        consts = []
        _all_tensors(self.packed, consts, set())
        consts = [t for t in consts if t.is_cuda]
        tensors = [(t, 1 if z else 0) for t, z in zip(self.bufs, self.buf_zero)] + [(t, 2) for t in consts]
        spans = sorted((t.data_ptr(), t.data_ptr() + t.numel() * t.element_size(), i) for i, (t, _) in enumerate(tensors)
                       if t.numel() > 0)

        def resolve(ptr):
            for lo, hi, i in spans:
                if lo <= ptr < hi:
                    return i, ptr - lo
            raise _lib.LwpError("export_blob: a plan argument points outside every known tensor")

        out = [struct.pack("<4sI", b"LWPB", 1),
               struct.pack("<7i", _PREC[self.precision][0], self.n, self.H, self.W, len(tensors), len(self.call_log),
                           self.num_compute_ops)]
        for t, kind in tensors:
            nbytes = t.numel() * t.element_size()
            out.append(struct.pack("<Qi", nbytes, kind))
            if kind == 2:
                raw = t.detach().contiguous().view(torch.uint8).cpu().numpy().tobytes()
                out.append(raw + b"\0" * (-len(raw) % 16))
        for name, args in self.call_log:
            argtypes = _lib.SIGNATURES[name][1][1:]
            out.append(struct.pack("<2i", _BLOB_FUNCS[name], len(args)))
            for a, ty in zip(args, argtypes):
                if ty is ctypes.c_void_p:
                    if a is None:
                        out.append(struct.pack("<i", 3))
                    else:
                        tid, off = resolve(int(a))
                        out.append(struct.pack("<iiQ", 2, tid, off))
                elif ty is ctypes.c_double:
                    out.append(struct.pack("<id", 1, float(a)))
                elif ty is ctypes.c_int:
                    out.append(struct.pack("<iq", 0, int(a)))
                else:   # POINTER(c_double): the three channel means of the uint8 stem
                    out.append(struct.pack("<i3d", 4, *[float(a[i]) for i in range(3)]))
        heads = [resolve(t.data_ptr())[0] for t in self.heads_f32]
        nchw = [resolve(t.data_ptr())[0] for t in self.outputs]
        out.append(struct.pack("<2i", len(heads), len(nchw)))
        out.append(struct.pack("<%di" % (len(heads) + len(nchw)), *(heads + nchw)))
        return b"".join(out)

    # -- execution ----------------------------------------------------------------------------
    def run(self, x, first=0, last=None):
        """Enqueue ops [first, last) on the current stream.  x: contiguous cuda tensor, float32 [n,3,H,W] or (plans
        built with input_u8) uint8 [n,H,W,3]."""
        if self.input_u8 is None:
            assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape) == (self.n, 3, self.H, self.W)
        else:
            assert x.is_cuda and x.dtype == torch.uint8 and x.is_contiguous() and tuple(x.shape) == (self.n, self.H, self.W, 3)
        last = len(self.op_names) if last is None else last
        _lib.check(self.lib.lwp_plan_run_range(self.handle, x.data_ptr(), first, last, _lib.current_stream()),
                   "lwp_plan_run")

    def run_compute(self, x, alt=0):
        """All layers, without the NCHW hand-off copies (the fused pipeline reads heads_f32 directly).  alt=1 (plans built
        with double_heads): the last stage's heads go to self.heads_alt instead of self.heads_f32[-1]."""
        if alt and self._alt_span is not None:
            self.run(x, 0, self._last_heads_span[0])
            self.run(x, self._alt_span[0], self._alt_span[1])
        else:
            self.run(x, 0, self.num_compute_ops)

    def error_flag(self):
        return self.lib.lwp_plan_error_flag(self.handle)

    @property
    def num_launches(self):
        return len(self.op_names)


class NetEngine:
    MAX_CACHED_PLANS = 12   # LRU bound: val.infer over variable-width images creates one plan per (scale, shape)

    def __init__(self, net):
        _lib.require_cuda()
        self.net = net
        p = next(net.parameters())
        if not p.is_cuda:
            raise _lib.LwpError("module parameters are on the CPU; call net.cuda() (there is no CPU path)")
        # the buffer plan hard-codes the reference's default widths (CONCAT_LD = 128 + 64, HEAD_LD = 64): refuse
        # anything else loudly instead of writing heads into a neighbouring pixel's row
        if net.num_channels != 128 or net.num_heatmaps + net.num_pafs > HEAD_LD:
            raise ValueError("the sm_100a engine supports num_channels=128 and num_heatmaps+num_pafs<=%d "
                             "(got num_channels=%d, %d+%d)" % (HEAD_LD, net.num_channels, net.num_heatmaps, net.num_pafs))
        self.device = p.device
        self._packed = {}
        self._plans = collections.OrderedDict()

    def new_plan(self, precision, n, H, W, input_u8=None, double_heads=False):
        """A private, uncached plan (own activation buffers): every PosePipeline chunk takes one, so two pipelines
        (or a pipeline and net.forward) of the same shape never share buffers across streams; it lives as long as
        its owner."""
        net = self.net
        with torch.cuda.device(self.device):
            return Plan(self.packed(precision), precision, n, H, W, 1 + len(net.refinement_stages),
                        net.num_heatmaps, net.num_pafs, self.device, input_u8=input_u8, double_heads=double_heads)

    def packed(self, precision):
        if precision not in _PREC:
            raise ValueError("precision must be 'bf16' or 'tf32'")
        if precision not in self._packed:
            with torch.no_grad():
                self._packed[precision] = _Packed(self.net, _PREC[precision][1])
        return self._packed[precision]

    def plan(self, precision, n, H, W, slot=0, input_u8=None):
        """Cached launch plan (with its own activation buffers) for one shape, shared by net.forward / infer_fast /
        val.infer (synchronous, one stream); input_u8 = (mean3, scale) makes the stem take raw uint8 [n,H,W,3]
        frames.  At most MAX_CACHED_PLANS shapes stay alive (least recently used goes first)."""
        key = (precision, n, H, W, slot, None if input_u8 is None else (tuple(input_u8[0]), float(input_u8[1])))
        if key in self._plans:
            self._plans.move_to_end(key)
            return self._plans[key]
        net = self.net
        with torch.cuda.device(self.device):
            while len(self._plans) >= self.MAX_CACHED_PLANS:
                # least recently used plan: its kernels may still be in flight, so drain the device before its
                # buffers and tensor maps go away (a pipeline that still holds the plan keeps it alive regardless)
                torch.cuda.synchronize(self.device)
                self._plans.popitem(last=False)
        plan = self.new_plan(precision, n, H, W, input_u8=input_u8)
        self._plans[key] = plan
        return plan

    def forward(self, x, precision="tf32"):
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected input [N,3,H,W]")
        x = x.contiguous().float()
        n, _, H, W = x.shape
        plan = self.plan(precision, n, H, W)
        with torch.cuda.device(self.device):
            plan.run(x)
        return [t.clone() for t in plan.outputs]
