"""Batched device API of the post-processing half of the hot path (torch tensors on cuda in, torch
tensors on cuda out; all compute in the C-ABI kernels of csrc/postproc.cu).

  upsample_cubic              cv2.resize(..., INTER_CUBIC)  (reference demo.py:72,76; val.py:98-107)
  extract_keypoints_batched   18 x extract_keypoints        (reference modules/keypoints.py:16-48)
  group_keypoints_batched     group_keypoints               (reference modules/keypoints.py:51-201)
"""
import numpy as np
import torch

from . import _lib

NUM_KPT_TYPES = 18
NUM_LIMBS = 19
POSE_ENTRY = 20


class CapacityOverflow(_lib.LwpError):
    """A fixed-capacity device buffer was too small for some image; results for it are invalid."""


def _ptr(t):
    return t.data_ptr() if t is not None else None


def upsample_cubic(src, channels=None, fx=None, fy=None, dsize=None, out=None, channel_offset=0, crop=None,
                   accumulate_divisor=None):
    """src: float32 cuda tensor [n, h, w, ld] (NHWC; channels [channel_offset, channel_offset + channels)
    of every pixel are resized).  Either fx/fy (like cv2.resize(src, (0, 0), fx=, fy=)) or dsize=(W, H).
    crop=(top, left, bottom, right): resize src[:, top:h-bottom, left:w-right] without copying it (val.py:99,106);
    accumulate_divisor=k (with out=): out += resized / k in float32 (the running average of val.py:101,108).
    Returns [n, H, W, channels] float32."""
    L = _lib.load()
    assert src.is_cuda and src.dtype == torch.float32 and src.dim() == 4 and src.is_contiguous()
    n, h, w, ld = src.shape
    c = ld - channel_offset if channels is None else int(channels)
    assert 0 <= channel_offset and channel_offset + c <= ld
    top, left, bottom, right = (0, 0, 0, 0) if crop is None else [int(v) for v in crop]
    hc, wc = h - top - bottom, w - left - right
    assert hc > 0 and wc > 0
    if dsize is None:
        inv_x, inv_y = float(fx), float(fy)
        W, H = int(np.rint(wc * inv_x)), int(np.rint(hc * inv_y))
    else:
        W, H = int(dsize[0]), int(dsize[1])
        inv_x, inv_y = W / wc, H / hc
    if out is None:
        assert accumulate_divisor is None
        out = torch.empty((n, H, W, c), dtype=torch.float32, device=src.device)
    else:
        assert out.shape == (n, H, W, c) and out.is_contiguous() and out.dtype == torch.float32
    _lib.check(L.lwp_upsample_cubic_ex(_ptr(src) + 4 * (channel_offset + (top * w + left) * ld), n, hc, wc, c, ld, w * ld, h * w * ld,
                                       _ptr(out), H, W, inv_x, inv_y, float(accumulate_divisor or 0.0), _lib.current_stream()),
               "lwp_upsample_cubic")
    return out


def resize_pad_u8(frames, fx=None, fy=None, dsize=None, padded=None, top=0, left=0, pad_value=(128, 128, 128), out=None):
    """Frame preparation on the device (reference demo.py:59,61-62): uint8 cuda [n, h, w, 3] -> cubic resize (factors fx/fy
    like cv2.resize(img, (0, 0), fx=, fy=) or dsize=(W, H)) placed at (top, left) of a padded=(Hp, Wp) frame filled with
    pad_value.  Returns uint8 [n, Hp, Wp, 3] -- what a stem built with input_format='u8_nhwc' consumes."""
    L = _lib.load()
    assert frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[3] == 3 and frames.is_contiguous()
    n, h, w, _ = frames.shape
    if dsize is None:
        inv_x, inv_y = float(fx), float(fy)
        W, H = int(np.rint(w * inv_x)), int(np.rint(h * inv_y))
    else:
        W, H = int(dsize[0]), int(dsize[1])
        inv_x, inv_y = W / w, H / h
    Hp, Wp = (H, W) if padded is None else (int(padded[0]), int(padded[1]))
    if out is None:
        out = torch.empty((n, Hp, Wp, 3), dtype=torch.uint8, device=frames.device)
    else:
        assert tuple(out.shape) == (n, Hp, Wp, 3) and out.dtype == torch.uint8 and out.is_contiguous()
    _lib.check(L.lwp_resize_pad_u8(_ptr(frames), n, h, w, _ptr(out), Hp, Wp, H, W, int(top), int(left), inv_x, inv_y,
                                   int(pad_value[0]), int(pad_value[1]), int(pad_value[2]), _lib.current_stream()),
               "lwp_resize_pad_u8")
    return out


def infer_fast_geometry(height, width, net_input_height_size, stride=8):
    """Bookkeeping of demo.infer_fast (demo.py:57-62) without touching pixels: (scale, resized (H, W), padded (Hp, Wp),
    pad [top, left, bottom, right])."""
    import math
    scale = net_input_height_size / height
    H, W = int(np.rint(height * scale)), int(np.rint(width * scale))
    hmin = math.ceil(net_input_height_size / float(stride)) * stride
    Wp = math.ceil(max(W, net_input_height_size) / float(stride)) * stride
    hh = min(net_input_height_size, H)
    top, left = int(math.floor((hmin - hh) / 2.0)), int(math.floor((Wp - W) / 2.0))
    bottom = int(hmin - hh - top)
    return scale, (H, W), (H + top + bottom, int(Wp)), [top, left, bottom, int(Wp - W - left)]


class KeypointBatch:
    """Device-side result of extract_keypoints_batched."""

    def __init__(self, n, n_ch, cap_kpts, device):
        self.n, self.n_ch, self.cap_kpts = n, n_ch, cap_kpts
        self.kpts = torch.empty((n, n_ch, cap_kpts, 4), dtype=torch.int32, device=device)  # lwp_keypoint
        self.counts = torch.empty((n, n_ch), dtype=torch.int32, device=device)
        self.kpt_start = torch.empty((n, n_ch + 1), dtype=torch.int32, device=device)
        self.overflow = torch.zeros((n,), dtype=torch.int32, device=device)

    def to_host(self):
        """(kpts int32 [n,n_ch,cap,4] with column 2 = float32 bits, counts, kpt_start, overflow) as numpy."""
        return (self.kpts.cpu().numpy(), self.counts.cpu().numpy(), self.kpt_start.cpu().numpy(),
                self.overflow.cpu().numpy())


def keypoint_lists(kpts_h, counts_h, kpt_start_h, img):
    """Rebuild the reference's all_keypoints_by_type for image `img`: per channel a list of
    (np.int64 x, np.int64 y, np.float32 score, int id) tuples (modules/keypoints.py:43-47)."""
    out = []
    n_ch = counts_h.shape[1]
    for c in range(n_ch):
        cnt = int(counts_h[img, c])
        rows = kpts_h[img, c, :cnt]
        scores = rows[:, 2].copy().view(np.float32)
        base = int(kpt_start_h[img, c])
        out.append([(np.int64(rows[j, 0]), np.int64(rows[j, 1]), scores[j], base + j) for j in range(cnt)])
    return out


def extract_keypoints_batched(hm, n_ch=NUM_KPT_TYPES, cap_kpts=128, cap_candidates=2048, workspace=None, out=None):
    """hm: float32 cuda [n, H, W, ld]; channels 0..n_ch-1 are scanned.  Returns a KeypointBatch.
    The caller must check `overflow` (see `raise_on_overflow`)."""
    L = _lib.load()
    assert hm.is_cuda and hm.dtype == torch.float32 and hm.dim() == 4 and hm.is_contiguous()
    n, H, W, ld = hm.shape
    kb = out if out is not None else KeypointBatch(n, n_ch, cap_kpts, hm.device)
    assert kb.n == n and kb.n_ch == n_ch and kb.cap_kpts == cap_kpts
    ws_bytes = L.lwp_extract_workspace_bytes(n, n_ch, cap_candidates)
    if workspace is None or workspace.numel() < ws_bytes:
        workspace = torch.empty((ws_bytes,), dtype=torch.uint8, device=hm.device)
    _lib.check(L.lwp_extract_keypoints(_ptr(hm), n, H, W, ld, n_ch, _ptr(kb.kpts), _ptr(kb.counts),
                                       _ptr(kb.kpt_start), cap_kpts, cap_candidates, _ptr(workspace),
                                       workspace.numel(), _ptr(kb.overflow), _lib.current_stream()),
               "lwp_extract_keypoints")
    kb._ws = workspace  # keep alive until the stream has consumed it
    return kb


def group_keypoints_batched(kb, pafs, demo=False, min_paf_score=0.05, cap_poses=128, cap_connections=2048,
                            workspace=None, out=None):
    """kb: KeypointBatch with 18 channels; pafs: float32 cuda [n, H, W, ld>=38] (up-sampled).
    Returns (pose_entries float64 [n, cap_poses, 20], n_poses int32 [n]); overflow is OR-ed into kb.overflow."""
    L = _lib.load()
    assert kb.n_ch == NUM_KPT_TYPES
    assert pafs.is_cuda and pafs.dtype == torch.float32 and pafs.dim() == 4 and pafs.is_contiguous()
    n, H, W, ld = pafs.shape
    assert n == kb.n
    if out is None:
        pose_entries = torch.empty((n, cap_poses, POSE_ENTRY), dtype=torch.float64, device=pafs.device)
        n_poses = torch.empty((n,), dtype=torch.int32, device=pafs.device)
    else:
        pose_entries, n_poses = out
    ws_bytes = L.lwp_group_workspace_bytes(n, kb.cap_kpts, cap_connections, cap_poses)
    if workspace is None or workspace.numel() < ws_bytes:
        workspace = torch.empty((ws_bytes,), dtype=torch.uint8, device=pafs.device)
    _lib.check(L.lwp_group_keypoints(_ptr(kb.kpts), _ptr(kb.counts), _ptr(kb.kpt_start), kb.cap_kpts, _ptr(pafs), n,
                                     H, W, ld, int(bool(demo)), float(min_paf_score), _ptr(pose_entries),
                                     _ptr(n_poses), cap_poses, cap_connections, _ptr(workspace), workspace.numel(),
                                     _ptr(kb.overflow), _lib.current_stream()),
               "lwp_group_keypoints")
    kb._gws = workspace
    return pose_entries, n_poses


def extract_keypoints_fused(maps, ratio, n_ch=NUM_KPT_TYPES, c_layout=19, channel_offset=0, cap_kpts=128,
                            cap_candidates=2048, workspace=None, out=None):
    """Peaks of the (never materialised) `ratio` x cubic up-sampling of `maps`: float32 cuda [n, h, w, ld], heat-map
    channels [channel_offset, channel_offset + c_layout).  Same result as upsample_cubic + extract_keypoints_batched."""
    L = _lib.load()
    assert maps.is_cuda and maps.dtype == torch.float32 and maps.dim() == 4 and maps.is_contiguous()
    n, h, w, ld = maps.shape
    H, W = int(np.rint(h * float(ratio))), int(np.rint(w * float(ratio)))
    kb = out if out is not None else KeypointBatch(n, n_ch, cap_kpts, maps.device)
    assert kb.n == n and kb.n_ch == n_ch and kb.cap_kpts == cap_kpts
    ws_bytes = L.lwp_extract_workspace_bytes(n, n_ch, cap_candidates)
    if workspace is None or workspace.numel() < ws_bytes:
        workspace = torch.empty((ws_bytes,), dtype=torch.uint8, device=maps.device)
    _lib.check(L.lwp_extract_keypoints_fused(_ptr(maps) + 4 * channel_offset, n, h, w, ld, n_ch, c_layout, H, W,
                                             float(ratio), float(ratio), _ptr(kb.kpts), _ptr(kb.counts),
                                             _ptr(kb.kpt_start), cap_kpts, cap_candidates, _ptr(workspace),
                                             workspace.numel(), _ptr(kb.overflow), _lib.current_stream()),
               "lwp_extract_keypoints_fused")
    kb._ws = workspace
    kb.up_size = (H, W)
    return kb


def group_keypoints_fused(kb, maps, ratio, channel_offset=19, demo=False, min_paf_score=0.05, cap_poses=128,
                          cap_connections=2048, workspace=None, out=None):
    """PAF grouping with every PAF sample computed on the fly from `maps` (float32 cuda [n, h, w, ld], the 38 PAF
    channels start at channel_offset).  Same result as upsample_cubic + group_keypoints_batched."""
    L = _lib.load()
    assert kb.n_ch == NUM_KPT_TYPES
    assert maps.is_cuda and maps.dtype == torch.float32 and maps.dim() == 4 and maps.is_contiguous()
    n, h, w, ld = maps.shape
    assert n == kb.n and channel_offset + 38 <= ld
    H, W = int(np.rint(h * float(ratio))), int(np.rint(w * float(ratio)))
    if out is None:
        pose_entries = torch.empty((n, cap_poses, POSE_ENTRY), dtype=torch.float64, device=maps.device)
        n_poses = torch.empty((n,), dtype=torch.int32, device=maps.device)
    else:
        pose_entries, n_poses = out
    ws_bytes = L.lwp_group_workspace_bytes(n, kb.cap_kpts, cap_connections, cap_poses) + L.lwp_paf_pack_bytes(n, h, w)
    if workspace is None or workspace.numel() < ws_bytes:
        workspace = torch.empty((ws_bytes,), dtype=torch.uint8, device=maps.device)
    _lib.check(L.lwp_group_keypoints_fused(_ptr(kb.kpts), _ptr(kb.counts), _ptr(kb.kpt_start), kb.cap_kpts,
                                           _ptr(maps) + 4 * channel_offset, n, h, w, ld, H, W, float(ratio),
                                           float(ratio), int(bool(demo)), float(min_paf_score), _ptr(pose_entries),
                                           _ptr(n_poses), cap_poses, cap_connections, _ptr(workspace),
                                           workspace.numel(), _ptr(kb.overflow), _lib.current_stream()),
               "lwp_group_keypoints_fused")
    kb._gws = workspace
    return pose_entries, n_poses


def pose_convert_xform(n, pad, scale, device):
    """Per-image (left pad, top pad, scale) table of pose_convert on the device.  Uploading it is a pageable, host-blocking
    copy: callers that convert every batch with the same geometry (PosePipeline) build it once."""
    pads = np.asarray(pad, np.float64).reshape(-1, 4)
    scales = np.asarray(scale, np.float64).reshape(-1)
    xf = np.empty((n, 3), np.float64)
    xf[:, 0] = pads[:, 1] if pads.shape[0] == n else pads[0, 1]
    xf[:, 1] = pads[:, 0] if pads.shape[0] == n else pads[0, 0]
    xf[:, 2] = scales if scales.shape[0] == n else scales[0]
    return torch.from_numpy(xf).to(device)


def pose_convert(pose_entries, n_poses, kb, stride=8, upsample_ratio=4, pad=(0, 0, 0, 0), scale=1.0, out=None, xform=None):
    """Batched result post-conversion on the device (reference demo.py:101-115 + modules/pose.py:30-39).
    pose_entries float64 [n, cap, 20] / n_poses int32 [n] as returned by group_keypoints_*; kb: the KeypointBatch.
    pad = [top, left, bottom, right] and scale as returned by infer_fast -- one pair for the whole batch, or per-image
    sequences.  Returns (pose_kpts int32 [n, cap, 18, 2], bbox int32 [n, cap, 4], confidence float64 [n, cap])."""
    L = _lib.load()
    n, cap = pose_entries.shape[0], pose_entries.shape[1]
    dev = pose_entries.device
    if xform is None:
        xform = pose_convert_xform(n, pad, scale, dev)
    if out is None:
        out = (torch.empty((n, cap, NUM_KPT_TYPES, 2), dtype=torch.int32, device=dev),
               torch.empty((n, cap, 4), dtype=torch.int32, device=dev),
               torch.empty((n, cap), dtype=torch.float64, device=dev))
    pk, bb, conf = out
    _lib.check(L.lwp_pose_convert(_ptr(pose_entries), _ptr(n_poses), cap, _ptr(kb.kpts), _ptr(kb.kpt_start), kb.cap_kpts, n,
                                  float(stride), float(upsample_ratio), _ptr(xform), _ptr(pk), _ptr(bb), _ptr(conf),
                                  _lib.current_stream()), "lwp_pose_convert")
    kb._xform = xform  # keep alive until the stream has consumed it
    return pk, bb, conf


def raise_on_overflow(overflow_h):
    bad = np.nonzero(np.asarray(overflow_h))[0]
    if bad.size:
        raise CapacityOverflow("capacity exceeded for image(s) %s; raise cap_kpts / cap_candidates / "
                               "cap_connections / cap_poses" % bad.tolist()[:8])


def pose_entries_array(pose_entries_h, n_poses_h, img):
    """The reference's return value: float64 [P, 20], or shape (0,) when no pose survives
    (np.asarray([]) at modules/keypoints.py:200)."""
    cnt = int(n_poses_h[img])
    return np.asarray([pose_entries_h[img, j].copy() for j in range(cnt)])
