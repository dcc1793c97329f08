"""ctypes front-end of oracle/lwp_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Function signatures and return types mirror the reference so parity tests read
like the reference's call sites:
  resize_cubic       cv2.resize(src, (0,0), fx=, fy=, INTER_CUBIC) / cv2.resize(src, (W,H), INTER_CUBIC)
                     (demo.py:72,76; val.py:98-107)
  extract_keypoints  modules/keypoints.py:16-48
  group_keypoints    modules/keypoints.py:51-201
"""
import ctypes

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(_build.build())
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int)
        dp = ctypes.POINTER(ctypes.c_double)
        L.orc_resize_cubic.restype = ctypes.c_int
        L.orc_resize_cubic.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_double, ctypes.c_double]
        L.orc_extract_keypoints.restype = ctypes.c_int
        L.orc_extract_keypoints.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.c_long,
                                            ip, ip, fp, ctypes.c_int]
        L.orc_group_keypoints.restype = ctypes.c_int
        L.orc_group_keypoints.argtypes = [ip, ip, ip, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_double, dp, ctypes.c_int]
        L.orc_resize_pad_u8.restype = ctypes.c_int
        L.orc_resize_pad_u8.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p] + [ctypes.c_int] * 6 + [
            ctypes.c_double, ctypes.c_double, ip]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def resize_cubic(src, fx=None, fy=None, dsize=None):
    """float32 HWC cubic resize with OpenCV's bits.  Either fx/fy or dsize=(W, H)."""
    src = np.ascontiguousarray(src, dtype=np.float32)
    squeeze = src.ndim == 2
    if squeeze:
        src = src[:, :, None]
    h, w, C = src.shape
    if dsize is None:
        inv_x, inv_y = float(fx), float(fy)
        W, H = int(np.rint(w * inv_x)), int(np.rint(h * inv_y))  # saturate_cast<int>(ssize*inv_scale)
    else:
        W, H = int(dsize[0]), int(dsize[1])
        inv_x, inv_y = W / w, H / h  # (double)dsize.width / ssize.width
    dst = np.empty((H, W, C), np.float32)
    rc = lib().orc_resize_cubic(_fp(src), h, w, C, _fp(dst), H, W, inv_x, inv_y)
    if rc != 0:
        raise MemoryError("orc_resize_cubic")
    return dst[:, :, 0] if squeeze else dst


def extract_keypoints(heatmap, all_keypoints, total_keypoint_num, cap=1 << 16):
    """Same contract as the reference: thresholds `heatmap` in place, appends the list of
    (x: np.int64, y: np.int64, score: np.float32, id: int) tuples, returns the count."""
    assert heatmap.dtype == np.float32 and heatmap.ndim == 2
    H, W = heatmap.shape
    rs, cs = heatmap.strides[0] // 4, heatmap.strides[1] // 4
    ox = np.empty(cap, np.int32)
    oy = np.empty(cap, np.int32)
    osc = np.empty(cap, np.float32)
    n = lib().orc_extract_keypoints(_fp(heatmap), H, W, rs, cs, _ip(ox), _ip(oy), _fp(osc), cap)
    if n < 0 or n > cap:
        raise RuntimeError("orc_extract_keypoints overflow/alloc failure: %d" % n)
    all_keypoints.append([(np.int64(ox[i]), np.int64(oy[i]), osc[i], total_keypoint_num + i) for i in range(n)])
    return n


def flatten_keypoints(all_keypoints_by_type):
    start = np.zeros(len(all_keypoints_by_type) + 1, np.int32)
    xs, ys, ss = [], [], []
    for t, lst in enumerate(all_keypoints_by_type):
        start[t + 1] = start[t] + len(lst)
        for k in lst:
            xs.append(int(k[0])); ys.append(int(k[1])); ss.append(np.float32(k[2]))
    return (start, np.asarray(xs, np.int32).reshape(-1), np.asarray(ys, np.int32).reshape(-1),
            np.asarray(ss, np.float32).reshape(-1))


def group_keypoints(all_keypoints_by_type, pafs, pose_entry_size=20, min_paf_score=0.05, demo=False,
                    cap_poses=4096):
    assert pose_entry_size == 20
    pafs = np.ascontiguousarray(pafs, dtype=np.float32)
    H, W, C = pafs.shape
    assert C == 38
    all_keypoints = np.array([item for sublist in all_keypoints_by_type for item in sublist])
    start, kx, ky, ks = flatten_keypoints(all_keypoints_by_type)
    out = np.empty((cap_poses, 20), np.float64)
    n = lib().orc_group_keypoints(_ip(start), _ip(kx), _ip(ky), _fp(ks), _fp(pafs), H, W, int(bool(demo)),
                                  float(min_paf_score), out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                  cap_poses)
    if n < 0 or n > cap_poses:
        raise RuntimeError("orc_group_keypoints overflow/alloc failure: %d" % n)
    pose_entries = np.asarray([out[i].copy() for i in range(n)])  # shape (0,) when empty, like the reference
    return pose_entries, all_keypoints


def pose_convert(pose_entries, all_keypoints, stride, upsample_ratio, pad, scale):
    """Plain-Python restatement of the reference's result post-conversion (demo.py:101-115) and of Pose.get_bbox
    (modules/pose.py:30-39; cv2.boundingRect of integer points = min corner and max - min + 1 extents).
    Returns (pose_keypoints int32 [P, 18, 2], bbox int32 [P, 4], confidence float64 [P]); all_keypoints is NOT modified."""
    allk = np.array(all_keypoints, dtype=np.float64).reshape(-1, 4).copy()
    for k in range(allk.shape[0]):
        allk[k, 0] = (allk[k, 0] * stride / upsample_ratio - pad[1]) / scale
        allk[k, 1] = (allk[k, 1] * stride / upsample_ratio - pad[0]) / scale
    poses = np.asarray(pose_entries, np.float64).reshape(-1, 20)
    kp = -np.ones((poses.shape[0], 18, 2), np.int32)
    bbox = np.zeros((poses.shape[0], 4), np.int32)
    for n in range(poses.shape[0]):
        for k in range(18):
            if poses[n, k] != -1.0:
                kp[n, k, 0] = int(allk[int(poses[n, k]), 0])
                kp[n, k, 1] = int(allk[int(poses[n, k]), 1])
        found = kp[n][kp[n][:, 0] != -1]
        if len(found):
            x0, y0 = found[:, 0].min(), found[:, 1].min()
            bbox[n] = (x0, y0, found[:, 0].max() - x0 + 1, found[:, 1].max() - y0 + 1)
    return kp, bbox, poses[:, 18].copy()


def resize_pad_u8(img, fx=None, fy=None, dsize=None, padded=None, top=0, left=0, pad_value=(128, 128, 128)):
    """uint8 [h, w, 3] cubic resize with the bits of OpenCV's generic fixed-point path (cv2.resize(img, (0, 0), fx=, fy=,
    INTER_CUBIC) of demo.py:59 with IPP off), optionally placed at (top, left) of a padded=(Hp, Wp) frame filled with
    pad_value (val.py:36-49)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w, _ = img.shape
    if dsize is None:
        inv_x, inv_y = float(fx), float(fy)
        W, H = int(np.rint(w * inv_x)), int(np.rint(h * inv_y))
    else:
        W, H = int(dsize[0]), int(dsize[1])
        inv_x, inv_y = W / w, H / h
    Hp, Wp = (H, W) if padded is None else padded
    dst = np.empty((Hp, Wp, 3), np.uint8)
    pad3 = np.asarray(pad_value, np.int32)
    rc = lib().orc_resize_pad_u8(img.ctypes.data, h, w, dst.ctypes.data, Hp, Wp, H, W, int(top), int(left), inv_x, inv_y, _ip(pad3))
    if rc != 0:
        raise MemoryError("orc_resize_pad_u8")
    return dst
