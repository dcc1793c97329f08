"""Drop-in mirror of the reference's modules/keypoints.py (same names, arguments, return types and side
effects) -- the compute runs in the CUDA kernels of csrc/postproc.cu, there is no NumPy fallback.

  extract_keypoints(heatmap, all_keypoints, total_keypoint_num) -> int     (reference :16-48)
  group_keypoints(all_keypoints_by_type, pafs, pose_entry_size=20,
                  min_paf_score=0.05, demo=False) -> (pose_entries, all_keypoints)   (reference :51-201)

For batches use lwpose_b200.postproc / lwpose_b200.pipeline, which keep everything on the device.
"""
import numpy as np

# limb tables, reference modules/keypoints.py:5-8 (modules/pose.py:4 imports them from here)
BODY_PARTS_KPT_IDS = [[1, 2], [1, 5], [2, 3], [3, 4], [5, 6], [6, 7], [1, 8], [8, 9], [9, 10], [1, 11],
                      [11, 12], [12, 13], [1, 0], [0, 14], [14, 16], [0, 15], [15, 17], [2, 16], [5, 17]]
BODY_PARTS_PAF_IDS = ([12, 13], [20, 21], [14, 15], [16, 17], [22, 23], [24, 25], [0, 1], [2, 3], [4, 5],
                      [6, 7], [8, 9], [10, 11], [28, 29], [30, 31], [34, 35], [32, 33], [36, 37], [18, 19], [26, 27])

_MAX_CAP = 1 << 15


def extract_keypoints(heatmap, all_keypoints, total_keypoint_num):
    """Peaks of one heat-map channel.  `heatmap`: float32 [H, W] NumPy array (any strides, e.g. the
    view heatmaps[:, :, k]); it is thresholded in place like the reference does (:17)."""
    import torch
    from .. import _lib, postproc
    _lib.require_cuda()
    if heatmap.dtype != np.float32 or heatmap.ndim != 2:
        raise TypeError("heatmap must be a 2-D float32 array")
    H, W = heatmap.shape
    dev = torch.from_numpy(np.ascontiguousarray(heatmap)).cuda().view(1, H, W, 1)
    cap_k, cap_c = 256, 4096
    while True:
        kb = postproc.extract_keypoints_batched(dev, n_ch=1, cap_kpts=cap_k, cap_candidates=cap_c)
        kpts_h, counts_h, start_h, ovf = kb.to_host()
        if not ovf[0]:
            break
        if cap_c >= _MAX_CAP // 2 and cap_k >= cap_c:
            postproc.raise_on_overflow(ovf)
        cap_c = min(cap_c * 2, _MAX_CAP // 2)
        cap_k = min(cap_k * 4, cap_c)
    heatmap[heatmap < 0.1] = 0  # the reference's visible side effect on the caller's array
    lst = postproc.keypoint_lists(kpts_h, counts_h, start_h, 0)[0]
    all_keypoints.append([(x, y, s, total_keypoint_num + i) for i, (x, y, s, _) in enumerate(lst)])
    return len(lst)


def _upload_keypoints(all_keypoints_by_type, device):
    import torch
    from .. import postproc
    if len(all_keypoints_by_type) != postproc.NUM_KPT_TYPES:
        raise ValueError("expected 18 key-point lists")
    cap = max(1, max(len(l) for l in all_keypoints_by_type))
    kp = np.zeros((1, 18, cap, 4), np.int32)
    counts = np.zeros((1, 18), np.int32)
    start = np.zeros((1, 19), np.int32)
    nxt = 0
    for c, lst in enumerate(all_keypoints_by_type):
        counts[0, c] = len(lst)
        start[0, c] = nxt
        for j, k in enumerate(lst):
            if int(k[3]) != nxt + j:
                raise ValueError("key-point ids must be consecutive over the 18 channels (as demo.py:95-98 "
                                 "produces them)")
            kp[0, c, j, 0], kp[0, c, j, 1] = int(k[0]), int(k[1])
            kp[0, c, j, 2] = np.float32(k[2]).view(np.int32)
            kp[0, c, j, 3] = nxt + j
        nxt += len(lst)
    start[0, 18] = nxt
    kb = postproc.KeypointBatch(1, 18, cap, device)
    kb.kpts.copy_(torch.from_numpy(kp))
    kb.counts.copy_(torch.from_numpy(counts))
    kb.kpt_start.copy_(torch.from_numpy(start))
    return kb


def group_keypoints(all_keypoints_by_type, pafs, pose_entry_size=20, min_paf_score=0.05, demo=False):
    """PAF grouping of one image.  pafs: float32 [H, W, 38] NumPy array (the up-sampled PAFs)."""
    import torch
    from .. import _lib, postproc
    _lib.require_cuda()
    if pose_entry_size != postproc.POSE_ENTRY:
        raise ValueError("pose_entry_size must be 20")
    all_keypoints = np.array([item for sublist in all_keypoints_by_type for item in sublist])
    pafs_d = torch.from_numpy(np.ascontiguousarray(pafs, dtype=np.float32)).cuda().unsqueeze(0)
    kb = _upload_keypoints(all_keypoints_by_type, pafs_d.device)
    cap_p, cap_c = 256, 4096
    while True:
        kb.overflow.zero_()
        pose_d, n_d = postproc.group_keypoints_batched(kb, pafs_d, demo=demo, min_paf_score=min_paf_score,
                                                       cap_poses=cap_p, cap_connections=cap_c)
        ovf = kb.overflow.cpu().numpy()
        if not ovf[0]:
            break
        if cap_p >= _MAX_CAP:
            postproc.raise_on_overflow(ovf)
        cap_p *= 4
        cap_c = min(cap_c * 2, 8192)
    pose_entries = postproc.pose_entries_array(pose_d.cpu().numpy(), n_d.cpu().numpy(), 0)
    return pose_entries, all_keypoints
