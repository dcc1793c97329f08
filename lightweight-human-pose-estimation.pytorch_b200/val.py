"""Drop-in mirror of the inference helpers of the reference's val.py: `normalize` (:30-33), `pad_width`
(:36-49) and `infer` (:81-110, multi-scale inference with averaging).  The network forward, both cubic
resizes per scale and the running average run on the GPU; only the input image preparation (host cv2
resize of the frame, a "next" row of SURVEY.md section 8f) stays on the host like in the reference.
`infer_batch` is the batched device form of `infer` for BASELINE.json configs[4] (frames of one size sharded over the
GPUs): here the per-scale input resize runs on the GPU as well.  `convert_to_coco_format` (:52-78) is mirrored;
the COCO dataset / pycocotools harness (`evaluate`, `run_coco_eval`) is out of scope."""
import math

import numpy as np


def normalize(img, img_mean, img_scale):
    """(img - mean) * scale; float32 input promoted by the mean tuple exactly like the reference."""
    arr = np.array(img, dtype=np.float32)
    return (arr - img_mean) * img_scale


def pad_width(img, stride, pad_value, min_dims):
    """Centre-pad to min_dims rounded up to a stride multiple; returns (padded, [top, left, bottom, right]).
    Mutates min_dims in place like the reference does."""
    import cv2
    h, w = img.shape[0], img.shape[1]
    h = min(min_dims[0], h)
    min_dims[0] = math.ceil(min_dims[0] / float(stride)) * stride
    min_dims[1] = math.ceil(max(min_dims[1], w) / float(stride)) * stride
    top = int(math.floor((min_dims[0] - h) / 2.0))
    left = int(math.floor((min_dims[1] - w) / 2.0))
    pad = [top, left, int(min_dims[0] - h - top), int(min_dims[1] - w - left)]
    padded = cv2.copyMakeBorder(img, pad[0], pad[2], pad[1], pad[3], cv2.BORDER_CONSTANT, value=pad_value)
    return padded, pad


def _heads_on_device(net, padded_img):
    """Run the CUDA engine on one padded HWC image; returns the last stage's float32 heads [1,h,w,64]."""
    import torch
    from .engine import HEAD_LD
    x = torch.from_numpy(padded_img).permute(2, 0, 1).unsqueeze(0).float().cuda().contiguous()
    eng = net.engine()
    n, _, H, W = x.shape
    plan = eng.plan(net.precision, n, H, W)
    plan.run_compute(x)
    return plan.heads_f32[-1].view(n, H // 8, W // 8, HEAD_LD)


def infer(net, img, scales, base_height, stride, pad_value=(0, 0, 0), img_mean=(128, 128, 128), img_scale=1/256):
    """Multi-scale inference: returns (avg_heatmaps [h,w,19], avg_pafs [h,w,38]) float32 host arrays at the
    ORIGINAL image size, bit-identical to the reference's cv2 arithmetic given identical network outputs."""
    import cv2
    import torch
    from . import postproc
    normed = normalize(img, img_mean, img_scale)
    height, width = normed.shape[0], normed.shape[1]
    ratios = [s * base_height / float(height) for s in scales]
    avg_h = torch.zeros((1, height, width, 19), dtype=torch.float32, device="cuda")
    avg_p = torch.zeros((1, height, width, 38), dtype=torch.float32, device="cuda")
    for ratio in ratios:
        scaled = cv2.resize(normed, (0, 0), fx=ratio, fy=ratio, interpolation=cv2.INTER_CUBIC)
        padded, pad = pad_width(scaled, stride, pad_value, [base_height, max(scaled.shape[1], base_height)])
        heads = _heads_on_device(net, padded)
        for off, ch, avg in ((0, 19, avg_h), (19, 38, avg_p)):
            up = postproc.upsample_cubic(heads, channels=ch, fx=stride, fy=stride, channel_offset=off)
            # crop (a view), resize to the image size and avg = avg + maps / len(scales) (float32 true division) in one kernel
            postproc.upsample_cubic(up, dsize=(width, height), crop=pad, out=avg, accumulate_divisor=len(ratios))
    return avg_h[0].cpu().numpy(), avg_p[0].cpu().numpy()


def convert_to_coco_format(pose_entries, all_keypoints):
    """Reference val.py:52-78: 17 COCO key-points (no neck) as [x + 0.5, y + 0.5, visibility] triples per pose and the
    pose score `pose score * max(0, count - 1)`; same return types (lists of Python lists / NumPy scalars)."""
    to_coco = (0, -1, 6, 8, 10, 5, 7, 9, 12, 14, 16, 11, 13, 15, 2, 1, 4, 3)
    coco_keypoints, scores = [], []
    for entry in pose_entries:
        if len(entry) == 0:
            continue
        triples = [0] * (17 * 3)
        for position, keypoint_id in enumerate(entry[:-2]):
            if position == 1:      # COCO has no neck
                continue
            cx = cy = visibility = 0
            if keypoint_id != -1:
                cx, cy = all_keypoints[int(keypoint_id), 0] + 0.5, all_keypoints[int(keypoint_id), 1] + 0.5
                visibility = 1
            triples[to_coco[position] * 3:to_coco[position] * 3 + 3] = [cx, cy, visibility]
        coco_keypoints.append(triples)
        scores.append(entry[-2] * max(0, (entry[-1] - 1)))
    return coco_keypoints, scores


def scale_geometry(height, width, scales, base_height, stride):
    """Per scale: (ratio, scaled (h, w) as cv2.resize rounds them, padded net input (H, W), pad [top, left, bottom, right])
    -- the bookkeeping of reference val.py:84-91 without touching pixels."""
    out = []
    for s in scales:
        ratio = s * base_height / float(height)
        hs, ws = int(np.rint(height * ratio)), int(np.rint(width * ratio))
        Hmin = int(math.ceil(base_height / float(stride)) * stride)
        W = int(math.ceil(max(ws, base_height) / float(stride)) * stride)
        hh = min(base_height, hs)   # pad_width pads for min(min_dims[0], h) rows: an image taller than base_height is not cropped
        top, left = int(math.floor((Hmin - hh) / 2.0)), int(math.floor((W - ws) / 2.0))
        bottom = int(Hmin - hh - top)
        out.append((ratio, (hs, ws), (hs + top + bottom, W), [top, left, bottom, int(W - ws - left)]))
    return out


def infer_batch(net, frames, scales, base_height, stride, img_mean=(128, 128, 128), img_scale=1/256):
    """Batched device form of `infer` (reference val.py:81-110) for frames of ONE size: uint8 BGR [B, h, w, 3] (host or cuda
    tensor / ndarray) -> (avg_heatmaps [B, h, w, 19], avg_pafs [B, h, w, 38]) float32 CUDA tensors at the original size.
    Everything runs on the GPU, the per-scale cubic resize of the normalised frame included (float32 arithmetic where the
    reference's cv2 call works in float64: network inputs differ by ~1e-7, far below the network tolerance; the output
    side -- x8 cubic, crop, cubic resize to (w, h), running average -- is the bit-exact OpenCV restatement)."""
    import torch
    from . import postproc
    from .engine import HEAD_LD
    x8 = torch.as_tensor(frames)
    if x8.dtype != torch.uint8 or x8.dim() != 4 or x8.shape[3] != 3:
        raise ValueError("frames must be uint8 [B, h, w, 3]")
    dev = net.engine().device
    x8 = x8.to(dev, non_blocking=True)
    B, height, width = x8.shape[0], x8.shape[1], x8.shape[2]
    mean = torch.tensor([float(m) for m in img_mean], dtype=torch.float32, device=dev)
    normed = ((x8.float() - mean) * float(img_scale)).contiguous()          # exact in float32 for uint8 pixels
    avg_h = torch.zeros((B, height, width, 19), dtype=torch.float32, device=dev)
    avg_p = torch.zeros((B, height, width, 38), dtype=torch.float32, device=dev)
    eng = net.engine()
    for ratio, (hs, ws), (H, W), pad in scale_geometry(height, width, scales, base_height, stride):
        scaled = postproc.upsample_cubic(normed, channels=3, fx=ratio, fy=ratio)
        assert tuple(scaled.shape[1:3]) == (hs, ws), (scaled.shape, hs, ws)
        if H % 8 or W % 8:
            raise ValueError("scaled frame %dx%d is not a multiple of 8 (the engine's input constraint)" % (H, W))
        x = torch.zeros((B, 3, H, W), dtype=torch.float32, device=dev)
        x[:, :, pad[0]:pad[0] + hs, pad[1]:pad[1] + ws] = scaled.permute(0, 3, 1, 2)
        plan = eng.plan(net.precision, B, H, W)
        plan.run_compute(x)
        heads = plan.heads_f32[-1].view(B, H // 8, W // 8, HEAD_LD)
        for off, ch, avg in ((0, 19, avg_h), (19, 38, avg_p)):
            up = postproc.upsample_cubic(heads, channels=ch, fx=stride, fy=stride, channel_offset=off)
            # crop (a view: no copy), resize to the frame size and avg += resized / len(scales) in ONE kernel
            postproc.upsample_cubic(up, dsize=(width, height), crop=pad, out=avg, accumulate_divisor=len(scales))
            del up
    return avg_h, avg_p


def evaluate_batch(net, frames, scales=(1,), base_height=368, stride=8, cap_kpts=128, cap_poses=256, maps_hook=None):
    """The per-image body of reference val.evaluate (:123-136) for a batch of equally sized frames, on the device:
    multi-scale inference, 18 x extract_keypoints on the averaged heat-maps, group_keypoints(demo=False), then
    convert_to_coco_format on the host tables.  Returns a list (one per frame) of (coco_keypoints, scores)."""
    import torch
    from . import postproc
    avg_h, avg_p = infer_batch(net, frames, list(scales), base_height, stride)
    if maps_hook is not None:   # callable(avg_heatmaps, avg_pafs) between inference and post-processing (benchmarks: person maps)
        maps_hook(avg_h, avg_p)
    kb = postproc.extract_keypoints_batched(avg_h, cap_kpts=cap_kpts)
    poses_d, n_d = postproc.group_keypoints_batched(kb, avg_p, demo=False, cap_poses=cap_poses)
    kp, cnt, st, ovf = kb.to_host()
    postproc.raise_on_overflow(ovf)
    poses_h, n_h = poses_d.cpu().numpy(), n_d.cpu().numpy()
    out = []
    for b in range(avg_h.shape[0]):
        by_type = postproc.keypoint_lists(kp, cnt, st, b)
        all_keypoints = np.array([item for sub in by_type for item in sub])
        out.append(convert_to_coco_format(postproc.pose_entries_array(poses_h, n_h, b), all_keypoints))
    return out
