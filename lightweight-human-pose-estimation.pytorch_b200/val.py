"""Drop-in mirror of the inference helpers of the reference's val.py: `normalize` (:30-33), `pad_width`
(:36-49) and `infer` (:81-110, multi-scale inference with averaging).  The network forward, both cubic
resizes per scale and the running average run on the GPU; only the input image preparation (host cv2
resize of the frame, a "next" row of SURVEY.md section 8f) stays on the host like in the reference.
The COCO evaluation harness (`evaluate`, `run_coco_eval`, `convert_to_coco_format`) is out of scope."""
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
    avg_h = torch.zeros((height, width, 19), dtype=torch.float32, device="cuda")
    avg_p = torch.zeros((height, width, 38), dtype=torch.float32, device="cuda")
    count = torch.tensor(float(len(ratios)), dtype=torch.float32, device="cuda")
    for ratio in ratios:
        scaled = cv2.resize(normed, (0, 0), fx=ratio, fy=ratio, interpolation=cv2.INTER_CUBIC)
        padded, pad = pad_width(scaled, stride, pad_value, [base_height, max(scaled.shape[1], base_height)])
        heads = _heads_on_device(net, padded)
        for off, ch, avg in ((0, 19, avg_h), (19, 38, avg_p)):
            up = postproc.upsample_cubic(heads, channels=ch, fx=stride, fy=stride, channel_offset=off)
            crop = up[:, pad[0]:up.shape[1] - pad[2], pad[1]:up.shape[2] - pad[3], :].contiguous()
            full = postproc.upsample_cubic(crop, dsize=(width, height))
            avg.add_(torch.div(full[0], count))  # avg = avg + maps / len(scales), float32 true division
    return avg_h.cpu().numpy(), avg_p.cpu().numpy()
