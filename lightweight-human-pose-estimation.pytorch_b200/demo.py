"""Drop-in mirror of `infer_fast` from the reference's demo.py:54-78.  The GUI loop, readers, tracking and
drawing of demo.py are out of scope; `run_frame` shows the per-frame call pattern of demo.py:91-100 on
top of the drop-in functions."""
import numpy as np

from .val import normalize, pad_width


def infer_fast(net, img, net_input_height_size, stride, upsample_ratio, cpu,
               pad_value=(0, 0, 0), img_mean=(128, 128, 128), img_scale=1/256, gpu_preprocess=False):
    """img: uint8 BGR [H,W,3].  Returns (heatmaps [h',w',19], pafs [h',w',38], scale, pad): float32 host
    arrays at upsample_ratio x the stride-8 grid, like the reference.  `cpu=True` is refused.
    gpu_preprocess=True (not in the reference's signature): the raw frame is uploaded as uint8 and the cubic resize,
    the normalisation and the pad (demo.py:59-62) run on the GPU -- bit-exact with OpenCV's generic uint8 cubic path;
    a cv2 built with IPP (like the one the default host path calls) differs from it by +-1 in ~5 % of the pixels."""
    import cv2
    import torch
    from . import postproc
    from .val import _heads_on_device
    if cpu:
        raise RuntimeError("lwpose_b200 has no CPU path (infer_fast(cpu=True))")
    height = img.shape[0]
    scale = net_input_height_size / height
    if gpu_preprocess:
        if tuple(pad_value) != (0, 0, 0) or stride != 8:
            raise ValueError("gpu_preprocess supports the reference's defaults pad_value=(0, 0, 0), stride=8")
        from .engine import HEAD_LD
        _, _, (Hp, Wp), pad = postproc.infer_fast_geometry(height, img.shape[1], net_input_height_size, stride)
        raw = torch.from_numpy(np.ascontiguousarray(img[None])).cuda()
        x8 = postproc.resize_pad_u8(raw, fx=scale, fy=scale, padded=(Hp, Wp), top=pad[0], left=pad[1],
                                    pad_value=tuple(int(round(float(m))) for m in img_mean))
        plan = net.engine().plan(net.precision, 1, Hp, Wp, input_u8=(tuple(img_mean), float(img_scale)))
        plan.run_compute(x8)
        heads = plan.heads_f32[-1].view(1, Hp // 8, Wp // 8, HEAD_LD)
        heat = postproc.upsample_cubic(heads, channels=19, fx=upsample_ratio, fy=upsample_ratio, channel_offset=0)
        pafs = postproc.upsample_cubic(heads, channels=38, fx=upsample_ratio, fy=upsample_ratio, channel_offset=19)
        torch.cuda.synchronize()
        return heat[0].cpu().numpy(), pafs[0].cpu().numpy(), scale, pad
    scaled = cv2.resize(img, (0, 0), fx=scale, fy=scale, interpolation=cv2.INTER_CUBIC)
    scaled = normalize(scaled, img_mean, img_scale)
    padded, pad = pad_width(scaled, stride, pad_value,
                            [net_input_height_size, max(scaled.shape[1], net_input_height_size)])
    heads = _heads_on_device(net, padded)
    heat = postproc.upsample_cubic(heads, channels=19, fx=upsample_ratio, fy=upsample_ratio, channel_offset=0)
    pafs = postproc.upsample_cubic(heads, channels=38, fx=upsample_ratio, fy=upsample_ratio, channel_offset=19)
    torch.cuda.synchronize()
    return heat[0].cpu().numpy(), pafs[0].cpu().numpy(), scale, pad


def run_frame(net, img, height_size=256, stride=8, upsample_ratio=4):
    """One iteration of run_demo's loop body (demo.py:93-103): returns (pose_entries, all_keypoints) with
    key-point coordinates mapped back to the original image."""
    from .modules.keypoints import extract_keypoints, group_keypoints
    heatmaps, pafs, scale, pad = infer_fast(net, img, height_size, stride, upsample_ratio, False)
    total, by_type = 0, []
    for k in range(18):
        total += extract_keypoints(heatmaps[:, :, k], by_type, total)
    pose_entries, all_keypoints = group_keypoints(by_type, pafs, demo=True)
    for k in range(all_keypoints.shape[0]):
        all_keypoints[k, 0] = (all_keypoints[k, 0] * stride / upsample_ratio - pad[1]) / scale
        all_keypoints[k, 1] = (all_keypoints[k, 1] * stride / upsample_ratio - pad[0]) / scale
    return pose_entries, all_keypoints
