"""Phase timing of configs[4] (val multi-scale) for one 16-frame chunk: infer_batch per scale (input resize / network /
output resizes), extract + group at the frame size, host conversion."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import postproc, synth, val
    from lwpose_b200.engine import HEAD_LD
    net = bench.make_net().cuda()
    net.precision = "bf16"
    B, Hf, Wf = 16, 480, 640
    frames = torch.from_numpy(synth.synthetic_frames(B, Hf, Wf, seed=300)).pin_memory()
    scales = [0.5, 1.0, 1.5, 2.0]

    def sync():
        torch.cuda.synchronize()
        return time.perf_counter()
    for rep in range(3):
        t0 = sync()
        res = val.evaluate_batch(net, frames, scales=scales, base_height=368)
        t1 = sync()
        avg_h, avg_p = val.infer_batch(net, frames, scales, 368, 8)
        t2 = sync()
        kb = postproc.extract_keypoints_batched(avg_h, cap_kpts=128)
        t3 = sync()
        poses_d, n_d = postproc.group_keypoints_batched(kb, avg_p, demo=False, cap_poses=256)
        t4 = sync()
        kp, cnt, st, ovf = kb.to_host()
        ph, nh = poses_d.cpu().numpy(), n_d.cpu().numpy()
        out = []
        for b in range(B):
            by_type = postproc.keypoint_lists(kp, cnt, st, b)
            allk = np.array([item for sub in by_type for item in sub])
            out.append(val.convert_to_coco_format(postproc.pose_entries_array(ph, nh, b), allk))
        t5 = sync()
        print("evaluate_batch %.2f ms | infer_batch %.2f | extract %.2f | group %.2f | host %.2f" %
              ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t5 - t4) * 1e3))
    # inside infer_batch: per scale
    eng = net.engine()
    x8 = frames.cuda()
    normed = ((x8.float() - 128.0) / 256.0).contiguous()
    for ratio, (hs, ws), (H, W), pad in val.scale_geometry(Hf, Wf, scales, 368, 8):
        t0 = sync()
        scaled = postproc.upsample_cubic(normed, channels=3, fx=ratio, fy=ratio)
        x = torch.zeros((B, 3, H, W), dtype=torch.float32, device="cuda")
        x[:, :, pad[0]:pad[0] + hs, pad[1]:pad[1] + ws] = scaled.permute(0, 3, 1, 2)
        t1 = sync()
        plan = eng.plan(net.precision, B, H, W)
        plan.run_compute(x)
        t2 = sync()
        heads = plan.heads_f32[-1].view(B, H // 8, W // 8, HEAD_LD)
        avg_h = torch.zeros((B, Hf, Wf, 19), dtype=torch.float32, device="cuda")
        avg_p = torch.zeros((B, Hf, Wf, 38), dtype=torch.float32, device="cuda")
        t2 = sync()
        ups = []
        for off, ch, avg in ((0, 19, avg_h), (19, 38, avg_p)):
            ta = sync()
            up = postproc.upsample_cubic(heads, channels=ch, fx=8, fy=8, channel_offset=off)
            tb = sync()
            postproc.upsample_cubic(up, dsize=(Wf, Hf), crop=pad, out=avg, accumulate_divisor=4)
            tc = sync()
            ups.append(((tb - ta) * 1e3, (tc - tb) * 1e3))
        print("scale %dx%d: input %.2f ms, network %.2f, x8 + resize: heat %.2f + %.2f, paf %.2f + %.2f" %
              (H, W, (t1 - t0) * 1e3, (t2 - t1) * 1e3, ups[0][0], ups[0][1], ups[1][0], ups[1][1]))


if __name__ == "__main__":
    main()
