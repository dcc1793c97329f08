"""Generate the golden fixtures under tests/golden/ by running the REAL reference (read-only at
/root/reference) and cv2 in the build container.  The GPU box has no reference, so the fixtures are
committed; re-run this script only when a generator in lwpose_b200.synth changes.

    python tests/golden/make_golden.py

Fixtures (all small):
  resize_golden.npz    cv2.resize INTER_CUBIC outputs (arrays for small cases, sha256 for large ones)
  postproc_golden.npz  extract_keypoints x18 + group_keypoints (demo True/False) on synthetic maps
  net_golden.npz       PoseEstimationWithMobileNet forward (CPU fp32) on seeded weights / inputs
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refimport  # noqa: E402
import lwpose_b200  # noqa: E402,F401
from lwpose_b200 import synth  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# C in {1, 3, 4} is excluded: OpenCV routes those through IPP (different last bits); the hot path
# only resizes 19- and 38-channel maps, which take the generic path restated by the oracle.
# (h, w, C, mode, arg): mode 'f' -> fx=fy=arg ; mode 'd' -> dsize=(W, H)=arg
RESIZE_CASES = [
    (12, 17, 19, "f", 4), (9, 11, 38, "f", 4), (7, 13, 6, "f", 8), (10, 9, 19, "d", (23, 31)),
    (16, 12, 38, "d", (37, 29)), (5, 6, 19, "d", (11, 7)), (8, 8, 2, "f", 4), (6, 10, 2, "d", (21, 13)),
    (46, 82, 19, "f", 4), (46, 82, 38, "f", 4), (32, 57, 19, "f", 4), (46, 62, 38, "f", 8),
    (368, 496, 19, "d", (640, 480)), (20, 30, 5, "d", (27, 17)), (20, 30, 7, "f", 1.5),
]


def resize_input(idx, h, w, C):
    rng = np.random.default_rng(1000 + idx)
    return rng.standard_normal((h, w, C)).astype(np.float32)


def gen_resize(out):
    import cv2
    for idx, (h, w, C, mode, arg) in enumerate(RESIZE_CASES):
        src = resize_input(idx, h, w, C)
        if mode == "f":
            dst = cv2.resize(src, (0, 0), fx=arg, fy=arg, interpolation=cv2.INTER_CUBIC)
        else:
            dst = cv2.resize(src, tuple(arg), interpolation=cv2.INTER_CUBIC)
        if dst.ndim == 2:
            dst = dst[:, :, None]
        out["resize_%d_sha" % idx] = np.array(sha(dst))
        out["resize_%d_shape" % idx] = np.array(dst.shape)
        if dst.size <= 20000:
            out["resize_%d_out" % idx] = dst


# (name, h, w, persons, noise, seed, zero_channels)
POSTPROC_CASES = [
    ("p1", 46, 82, 1, 0.0, 11, ()), ("p3", 46, 82, 3, 0.0, 12, ()), ("p8", 46, 82, 8, 0.0, 13, ()),
    ("p5n", 46, 82, 5, 0.03, 14, ()), ("p15n", 46, 82, 15, 0.03, 15, ()), ("p2_small", 32, 57, 2, 0.02, 16, ()),
    ("p4_missing", 46, 82, 4, 0.0, 17, (1, 2, 8)), ("p3_face", 46, 82, 3, 0.0, 18, tuple(range(1, 14))),
    ("noise_only", 20, 24, 0, 0.25, 19, ()), ("p30n", 46, 82, 30, 0.02, 20, ()),
]


def postproc_maps(case):
    name, h, w, persons, noise, seed, zero = case
    hm, paf, _ = synth.synthetic_pose_maps(1, h, w, seed=seed, noise=noise, persons=persons)
    hm, paf = hm[0], paf[0]
    for c in zero:
        hm[c] = 0
    return hm, paf


def run_reference_postproc(ref, hm, paf, demo, upsample=4):
    import cv2
    heat = cv2.resize(np.ascontiguousarray(hm.transpose(1, 2, 0)), (0, 0), fx=upsample, fy=upsample,
                      interpolation=cv2.INTER_CUBIC)
    pafs = cv2.resize(np.ascontiguousarray(paf.transpose(1, 2, 0)), (0, 0), fx=upsample, fy=upsample,
                      interpolation=cv2.INTER_CUBIC)
    total = 0
    by_type = []
    for k in range(18):
        total += ref.keypoints.extract_keypoints(heat[:, :, k], by_type, total)
    poses, allk = ref.keypoints.group_keypoints(by_type, pafs, demo=demo)
    return by_type, poses, allk


def pack_keypoints(by_type):
    rows = []
    for c, lst in enumerate(by_type):
        for (x, y, s, i) in lst:
            rows.append((c, int(x), int(y), int(np.float32(s).view(np.int32)), int(i)))
    return np.asarray(rows, np.int64).reshape(-1, 5)


def gen_postproc(out, ref):
    for case in POSTPROC_CASES:
        name = case[0]
        hm, paf = postproc_maps(case)
        out["pp_%s_in_sha" % name] = np.array(sha(hm) + sha(paf))
        for demo in (True, False):
            by_type, poses, allk = run_reference_postproc(ref, hm.copy(), paf.copy(), demo)
            tag = "pp_%s_%s" % (name, "demo" if demo else "val")
            out[tag + "_kpts"] = pack_keypoints(by_type)
            out[tag + "_poses"] = np.asarray(poses, np.float64).reshape(-1, 20) if len(poses) else np.zeros((0, 20))
            out[tag + "_allk"] = np.asarray(allk, np.float64).reshape(-1, 4) if len(allk) else np.zeros((0, 4))


# uint8 frame resize (demo.py:59): (h, w, scale numerator, denominator) ; scale = num / den as infer_fast computes it
U8_CASES = [(720, 1280, 256, 720), (480, 640, 368, 480), (180, 320, 256, 180), (375, 500, 368, 375), (97, 131, 194, 97)]


def gen_u8(out):
    """cv2's GENERIC uint8 cubic path: IPP is switched off for these calls (IPP-enabled builds differ by +-1 in ~5 % of
    the pixels; that arithmetic is not published and is not what the oracle / the GPU kernel restate)."""
    import cv2
    was = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)
    try:
        for i, (h, w, num, den) in enumerate(U8_CASES):
            img = np.random.default_rng(2000 + i).integers(0, 256, (h, w, 3), dtype=np.uint8)
            img[::7, ::5] = 255
            img[3::7, 2::5] = 0
            scale = num / den
            dst = cv2.resize(img, (0, 0), fx=scale, fy=scale, interpolation=cv2.INTER_CUBIC)
            out["u8_%d_sha" % i] = np.array(sha(dst))
            out["u8_%d_shape" % i] = np.array(dst.shape)
            if dst.size <= 120000:
                out["u8_%d_out" % i] = dst
        cv2.ipp.setUseIPP(True)
        img = np.random.default_rng(2000).integers(0, 256, (720, 1280, 3), dtype=np.uint8)
        img[::7, ::5] = 255
        img[3::7, 2::5] = 0
        d_ipp = cv2.resize(img, (0, 0), fx=256 / 720, fy=256 / 720, interpolation=cv2.INTER_CUBIC)
        cv2.ipp.setUseIPP(False)
        d_gen = cv2.resize(img, (0, 0), fx=256 / 720, fy=256 / 720, interpolation=cv2.INTER_CUBIC)
        diff = np.abs(d_ipp.astype(np.int32) - d_gen.astype(np.int32))
        out["u8_ipp_vs_generic"] = np.array([int(diff.max()), int((diff != 0).sum()), diff.size])
    finally:
        cv2.ipp.setUseIPP(was)


def u8_input(i):
    h, w, _, _ = U8_CASES[i]
    img = np.random.default_rng(2000 + i).integers(0, 256, (h, w, 3), dtype=np.uint8)
    img[::7, ::5] = 255
    img[3::7, 2::5] = 0
    return img


POSE_CASES = [("p3", (0, 3, 0, 5), 256 / 720), ("p8", (2, 0, 1, 0), 368 / 480), ("p15n", (0, 0, 0, 1), 0.3555555555555555)]


def ref_poses(ref, pose_mod, case_name, pad, scale, stride=8, upsample=4):
    """demo.py:100-115 through the real reference objects: returns (list of reference Pose, pose_entries, all_keypoints)."""
    case = [c for c in POSTPROC_CASES if c[0] == case_name][0]
    hm, paf = postproc_maps(case)
    _, pose_entries, all_keypoints = run_reference_postproc(ref, hm.copy(), paf.copy(), True, upsample)
    raw = np.array(all_keypoints, np.float64).copy()
    for kpt_id in range(all_keypoints.shape[0]):   # the reference's own statements (demo.py:101-103)
        all_keypoints[kpt_id, 0] = (all_keypoints[kpt_id, 0] * stride / upsample - pad[1]) / scale
        all_keypoints[kpt_id, 1] = (all_keypoints[kpt_id, 1] * stride / upsample - pad[0]) / scale
    poses = []
    for n in range(len(pose_entries)):
        pose_keypoints = np.ones((18, 2), dtype=np.int32) * -1
        for kpt_id in range(18):
            if pose_entries[n][kpt_id] != -1.0:
                pose_keypoints[kpt_id, 0] = int(all_keypoints[int(pose_entries[n][kpt_id]), 0])
                pose_keypoints[kpt_id, 1] = int(all_keypoints[int(pose_entries[n][kpt_id]), 1])
        poses.append(pose_mod.Pose(pose_keypoints, pose_entries[n][18]))
    return poses, np.asarray(pose_entries, np.float64).reshape(-1, 20), raw


def gen_pose(out, ref):
    import importlib
    pose_mod = importlib.import_module("modules.pose")   # the reference's Pose / track_poses
    for name, pad, scale in POSE_CASES:
        poses, entries, raw = ref_poses(ref, pose_mod, name, pad, scale)
        out["pose_%s_kpts" % name] = np.stack([p.keypoints for p in poses]).astype(np.int32)
        out["pose_%s_bbox" % name] = np.asarray([p.bbox for p in poses], np.int32)
        out["pose_%s_conf" % name] = np.asarray([p.confidence for p in poses], np.float64)
    # tracking over three "frames": the same persons drifting by a few pixels, smooth on and off
    for smooth in (False, True):
        pose_mod.Pose.last_id = -1
        prev = []
        for f in range(3):
            poses, _, _ = ref_poses(ref, pose_mod, "p8", (0, 0, 0, 0), 1.0)
            rng = np.random.default_rng(50 + f)
            for p in poses:
                m = p.keypoints[:, 0] != -1
                p.keypoints[m] += rng.integers(-3, 4, size=(int(m.sum()), 2)).astype(np.int32) + 2 * f
                p.bbox = pose_mod.Pose.get_bbox(p.keypoints)
            if f == 2:
                poses = poses[::-1][:-2]   # two persons leave, order changes
            out["track_s%d_f%d_in" % (smooth, f)] = np.stack([p.keypoints for p in poses]).astype(np.int32)
            out["track_s%d_f%d_conf" % (smooth, f)] = np.asarray([p.confidence for p in poses], np.float64)
            pose_mod.track_poses(prev, poses, smooth=smooth)
            out["track_s%d_f%d_ids" % (smooth, f)] = np.asarray([p.id for p in poses], np.int64)
            out["track_s%d_f%d_out" % (smooth, f)] = np.stack([p.keypoints for p in poses]).astype(np.int32)
            out["track_s%d_f%d_bbox" % (smooth, f)] = np.asarray([p.bbox for p in poses], np.int32)
            prev = poses
    sim = [[pose_mod.get_similarity(a, b) for b in prev] for a in prev]
    out["track_similarity"] = np.asarray(sim, np.int64)
    # convert_to_coco_format (val.py:52-78) on the val-mode (demo=False) tables of two cases
    for name in ("p8", "p15n"):
        case = [c for c in POSTPROC_CASES if c[0] == name][0]
        hm, paf = postproc_maps(case)
        _, pose_entries, all_keypoints = run_reference_postproc(ref, hm.copy(), paf.copy(), False)
        coco, scores = ref.val.convert_to_coco_format(pose_entries, all_keypoints)
        out["coco_%s_kpts" % name] = np.asarray(coco, np.float64)
        out["coco_%s_scores" % name] = np.asarray(scores, np.float64)
    # One-Euro filter trace
    oef = importlib.import_module("modules.one_euro_filter")
    f = oef.OneEuroFilter(freq=15, beta=0.1)
    out["one_euro_trace"] = np.asarray([f(v + (-1) ** (v % 2)) for v in range(12)], np.float64)


# (name, refinement stages, H, W, batch, head gain)
NET_CASES = [("r1_64x96", 1, 64, 96, 1, 1.0), ("r2_48x72", 2, 48, 72, 2, 1.0), ("r1_gain", 1, 64, 64, 1, 4.0)]


def gen_net(out, ref):
    import torch
    torch.set_num_threads(1)
    for name, R, H, W, B, gain in NET_CASES:
        torch.manual_seed(0)
        net = ref.with_mobilenet.PoseEstimationWithMobileNet(num_refinement_stages=R).eval()
        sd = net.state_dict()
        out["net_%s_init_sha" % name] = np.array(sha(np.concatenate([v.numpy().astype(np.float64).ravel()
                                                                     for v in sd.values()])))
        out["net_%s_keys" % name] = np.array(list(sd.keys()))
        synth.randomize_bn_(net, seed=7)
        if gain != 1.0:
            synth.apply_head_gain_(net, gain)
        x = synth.synthetic_net_input(B, H, W, seed=3)
        with torch.no_grad():
            ys = net(x)
        for i, y in enumerate(ys):
            out["net_%s_out%d" % (name, i)] = y.numpy()


def main():
    ref = refimport.load()
    r, p, n = {}, {}, {}
    gen_resize(r)
    np.savez_compressed(os.path.join(HERE, "resize_golden.npz"), **r)
    gen_postproc(p, ref)
    np.savez_compressed(os.path.join(HERE, "postproc_golden.npz"), **p)
    gen_net(n, ref)
    np.savez_compressed(os.path.join(HERE, "net_golden.npz"), **n)
    q = {}
    gen_pose(q, ref)
    np.savez_compressed(os.path.join(HERE, "pose_golden.npz"), **q)
    u = {}
    gen_u8(u)
    np.savez_compressed(os.path.join(HERE, "u8_golden.npz"), **u)
    for f in ("resize_golden.npz", "postproc_golden.npz", "net_golden.npz", "pose_golden.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
