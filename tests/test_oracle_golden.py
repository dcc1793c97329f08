"""Pins the oracle (oracle/) to the reference: golden vectors made by the real reference + cv2
(tests/golden/make_golden.py), and -- when /root/reference is present -- the live reference on extra
seeded inputs.  CPU only."""
import numpy as np
import pytest

import golden_cases as gc
import refimport
from oracle import postproc as orc


@pytest.mark.parametrize("idx,case", gc.resize_cases())
def test_resize_matches_cv2_golden(idx, case):
    g = gc.load("resize_golden.npz")
    h, w, C, mode, arg = case
    src = gc.resize_input(idx, h, w, C)
    dst = orc.resize_cubic(src, fx=arg, fy=arg) if mode == "f" else orc.resize_cubic(src, dsize=tuple(arg))
    assert tuple(dst.shape) == tuple(g["resize_%d_shape" % idx])
    assert gc.sha(dst) == str(g["resize_%d_sha" % idx])
    key = "resize_%d_out" % idx
    if key in g:
        assert np.array_equal(dst.view(np.int32), g[key].view(np.int32))


def _oracle_postproc(hm, paf, demo):
    heat = orc.resize_cubic(np.ascontiguousarray(hm.transpose(1, 2, 0)), fx=4, fy=4)
    pafs = orc.resize_cubic(np.ascontiguousarray(paf.transpose(1, 2, 0)), fx=4, fy=4)
    total, by_type = 0, []
    for k in range(18):
        total += orc.extract_keypoints(heat[:, :, k], by_type, total)
    poses, allk = orc.group_keypoints(by_type, pafs, demo=demo)
    return by_type, poses, allk


@pytest.mark.parametrize("case", gc.postproc_cases(), ids=lambda c: c[0])
@pytest.mark.parametrize("demo", [True, False])
def test_postproc_matches_reference_golden(case, demo):
    g = gc.load("postproc_golden.npz")
    hm, paf = gc.postproc_maps(case)
    assert gc.sha(hm) + gc.sha(paf) == str(g["pp_%s_in_sha" % case[0]]), "synthetic generator drifted"
    by_type, poses, allk = _oracle_postproc(hm, paf, demo)
    tag = "pp_%s_%s" % (case[0], "demo" if demo else "val")
    assert np.array_equal(gc.pack_keypoints(by_type), g[tag + "_kpts"])
    poses = np.asarray(poses, np.float64).reshape(-1, 20)
    assert poses.shape == g[tag + "_poses"].shape
    assert np.array_equal(poses.view(np.int64), g[tag + "_poses"].view(np.int64))
    allk = np.asarray(allk, np.float64).reshape(-1, 4)
    assert np.array_equal(allk.view(np.int64), g[tag + "_allk"].view(np.int64))


def test_empty_outputs_have_reference_shapes():
    hm = np.zeros((19, 8, 8), np.float32)
    paf = np.zeros((38, 8, 8), np.float32)
    by_type, poses, allk = _oracle_postproc(hm, paf, True)
    assert [len(l) for l in by_type] == [0] * 18
    assert poses.shape == (0,) and allk.shape == (0,)


@pytest.mark.skipif(not refimport.available(), reason="live reference not present (GPU box)")
@pytest.mark.parametrize("seed", range(6))
def test_postproc_matches_live_reference(seed):
    import cv2
    from lwpose_b200 import synth
    ref = refimport.load()
    rng = np.random.default_rng(seed)
    persons = int(rng.integers(0, 7))
    hm, paf, _ = synth.synthetic_pose_maps(1, 24, 31, seed=100 + seed, noise=float(rng.uniform(0, 0.08)),
                                           persons=persons)
    hm, paf = hm[0], paf[0]
    for demo in (True, False):
        heat = cv2.resize(np.ascontiguousarray(hm.transpose(1, 2, 0)), (0, 0), fx=4, fy=4,
                          interpolation=cv2.INTER_CUBIC)
        pafs = cv2.resize(np.ascontiguousarray(paf.transpose(1, 2, 0)), (0, 0), fx=4, fy=4,
                          interpolation=cv2.INTER_CUBIC)
        total, ref_by_type = 0, []
        for k in range(18):
            total += ref.keypoints.extract_keypoints(heat[:, :, k], ref_by_type, total)
        ref_poses, ref_allk = ref.keypoints.group_keypoints(ref_by_type, pafs, demo=demo)
        by_type, poses, allk = _oracle_postproc(hm, paf, demo)
        assert np.array_equal(gc.pack_keypoints(by_type), gc.pack_keypoints(ref_by_type))
        assert np.asarray(poses).shape == np.asarray(ref_poses).shape
        assert np.array_equal(np.asarray(poses, np.float64).view(np.int64),
                              np.asarray(ref_poses, np.float64).view(np.int64))


@pytest.mark.skipif(not refimport.available(), reason="live cv2 cross-check runs in the build container")
@pytest.mark.parametrize("seed", range(12))
def test_resize_matches_live_cv2(seed):
    import cv2
    rng = np.random.default_rng(seed)
    h, w, C = int(rng.integers(3, 40)), int(rng.integers(3, 40)), int(rng.choice([2, 5, 6, 7, 19, 38]))
    src = rng.standard_normal((h, w, C)).astype(np.float32)
    if seed % 2:
        W, H = int(rng.integers(w, 4 * w)), int(rng.integers(h, 4 * h))
        ref = cv2.resize(src, (W, H), interpolation=cv2.INTER_CUBIC)
        got = orc.resize_cubic(src, dsize=(W, H))
    else:
        f = float(rng.choice([2, 3, 4, 8, 1.5]))
        ref = cv2.resize(src, (0, 0), fx=f, fy=f, interpolation=cv2.INTER_CUBIC)
        got = orc.resize_cubic(src, fx=f, fy=f)
    ref = ref.reshape(got.shape)
    assert np.array_equal(got.view(np.int32), ref.view(np.int32)), (h, w, C, got.shape)


@pytest.mark.parametrize("case", gc.net_cases(), ids=lambda c: c[0])
def test_net_oracle_matches_reference_golden(case):
    """oracle/net.py (functional torch-CPU restatement) vs outputs of the real reference module."""
    import torch
    from lwpose_b200 import synth
    from lwpose_b200.models.with_mobilenet import PoseEstimationWithMobileNet
    from oracle import net as onet
    name, R, H, W, B, gain = case
    g = gc.load("net_golden.npz")
    torch.manual_seed(0)
    net = PoseEstimationWithMobileNet(num_refinement_stages=R).eval()
    sd = net.state_dict()
    import numpy as np
    assert gc.sha(np.concatenate([v.numpy().astype(np.float64).ravel() for v in sd.values()])) == \
        str(g["net_%s_init_sha" % name]), "seeded initialisation differs from the reference's"
    synth.randomize_bn_(net, seed=7)
    if gain != 1.0:
        synth.apply_head_gain_(net, gain)
    torch.set_num_threads(1)
    outs = onet.forward(net.state_dict(), synth.synthetic_net_input(B, H, W, seed=3))
    assert len(outs) == 2 * (1 + R)
    for i, y in enumerate(outs):
        ref = torch.from_numpy(g["net_%s_out%d" % (name, i)])
        assert float((y - ref).abs().max()) < 2e-6 * gain, i


def test_oracle_u8_resize_matches_cv2_generic_path_golden():
    """orc_resize_pad_u8 (OpenCV's generic fixed-point uint8 cubic path: demo.py:59) against cv2 outputs taken with IPP off;
    the fixtures also record how far an IPP-enabled cv2 is from that path (+-1 on ~5 % of the pixels)."""
    import golden_cases as gc
    from oracle import postproc as orc
    g = gc.load("u8_golden.npz")
    mg = gc._mg()
    for i, (h, w, num, den) in enumerate(mg.U8_CASES):
        r = orc.resize_pad_u8(mg.u8_input(i), fx=num / den, fy=num / den)
        assert tuple(r.shape) == tuple(g["u8_%d_shape" % i]) and gc.sha(r) == str(g["u8_%d_sha" % i])
    assert int(g["u8_ipp_vs_generic"][0]) == 1
    # pad placement (val.py:36-49): the resized frame sits at (top, left), the rest is the pad value
    img = mg.u8_input(4)
    r = orc.resize_pad_u8(img, fx=2.0, fy=2.0)
    p = orc.resize_pad_u8(img, fx=2.0, fy=2.0, padded=(200, 272), top=3, left=5, pad_value=(128, 127, 126))
    assert np.array_equal(p[3:3 + 194, 5:5 + 262], r) and tuple(p[0, 0]) == (128, 127, 126) and tuple(p[199, 271]) == (128, 127, 126)
