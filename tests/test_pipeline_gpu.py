"""GPU tests of the fused batched pipeline and of the drop-in infer_fast / infer mirrors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import synth
    from lwpose_b200.models.with_mobilenet import PoseEstimationWithMobileNet
    torch.manual_seed(0)
    net = PoseEstimationWithMobileNet(num_refinement_stages=1).eval()
    synth.randomize_bn_(net, seed=7)
    return torch, net.cuda()


def _oracle_post(heads_b, demo):
    from oracle import postproc as orc
    heat = orc.resize_cubic(np.ascontiguousarray(heads_b[:, :, :19]), fx=4, fy=4)
    pafs = orc.resize_cubic(np.ascontiguousarray(heads_b[:, :, 19:57]), fx=4, fy=4)
    total, by_type = 0, []
    for k in range(18):
        total += orc.extract_keypoints(heat[:, :, k], by_type, total)
    return by_type, orc.group_keypoints(by_type, pafs, demo=demo)


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_pipeline_end_to_end(env, precision):
    """Host frames in -> host pose tables out; heads within tolerance of the CPU fp32 oracle, and the
    pose tables bit-exact against the oracle run on the very maps the GPU produced."""
    torch, net = env
    import golden_cases as gc
    from lwpose_b200 import synth
    from lwpose_b200.pipeline import PosePipeline
    from oracle import net as onet
    B, H, W = 4, 128, 192
    hm, paf, _ = synth.synthetic_pose_maps(B, H // 8, W // 8, seed=9, max_persons=3)
    inj = np.zeros((B, H // 8, W // 8, 64), np.float32)
    inj[..., :19] = hm.transpose(0, 2, 3, 1)
    inj[..., 19:57] = paf.transpose(0, 2, 3, 1)
    inj_d = torch.from_numpy(inj).cuda()
    pipe = PosePipeline(net, B, H, W, precision=precision, demo=True, chunk=3,
                        heads_hook=lambda t, lo: t.add_(inj_d[lo:lo + t.shape[0]]))
    x = synth.synthetic_net_input(B, H, W, seed=2)
    res = pipe(x.pin_memory()).check()
    assert pipe.error_flag() == 0
    heads = pipe.heads.cpu().numpy()
    ref = onet.forward(net.state_dict(), x)
    ref_heads = np.concatenate([ref[-2].numpy(), ref[-1].numpy()], 1).transpose(0, 2, 3, 1) + inj[..., :57]
    tol = 1e-3 if precision == "tf32" else 3e-3
    assert np.abs(heads[..., :57] - ref_heads).max() < tol
    assert np.all(heads[..., 57:] == 0)
    n_total = 0
    for b in range(B):
        by_type, (ref_poses, ref_allk) = _oracle_post(heads[b], True)
        assert gc.pack_keypoints(res.keypoints_by_type(b)).tolist() == gc.pack_keypoints(by_type).tolist()
        poses, allk = res.frame(b)
        rp = np.asarray(ref_poses, np.float64).reshape(-1, 20)
        gp = np.asarray(poses, np.float64).reshape(-1, 20)
        assert rp.shape == gp.shape and np.array_equal(rp.view(np.int64), gp.view(np.int64))
        assert np.array_equal(np.asarray(allk, np.float64), np.asarray(ref_allk, np.float64))
        n_total += len(rp)
    assert n_total >= B  # the injected persons were found
    # device-resident entry point gives the same tables
    pipe.run_device(x.cuda())
    torch.cuda.synchronize()
    assert int(pipe.n_poses.sum()) == n_total


def test_infer_fast_matches_oracle(env):
    """demo.infer_fast mirror: same return signature; maps = cubic x4 of the network output (bit-exact
    resize of our own heads, heads within tolerance of the CPU oracle)."""
    torch, net = env
    from lwpose_b200 import demo
    from oracle import postproc as orc
    net.precision = "tf32"
    img = np.random.default_rng(0).integers(0, 256, (180, 320, 3), dtype=np.uint8)
    heat, pafs, scale, pad = demo.infer_fast(net, img, 128, 8, 4, False)
    assert heat.dtype == np.float32 and heat.shape[2] == 19 and pafs.shape[2] == 38
    assert scale == 128 / 180 and len(pad) == 4
    H, W = heat.shape[0] // 4 * 8, heat.shape[1] // 4 * 8
    plan = net.engine().plan("tf32", 1, H, W)
    heads = plan.heads_f32[-1].view(1, H // 8, W // 8, 64).cpu().numpy()[0]
    assert np.array_equal(heat.view(np.int32), orc.resize_cubic(heads[:, :, :19].copy(), fx=4, fy=4).view(np.int32))
    assert np.array_equal(pafs.view(np.int32), orc.resize_cubic(heads[:, :, 19:57].copy(), fx=4, fy=4).view(np.int32))
    with pytest.raises(RuntimeError):
        demo.infer_fast(net, img, 128, 8, 4, True)
    poses, allk = demo.run_frame(net, img, 128)
    assert poses.shape == (0,) or poses.shape[1] == 20


def test_val_infer_multiscale_matches_oracle(env):
    """val.infer mirror (multi-scale, x8 cubic, crop, resize to the image size, average)."""
    torch, net = env
    import cv2
    from lwpose_b200 import val
    from oracle import net as onet
    from oracle import postproc as orc
    net.precision = "tf32"
    img = np.random.default_rng(1).integers(0, 256, (96, 128, 3), dtype=np.uint8)
    scales, base = [0.5, 1.0], 64
    got_h, got_p = val.infer(net, img, scales, base, 8)
    assert got_h.shape == (96, 128, 19) and got_p.shape == (96, 128, 38) and got_h.dtype == np.float32
    # oracle restatement of val.py:81-110 with the CPU fp32 network
    normed = val.normalize(img, (128, 128, 128), 1 / 256)
    avg_h = np.zeros((96, 128, 19), np.float32)
    avg_p = np.zeros((96, 128, 38), np.float32)
    for s in scales:
        ratio = s * base / 96.0
        scaled = cv2.resize(normed, (0, 0), fx=ratio, fy=ratio, interpolation=cv2.INTER_CUBIC)
        padded, pad = val.pad_width(scaled, 8, (0, 0, 0), [base, max(scaled.shape[1], base)])
        x = torch.from_numpy(padded).permute(2, 0, 1).unsqueeze(0).float()
        outs = onet.forward(net.state_dict(), x)
        for o, avg in ((outs[-2], avg_h), (outs[-1], avg_p)):
            m = np.ascontiguousarray(o[0].numpy().transpose(1, 2, 0))
            m = orc.resize_cubic(m, fx=8, fy=8)
            m = m[pad[0]:m.shape[0] - pad[2], pad[1]:m.shape[1] - pad[3], :]
            m = orc.resize_cubic(np.ascontiguousarray(m), dsize=(128, 96))
            avg += m / len(scales)
    assert np.abs(got_h - avg_h).max() < 1e-3 and np.abs(got_p - avg_p).max() < 1e-3


def test_config1_infer_fast_720p_frame(env):
    """BASELINE.json configs[0]: one 720x1280 BGR frame through infer_fast at height 256 -> net input
    1x3x256x456 (pad [0,0,0,1]), then the drop-in extract/group calls (demo=True)."""
    torch, net = env
    from lwpose_b200 import demo
    from oracle import net as onet
    from lwpose_b200 import val
    import cv2
    net.precision = "tf32"
    img = np.random.default_rng(0).integers(0, 256, (720, 1280, 3), dtype=np.uint8)
    heat, pafs, scale, pad = demo.infer_fast(net, img, 256, 8, 4, False)
    assert heat.shape == (128, 228, 19) and pafs.shape == (128, 228, 38)
    assert pad == [0, 0, 0, 1] and scale == 256 / 720
    # same pre-processing as the reference, network through the CPU oracle
    scaled = cv2.resize(img, (0, 0), fx=scale, fy=scale, interpolation=cv2.INTER_CUBIC)
    padded, _ = val.pad_width(val.normalize(scaled, (128, 128, 128), 1 / 256), 8, (0, 0, 0), [256, max(scaled.shape[1], 256)])
    x = torch.from_numpy(padded).permute(2, 0, 1).unsqueeze(0).float()
    assert tuple(x.shape) == (1, 3, 256, 456)
    ref = onet.forward(net.state_dict(), x)
    from oracle import postproc as orc
    ref_heat = orc.resize_cubic(np.ascontiguousarray(ref[-2][0].numpy().transpose(1, 2, 0)), fx=4, fy=4)
    assert np.abs(heat - ref_heat).max() < 1e-3
    poses, allk = demo.run_frame(net, img, 256)
    assert poses.shape == (0,) and allk.shape == (0,)   # random-init weights stay below the 0.1 threshold


def test_three_refinement_stages_pipeline(env):
    """BASELINE.json configs[3] at a small size: R = 3 network through the batched pipeline."""
    torch, _ = env
    from lwpose_b200 import synth
    from lwpose_b200.models.with_mobilenet import PoseEstimationWithMobileNet
    from lwpose_b200.pipeline import PosePipeline
    from oracle import net as onet
    torch.manual_seed(0)
    net = PoseEstimationWithMobileNet(num_refinement_stages=3).eval()
    synth.randomize_bn_(net, seed=5)
    x = synth.synthetic_net_input(3, 64, 96, seed=4)
    ref = onet.forward(net.state_dict(), x)
    net = net.cuda()
    pipe = PosePipeline(net, 3, 64, 96, precision="bf16")
    pipe(x.pin_memory()).check()
    assert pipe.error_flag() == 0
    heads = pipe.heads.cpu().numpy()
    ref_heads = np.concatenate([ref[-2].numpy(), ref[-1].numpy()], 1).transpose(0, 2, 3, 1)
    assert np.abs(heads[..., :57] - ref_heads).max() < 3e-3
    outs = net(x.cuda())
    assert len(outs) == 8
    for o, r in zip(outs, ref):
        assert float((o.cpu() - r).abs().max()) < 1e-3   # module forward defaults to tf32


def test_streaming_submit_collect_order(env):
    """Two batches in flight come back in submission order with their own results."""
    torch, net = env
    from lwpose_b200 import synth
    from lwpose_b200.pipeline import PosePipeline
    B, H, W = 2, 64, 96
    marks = []
    for k in (1, 2):
        hm, paf, _ = synth.synthetic_pose_maps(B, H // 8, W // 8, seed=20 + k, persons=k)
        inj = np.zeros((B, H // 8, W // 8, 64), np.float32)
        inj[..., :19] = hm.transpose(0, 2, 3, 1)
        inj[..., 19:57] = paf.transpose(0, 2, 3, 1)
        marks.append(torch.from_numpy(inj).cuda())
    state = {"i": 0}
    pipe = PosePipeline(net, B, H, W, precision="bf16", heads_hook=lambda t, lo: t.add_(marks[state["i"]][lo:lo + t.shape[0]]))
    x = synth.synthetic_net_input(B, H, W, seed=2).pin_memory()
    expect = []
    for k in (0, 1):   # synchronous calls give the expected tables of each batch
        state["i"] = k
        r = pipe(x).check()
        expect.append((r.n_poses.copy(), r.pose_entries.copy()))
    assert expect[0][0].sum() >= 1 and not np.array_equal(expect[0][1], expect[1][1])
    state["i"] = 0
    pipe.submit(x)
    state["i"] = 1
    pipe.submit(x)
    with pytest.raises(RuntimeError):
        pipe.submit(x)   # depth 2
    for k in (0, 1):
        r = pipe.collect()
        assert np.array_equal(r.n_poses, expect[k][0])
        for b in range(B):
            n = int(r.n_poses[b])
            assert np.array_equal(r.pose_entries[b, :n], expect[k][1][b, :n])


@pytest.mark.parametrize("norm", [((128, 128, 128), 1 / 256, False), ((128, 128, 128), 1 / 256, True),
                                  ((104.5, 117.0, 123.25), 1 / 255.0, False), ((0, 255, 7), 0.5, False)],
                         ids=["default_f32_exact", "default_forced_f64", "general_f64", "pow2_f32_exact"])
def test_uint8_frames_equal_normalised_float_input(env, norm, monkeypatch):
    """input_format='u8_nhwc': raw BGR frames with val.normalize fused into the stem give bit-identical heads and pose
    tables to feeding the reference's normalised float32 NCHW tensor -- on the exact float32 short cut the stem takes
    for integer means and a power-of-two scale (the reference's constants) and on the general float64 path."""
    torch, net = env
    from lwpose_b200 import synth, val
    from lwpose_b200.pipeline import PosePipeline
    mean, scale, force_f64 = norm
    if force_f64:
        monkeypatch.setenv("LWP_STEM_F64_NORM", "1")
    B, H, W = 3, 64, 96
    frames = synth.synthetic_frames(B, H, W, seed=3)
    x_f32 = torch.from_numpy(np.stack([val.normalize(f, mean, scale) for f in frames])).permute(0, 3, 1, 2).float().contiguous()
    pf = PosePipeline(net, B, H, W, precision="bf16")
    p8 = PosePipeline(net, B, H, W, precision="bf16", input_format="u8_nhwc", img_mean=mean, img_scale=scale)
    rf = pf(x_f32.pin_memory()).check()
    hf = pf.heads.cpu().numpy()
    r8 = p8(torch.from_numpy(frames).pin_memory()).check()
    h8 = p8.heads.cpu().numpy()
    assert np.array_equal(hf.view(np.int32), h8.view(np.int32))
    assert np.array_equal(rf.n_poses, r8.n_poses)
    assert p8.h2d_bytes * 4 == pf.h2d_bytes


def test_pipeline_convert_gives_pose_objects(env):
    """PosePipeline(convert=...): the result post-conversion of demo.py:101-115 runs on the device; PoseResult.poses(i)
    returns Pose objects whose key-points / boxes / confidences equal the oracle's restatement on the same tables."""
    torch, net = env
    from lwpose_b200 import synth
    from lwpose_b200.pipeline import PosePipeline
    from oracle import postproc as orc
    B, H, W = 3, 128, 192
    hm, paf, _ = synth.synthetic_pose_maps(B, H // 8, W // 8, seed=41, max_persons=3)
    inj = np.zeros((B, H // 8, W // 8, 64), np.float32)
    inj[..., :19] = hm.transpose(0, 2, 3, 1)
    inj[..., 19:57] = paf.transpose(0, 2, 3, 1)
    inj_d = torch.from_numpy(inj).cuda()
    pads = [[0, 3, 0, 5], [2, 0, 1, 0], [0, 0, 0, 1]]
    scales = [256 / 720, 0.5, 1.25]
    pipe = PosePipeline(net, B, H, W, precision="bf16", demo=True, chunk=2, convert=dict(pad=pads, scale=scales),
                        heads_hook=lambda t, lo: t.add_(inj_d[lo:lo + t.shape[0]]))
    res = pipe(synth.synthetic_net_input(B, H, W, seed=2).pin_memory()).check()
    total = 0
    for b in range(B):
        poses, allk = res.frame(b)
        kp, bbox, conf = orc.pose_convert(poses, allk, 8, 4, pads[b], scales[b])
        objs = res.poses(b)
        assert len(objs) == kp.shape[0]
        for j, p in enumerate(objs):
            assert np.array_equal(p.keypoints, kp[j]) and tuple(p.bbox) == tuple(int(v) for v in bbox[j])
            assert p.confidence == conf[j]
        total += len(objs)
    assert total >= B


@pytest.mark.parametrize("overlap", [True, False])
def test_one_crowded_frame_is_retried_not_fatal(env, overlap):
    """One frame exceeds the pipeline's fixed table capacities: only that frame is re-processed (larger tables) by
    collect(); every frame's result -- the crowded one included -- is bit-exact against the oracle, nothing raises."""
    torch, net = env
    import golden_cases as gc
    from lwpose_b200 import synth
    from lwpose_b200.pipeline import PosePipeline
    B, H, W = 3, 368, 656
    maps = [synth.synthetic_pose_maps(1, H // 8, W // 8, seed=60 + b, persons=p) for b, p in enumerate((1, 12, 2))]
    inj = np.zeros((B, H // 8, W // 8, 64), np.float32)
    for b, (hm, paf, _) in enumerate(maps):
        inj[b, :, :, :19] = hm[0].transpose(1, 2, 0)
        inj[b, :, :, 19:57] = paf[0].transpose(1, 2, 0)
    inj_d = torch.from_numpy(inj).cuda()
    pipe = PosePipeline(net, B, H, W, precision="bf16", demo=True, cap_kpts=8, cap_poses=8, overlap_postproc=overlap,
                        heads_hook=lambda t, lo: t.add_(inj_d[lo:lo + t.shape[0]]))
    x = synth.synthetic_net_input(B, H, W, seed=2).pin_memory()
    res = pipe(x)
    assert res.overflow.tolist() == [0, 1, 0] and list(res.retried) == [1]
    res.check()
    heads = pipe.heads.cpu().numpy()
    for b in range(B):
        by_type, (ref_poses, ref_allk) = _oracle_post(heads[b], True)
        assert gc.pack_keypoints(res.keypoints_by_type(b)).tolist() == gc.pack_keypoints(by_type).tolist()
        poses, _ = res.frame(b)
        rp = np.asarray(ref_poses, np.float64).reshape(-1, 20)
        gp = np.asarray(poses, np.float64).reshape(-1, 20)
        assert rp.shape == gp.shape and np.array_equal(rp.view(np.int64), gp.view(np.int64))
    assert len(res.frame(1)[0]) >= 12


def test_val_infer_batch_matches_oracle(env):
    """val.infer_batch (configs[4]: batched multi-scale inference entirely on the device) against the oracle restatement of
    val.py:81-110 (host cv2 float64 input resize, CPU fp32 network, OpenCV-exact output resizes) frame by frame, and
    evaluate_batch's COCO conversion against the oracle post-processing of the device's own averaged maps."""
    torch, net = env
    import cv2
    import golden_cases as gc
    from lwpose_b200 import postproc, val
    from oracle import net as onet
    from oracle import postproc as orc
    net.precision = "tf32"
    frames = np.random.default_rng(5).integers(0, 256, (2, 96, 128, 3), dtype=np.uint8)
    scales, base = [0.5, 1.0, 1.5], 64
    got_h, got_p = val.infer_batch(net, frames, scales, base, 8)
    assert tuple(got_h.shape) == (2, 96, 128, 19) and tuple(got_p.shape) == (2, 96, 128, 38)
    for b in range(2):
        normed = val.normalize(frames[b], (128, 128, 128), 1 / 256)
        avg_h = np.zeros((96, 128, 19), np.float32)
        avg_p = np.zeros((96, 128, 38), np.float32)
        for s in scales:
            ratio = s * base / 96.0
            scaled = cv2.resize(normed, (0, 0), fx=ratio, fy=ratio, interpolation=cv2.INTER_CUBIC)
            padded, pad = val.pad_width(scaled, 8, (0, 0, 0), [base, max(scaled.shape[1], base)])
            x = torch.from_numpy(padded).permute(2, 0, 1).unsqueeze(0).float()
            outs = onet.forward(net.state_dict(), x)
            for o, avg in ((outs[-2], avg_h), (outs[-1], avg_p)):
                m = orc.resize_cubic(np.ascontiguousarray(o[0].numpy().transpose(1, 2, 0)), fx=8, fy=8)
                m = m[pad[0]:m.shape[0] - pad[2], pad[1]:m.shape[1] - pad[3], :]
                avg += orc.resize_cubic(np.ascontiguousarray(m), dsize=(128, 96)) / len(scales)
        assert np.abs(got_h[b].cpu().numpy() - avg_h).max() < 1e-3 and np.abs(got_p[b].cpu().numpy() - avg_p).max() < 1e-3
    # end-to-end body of val.evaluate on the device vs the oracle post-processing of the same averaged maps
    hm, paf, _ = __import__("lwpose_b200").synth.synthetic_pose_maps(2, 96, 128, seed=3, persons=2)
    ah = torch.from_numpy(np.ascontiguousarray(hm.transpose(0, 2, 3, 1))).cuda()
    ap = torch.from_numpy(np.ascontiguousarray(paf.transpose(0, 2, 3, 1))).cuda()
    kb = postproc.extract_keypoints_batched(ah)
    poses_d, n_d = postproc.group_keypoints_batched(kb, ap, demo=False)
    kp, cnt, st, _ = kb.to_host()
    for b in range(2):
        by_type = postproc.keypoint_lists(kp, cnt, st, b)
        allk = np.array([item for sub in by_type for item in sub])
        got = val.convert_to_coco_format(postproc.pose_entries_array(poses_d.cpu().numpy(), n_d.cpu().numpy(), b), allk)
        total, ref_by = 0, []
        heat = ah[b].cpu().numpy().copy()
        for k in range(18):
            total += orc.extract_keypoints(heat[:, :, k], ref_by, total)
        rp, ra = orc.group_keypoints(ref_by, ap[b].cpu().numpy(), demo=False)
        want = val.convert_to_coco_format(rp, ra)
        assert len(got[0]) == len(want[0]) >= 2 and got[0] == want[0] and [float(v) for v in got[1]] == [float(v) for v in want[1]]
    res = val.evaluate_batch(net, frames, scales=[1.0], base_height=base)
    assert len(res) == 2 and all(len(r) == 2 for r in res)


def test_raw_camera_frames_pipeline(env):
    """input_format='u8_raw' (row f1): 180x320 camera frames are resized (cubic, OpenCV generic-path bits) and padded on the GPU;
    heads and pose tables are bit-identical to feeding the pipeline the frames prepared by the oracle on the host."""
    torch, net = env
    from lwpose_b200 import postproc, synth
    from lwpose_b200.pipeline import PosePipeline
    from oracle import postproc as orc
    B, h, w, hn = 2, 180, 320, 128
    scale, (H, W), (Hp, Wp), pad = postproc.infer_fast_geometry(h, w, hn)
    raw = synth.synthetic_frames(B, h, w, seed=8)
    host = np.stack([orc.resize_pad_u8(f, fx=scale, fy=scale, padded=(Hp, Wp), top=pad[0], left=pad[1]) for f in raw])
    p_raw = PosePipeline(net, B, Hp, Wp, precision="bf16", input_format="u8_raw", raw_size=(h, w))
    p_u8 = PosePipeline(net, B, Hp, Wp, precision="bf16", input_format="u8_nhwc")
    r1 = p_raw(torch.from_numpy(raw).pin_memory()).check()
    h1 = p_raw.heads.cpu().numpy()
    r2 = p_u8(torch.from_numpy(host).pin_memory()).check()
    h2 = p_u8.heads.cpu().numpy()
    assert np.array_equal(h1.view(np.int32), h2.view(np.int32)) and np.array_equal(r1.n_poses, r2.n_poses)
    assert p_raw.h2d_bytes == B * h * w * 3
    with pytest.raises(ValueError):
        PosePipeline(net, B, Hp, Wp + 8, precision="bf16", input_format="u8_raw", raw_size=(h, w))


def test_infer_fast_gpu_preprocess(env):
    """infer_fast(gpu_preprocess=True): same return signature; the maps equal those of the host path within the network
    tolerance (the host cv2 may be an IPP build: +-1 LSB on a few per cent of the input pixels)."""
    torch, net = env
    from lwpose_b200 import demo
    net.precision = "tf32"
    img = np.random.default_rng(0).integers(0, 256, (180, 320, 3), dtype=np.uint8)
    h0, p0, s0, pad0 = demo.infer_fast(net, img, 128, 8, 4, False)
    h1, p1, s1, pad1 = demo.infer_fast(net, img, 128, 8, 4, False, gpu_preprocess=True)
    assert h0.shape == h1.shape and p0.shape == p1.shape and s0 == s1 and list(pad0) == list(pad1)
    assert np.abs(h0 - h1).max() < 1e-3 and np.abs(p0 - p1).max() < 1e-3
