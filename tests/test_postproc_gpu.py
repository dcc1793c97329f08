"""GPU parity of the post-processing kernels (through the C ABI) against the golden vectors made by the
real reference and against the oracle.  Bit-exact: key-point tuples, pose entries, up-sampled maps."""
import numpy as np
import pytest

import golden_cases as gc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import lwpose_b200  # noqa: F401
    return torch


@pytest.mark.parametrize("idx,case", gc.resize_cases())
def test_upsample_bit_exact_vs_cv2_golden(torch_cuda, idx, case):
    torch = torch_cuda
    from lwpose_b200 import postproc
    g = gc.load("resize_golden.npz")
    h, w, C, mode, arg = case
    src = gc.resize_input(idx, h, w, C)
    d = torch.from_numpy(src).cuda().unsqueeze(0)
    out = postproc.upsample_cubic(d, fx=arg, fy=arg) if mode == "f" else postproc.upsample_cubic(d, dsize=tuple(arg))
    got = out[0].cpu().numpy()
    assert tuple(got.shape) == tuple(g["resize_%d_shape" % idx])
    assert gc.sha(got) == str(g["resize_%d_sha" % idx])


def test_upsample_strided_source_and_batch(torch_cuda):
    """Channels taken out of a wider pixel (ld 64) for a batch: what the fused pipeline does."""
    torch = torch_cuda
    from lwpose_b200 import postproc
    from oracle import postproc as orc
    rng = np.random.default_rng(5)
    src = rng.standard_normal((3, 11, 13, 64)).astype(np.float32)
    d = torch.from_numpy(src).cuda()
    got = postproc.upsample_cubic(d, channels=57, fx=4, fy=4).cpu().numpy()
    for b in range(3):
        ref = orc.resize_cubic(src[b, :, :, :57], fx=4, fy=4)
        assert np.array_equal(got[b].view(np.int32), ref.view(np.int32))


@pytest.mark.parametrize("ratio", [2, 4, 8])
def test_upsample_pow2_kernel_equals_generic_kernel_and_oracle(torch_cuda, ratio, monkeypatch):
    """The periodic-geometry kernel for x2 / x4 / x8 (one thread = one channel of one source cell, R x R outputs), the generic
    materialising kernel and the oracle agree bit for bit (19- and 38-channel layouts out of a 64-wide pixel, odd sizes)."""
    torch = torch_cuda
    from lwpose_b200 import postproc
    from oracle import postproc as orc
    rng = np.random.default_rng(17 + ratio)
    src = rng.standard_normal((2, 23, 41, 64)).astype(np.float32)
    d = torch.from_numpy(src).cuda()
    for off, ch in ((0, 19), (19, 38)):
        monkeypatch.delenv("LWP_UPSAMPLE_GENERIC", raising=False)
        fast = postproc.upsample_cubic(d, channels=ch, fx=ratio, fy=ratio, channel_offset=off).cpu().numpy()
        monkeypatch.setenv("LWP_UPSAMPLE_GENERIC", "1")
        slow = postproc.upsample_cubic(d, channels=ch, fx=ratio, fy=ratio, channel_offset=off).cpu().numpy()
        assert np.array_equal(fast.view(np.int32), slow.view(np.int32))
        ref = orc.resize_cubic(np.ascontiguousarray(src[0, :, :, off:off + ch]), fx=ratio, fy=ratio)
        assert np.array_equal(fast[0].view(np.int32), ref.view(np.int32))


def test_upsample_cropped_view_and_running_average(torch_cuda):
    """lwp_upsample_cubic_ex: resizing a cropped VIEW equals resizing the cropped copy, and the accumulate mode equals
    avg + resized / k in float32 (val.py:99-101), bit for bit."""
    torch = torch_cuda
    from lwpose_b200 import postproc
    rng = np.random.default_rng(23)
    d = torch.from_numpy(rng.standard_normal((2, 40, 56, 38)).astype(np.float32)).cuda()
    pad = (3, 5, 2, 7)
    crop = d[:, pad[0]:40 - pad[2], pad[1]:56 - pad[3], :].contiguous()
    ref = postproc.upsample_cubic(crop, dsize=(61, 47))
    got = postproc.upsample_cubic(d, dsize=(61, 47), crop=pad)
    assert torch.equal(ref.view(torch.int32), got.view(torch.int32))
    avg = torch.from_numpy(rng.standard_normal((2, 47, 61, 38)).astype(np.float32)).cuda()
    want = avg + torch.div(ref, torch.tensor(4.0, device="cuda"))
    postproc.upsample_cubic(d, dsize=(61, 47), crop=pad, out=avg, accumulate_divisor=4)
    assert torch.equal(want.view(torch.int32), avg.view(torch.int32))


def _device_postproc(torch, hm, paf, demo, cap_kpts=256, cap_cand=4096, cap_poses=256, cap_conn=4096):
    from lwpose_b200 import postproc
    hm_d = torch.from_numpy(np.ascontiguousarray(hm.transpose(0, 2, 3, 1))).cuda()
    paf_d = torch.from_numpy(np.ascontiguousarray(paf.transpose(0, 2, 3, 1))).cuda()
    heat = postproc.upsample_cubic(hm_d, fx=4, fy=4)
    pafs = postproc.upsample_cubic(paf_d, fx=4, fy=4)
    kb = postproc.extract_keypoints_batched(heat, cap_kpts=cap_kpts, cap_candidates=cap_cand)
    poses_d, n_d = postproc.group_keypoints_batched(kb, pafs, demo=demo, cap_poses=cap_poses,
                                                    cap_connections=cap_conn)
    kpts_h, counts_h, start_h, ovf = kb.to_host()
    postproc.raise_on_overflow(ovf)
    return kpts_h, counts_h, start_h, poses_d.cpu().numpy(), n_d.cpu().numpy()


@pytest.mark.parametrize("case", gc.postproc_cases(), ids=lambda c: c[0])
@pytest.mark.parametrize("demo", [True, False])
def test_postproc_bit_exact_vs_reference_golden(torch_cuda, case, demo):
    from lwpose_b200 import postproc
    g = gc.load("postproc_golden.npz")
    hm, paf = gc.postproc_maps(case)
    kpts_h, counts_h, start_h, poses_h, n_h = _device_postproc(torch_cuda, hm[None], paf[None], demo,
                                                               cap_poses=2048, cap_conn=8192)
    by_type = postproc.keypoint_lists(kpts_h, counts_h, start_h, 0)
    tag = "pp_%s_%s" % (case[0], "demo" if demo else "val")
    assert np.array_equal(gc.pack_keypoints(by_type), g[tag + "_kpts"])
    poses = np.asarray(postproc.pose_entries_array(poses_h, n_h, 0), np.float64).reshape(-1, 20)
    assert poses.shape == g[tag + "_poses"].shape
    assert np.array_equal(poses.view(np.int64), g[tag + "_poses"].view(np.int64))


def test_postproc_batch_vs_oracle(torch_cuda):
    """A batch with 1..12 persons + noise, every frame checked against the oracle (bit-exact)."""
    from lwpose_b200 import postproc, synth
    from oracle import postproc as orc
    hm, paf, _ = synth.synthetic_pose_maps(12, 46, 82, seed=3, noise=0.02, max_persons=12)
    for demo in (True, False):
        kpts_h, counts_h, start_h, poses_h, n_h = _device_postproc(torch_cuda, hm, paf, demo)
        for b in range(hm.shape[0]):
            heat = orc.resize_cubic(np.ascontiguousarray(hm[b].transpose(1, 2, 0)), fx=4, fy=4)
            pafs = orc.resize_cubic(np.ascontiguousarray(paf[b].transpose(1, 2, 0)), fx=4, fy=4)
            total, ref_by_type = 0, []
            for k in range(18):
                total += orc.extract_keypoints(heat[:, :, k], ref_by_type, total)
            ref_poses, _ = orc.group_keypoints(ref_by_type, pafs, demo=demo)
            got = postproc.keypoint_lists(kpts_h, counts_h, start_h, b)
            assert np.array_equal(gc.pack_keypoints(got), gc.pack_keypoints(ref_by_type)), b
            gp = np.asarray(postproc.pose_entries_array(poses_h, n_h, b), np.float64).reshape(-1, 20)
            rp = np.asarray(ref_poses, np.float64).reshape(-1, 20)
            assert gp.shape == rp.shape and np.array_equal(gp.view(np.int64), rp.view(np.int64)), b


def test_overflow_is_reported_not_silent(torch_cuda):
    from lwpose_b200 import postproc
    case = [c for c in gc.postproc_cases() if c[0] == "noise_only"][0]
    hm, paf = gc.postproc_maps(case)
    with pytest.raises(postproc.CapacityOverflow):
        _device_postproc(torch_cuda, hm[None], paf[None], True, cap_kpts=8, cap_cand=64)


def test_dropin_functions_match_golden(torch_cuda):
    """The reference-signature wrappers: host arrays in, Python lists / float64 arrays out, input
    heat-map thresholded in place."""
    from lwpose_b200.modules.keypoints import extract_keypoints, group_keypoints
    from oracle import postproc as orc
    g = gc.load("postproc_golden.npz")
    case = [c for c in gc.postproc_cases() if c[0] == "p5n"][0]
    hm, paf = gc.postproc_maps(case)
    heat = orc.resize_cubic(np.ascontiguousarray(hm.transpose(1, 2, 0)), fx=4, fy=4)
    pafs = orc.resize_cubic(np.ascontiguousarray(paf.transpose(1, 2, 0)), fx=4, fy=4)
    before = heat.copy()
    total, by_type = 0, []
    for k in range(18):
        total += extract_keypoints(heat[:, :, k], by_type, total)
    assert np.array_equal(gc.pack_keypoints(by_type), g["pp_p5n_demo_kpts"])
    expect = before[:, :, :18].copy()
    expect[expect < 0.1] = 0
    assert np.array_equal(heat[:, :, :18], expect)  # side effect of the reference (:17)
    x, y, s, i = by_type[0][0]
    assert isinstance(x, np.int64) and isinstance(s, np.float32) and isinstance(i, int)
    for demo, tag in ((True, "demo"), (False, "val")):
        poses, allk = group_keypoints(by_type, pafs, demo=demo)
        assert poses.dtype == np.float64 and allk.dtype == np.float64
        assert np.array_equal(poses.view(np.int64), g["pp_p5n_%s_poses" % tag].view(np.int64))
        assert np.array_equal(allk.view(np.int64), g["pp_p5n_%s_allk" % tag].view(np.int64))
    empty = [[] for _ in range(18)]
    poses, allk = group_keypoints(empty, pafs, demo=True)
    assert poses.shape == (0,) and allk.shape == (0,)


def _fused_postproc(torch, hm, paf, demo, ratio=4, cap_kpts=256, cap_cand=4096, cap_poses=2048, cap_conn=8192):
    from lwpose_b200 import postproc
    n, _, h, w = hm.shape
    heads = torch.zeros((n, h, w, 64), dtype=torch.float32, device="cuda")
    heads[..., :19] = torch.from_numpy(np.ascontiguousarray(hm.transpose(0, 2, 3, 1))).cuda()
    heads[..., 19:57] = torch.from_numpy(np.ascontiguousarray(paf.transpose(0, 2, 3, 1))).cuda()
    kb = postproc.extract_keypoints_fused(heads, ratio, cap_kpts=cap_kpts, cap_candidates=cap_cand)
    poses_d, n_d = postproc.group_keypoints_fused(kb, heads, ratio, demo=demo, cap_poses=cap_poses,
                                                  cap_connections=cap_conn)
    kpts_h, counts_h, start_h, ovf = kb.to_host()
    postproc.raise_on_overflow(ovf)
    return kpts_h, counts_h, start_h, poses_d.cpu().numpy(), n_d.cpu().numpy()


@pytest.mark.parametrize("case", gc.postproc_cases(), ids=lambda c: c[0])
@pytest.mark.parametrize("demo", [True, False])
def test_fused_postproc_bit_exact_vs_reference_golden(torch_cuda, case, demo):
    """No up-sampled map in memory: peaks from smem-rebuilt tiles, PAF samples computed on the fly --
    still the reference's exact key-points and pose entries."""
    from lwpose_b200 import postproc
    g = gc.load("postproc_golden.npz")
    hm, paf = gc.postproc_maps(case)
    kpts_h, counts_h, start_h, poses_h, n_h = _fused_postproc(torch_cuda, hm[None], paf[None], demo)
    by_type = postproc.keypoint_lists(kpts_h, counts_h, start_h, 0)
    tag = "pp_%s_%s" % (case[0], "demo" if demo else "val")
    assert np.array_equal(gc.pack_keypoints(by_type), g[tag + "_kpts"])
    poses = np.asarray(postproc.pose_entries_array(poses_h, n_h, 0), np.float64).reshape(-1, 20)
    assert poses.shape == g[tag + "_poses"].shape
    assert np.array_equal(poses.view(np.int64), g[tag + "_poses"].view(np.int64))


@pytest.mark.parametrize("variant", [{"LWP_NO_PAF_PACK": "1"}, {"LWP_PAF_BLOCKS": "1"}, {"LWP_PAF_BLOCKS": "5"}],
                         ids=["unpacked", "packed_1_block", "packed_5_blocks"])
def test_paf_staging_variants_give_identical_pose_tables(torch_cuda, variant, monkeypatch):
    """The PAF line integral with the limb's channels staged from the re-packed planes (default when the workspace has
    lwp_paf_pack_bytes extra), straight from the head rows (LWP_NO_PAF_PACK=1), and with a (limb, image) spread over
    1 / 5 blocks: the same pose tables, bit for bit, on crowded noisy frames (odd map size: padded plane stride)."""
    from lwpose_b200 import synth
    hm, paf, _ = synth.synthetic_pose_maps(6, 45, 81, seed=11, noise=0.05, max_persons=30)
    ref = _fused_postproc(torch_cuda, hm, paf, True)
    for k, v in variant.items():
        monkeypatch.setenv(k, v)
    got = _fused_postproc(torch_cuda, hm, paf, True)
    for a, b in zip(ref, got):
        assert np.array_equal(np.asarray(a).view(np.uint8), np.asarray(b).view(np.uint8))
    assert int(ref[4].sum()) > 20


@pytest.mark.parametrize("demo", [True, False])
def test_config3_batch256_bit_exact_vs_oracle(torch_cuda, demo):
    """BASELINE.json configs[2] at full size: 256 synthetic frames of 19 heat-maps / 38 PAFs at 46x82 with 1..30
    persons (+ noise), up-sampled x4, through the fused batched path -- key-points and pose entries of EVERY frame
    bit-identical to the oracle."""
    from lwpose_b200 import postproc, synth
    from oracle import postproc as orc
    hm, paf, persons = synth.synthetic_pose_maps(256, 46, 82, seed=0, noise=0.02, max_persons=30)
    assert min(persons) == 1 and max(persons) == 30
    kpts_h, counts_h, start_h, poses_h, n_h = _fused_postproc(torch_cuda, hm, paf, demo, cap_kpts=256, cap_cand=8192,
                                                              cap_poses=512, cap_conn=8192)
    n_poses = 0
    for b in range(hm.shape[0]):
        heat = orc.resize_cubic(np.ascontiguousarray(hm[b].transpose(1, 2, 0)), fx=4, fy=4)
        pafs = orc.resize_cubic(np.ascontiguousarray(paf[b].transpose(1, 2, 0)), fx=4, fy=4)
        total, ref_by_type = 0, []
        for k in range(18):
            total += orc.extract_keypoints(heat[:, :, k], ref_by_type, total)
        ref_poses, _ = orc.group_keypoints(ref_by_type, pafs, demo=demo)
        got = postproc.keypoint_lists(kpts_h, counts_h, start_h, b)
        assert np.array_equal(gc.pack_keypoints(got), gc.pack_keypoints(ref_by_type)), b
        gp = np.asarray(postproc.pose_entries_array(poses_h, n_h, b), np.float64).reshape(-1, 20)
        rp = np.asarray(ref_poses, np.float64).reshape(-1, 20)
        assert gp.shape == rp.shape and np.array_equal(gp.view(np.int64), rp.view(np.int64)), b
        n_poses += gp.shape[0]
    assert n_poses >= sum(persons) // 2   # the synthetic persons are actually found


@pytest.mark.parametrize("shape,ratio", [((3, 13, 21), 4), ((2, 46, 82), 4), ((2, 9, 40), 8), ((1, 33, 7), 3)])
def test_fused_equals_materialised_on_noise(torch_cuda, shape, ratio):
    """Random maps with odd sizes / other ratios: fused and materialising paths give identical tables."""
    torch = torch_cuda
    from lwpose_b200 import postproc
    n, h, w = shape
    rng = np.random.default_rng(h * 100 + w)
    hm = (rng.standard_normal((n, 19, h, w)) * 0.12).astype(np.float32)
    paf = (rng.standard_normal((n, 38, h, w)) * 0.3).astype(np.float32)
    a = _fused_postproc(torch, hm, paf, False, ratio=ratio, cap_kpts=1024, cap_cand=8192)
    hm_d = torch.from_numpy(np.ascontiguousarray(hm.transpose(0, 2, 3, 1))).cuda()
    paf_d = torch.from_numpy(np.ascontiguousarray(paf.transpose(0, 2, 3, 1))).cuda()
    heat = postproc.upsample_cubic(hm_d, fx=ratio, fy=ratio)
    pafs = postproc.upsample_cubic(paf_d, fx=ratio, fy=ratio)
    kb = postproc.extract_keypoints_batched(heat, cap_kpts=1024, cap_candidates=8192)
    poses_d, n_d = postproc.group_keypoints_batched(kb, pafs, demo=False, cap_poses=2048, cap_connections=8192)
    kpts_h, counts_h, start_h, ovf = kb.to_host()
    postproc.raise_on_overflow(ovf)
    assert np.array_equal(a[1], counts_h) and np.array_equal(a[2], start_h)
    for b in range(n):
        assert np.array_equal(gc.pack_keypoints(postproc.keypoint_lists(a[0], a[1], a[2], b)),
                              gc.pack_keypoints(postproc.keypoint_lists(kpts_h, counts_h, start_h, b)))
    assert np.array_equal(a[4], n_d.cpu().numpy())
    ph = poses_d.cpu().numpy()
    for b in range(n):
        k = int(a[4][b])
        assert np.array_equal(a[3][b, :k].view(np.int64), ph[b, :k].view(np.int64))


def test_pose_convert_matches_reference_golden():
    """lwp_pose_convert (demo.py:101-115 + Pose.get_bbox on the device) against fixtures made by the real reference, through
    the batched device API and through PosePipeline(convert=...).poses()."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import golden_cases as gc
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import postproc
    g = gc.load("pose_golden.npz")
    mg = gc._mg()
    for name, pad, scale in mg.POSE_CASES:
        case = [c for c in gc.postproc_cases() if c[0] == name][0]
        hm, paf = gc.postproc_maps(case)
        heads = torch.zeros((1, hm.shape[1], hm.shape[2], 64), dtype=torch.float32, device="cuda")
        heads[0, :, :, :19] = torch.from_numpy(hm.transpose(1, 2, 0)).cuda()
        heads[0, :, :, 19:57] = torch.from_numpy(paf.transpose(1, 2, 0)).cuda()
        kb = postproc.extract_keypoints_fused(heads, 4)
        poses_d, n_d = postproc.group_keypoints_fused(kb, heads, 4, demo=True)
        pk, bb, conf = postproc.pose_convert(poses_d, n_d, kb, stride=8, upsample_ratio=4, pad=pad, scale=scale)
        n = int(n_d[0])
        assert n == g["pose_%s_kpts" % name].shape[0]
        assert np.array_equal(pk[0, :n].cpu().numpy(), g["pose_%s_kpts" % name])
        assert np.array_equal(bb[0, :n].cpu().numpy(), g["pose_%s_bbox" % name])
        assert np.array_equal(conf[0, :n].cpu().numpy().view(np.int64), g["pose_%s_conf" % name].view(np.int64))


def test_resize_pad_u8_bit_exact_vs_cv2_generic_golden():
    """lwp_resize_pad_u8 (row f1: the camera frame's cubic resize + centred pad on the GPU) against cv2 golden hashes (generic
    path) on 720x1280 -> 256x455 (+ pad to 256x456) and 480x640 -> 368x491 (+ pad to 368x496), and against the oracle with pad."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import golden_cases as gc
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import postproc
    from oracle import postproc as orc
    g = gc.load("u8_golden.npz")
    mg = gc._mg()
    for i, (h, w, num, den) in enumerate(mg.U8_CASES):
        img = mg.u8_input(i)
        d = torch.from_numpy(np.stack([img, img[::-1].copy()])).cuda()
        out = postproc.resize_pad_u8(d, fx=num / den, fy=num / den).cpu().numpy()
        assert gc.sha(out[0]) == str(g["u8_%d_sha" % i]), i
        assert np.array_equal(out[1], orc.resize_pad_u8(img[::-1].copy(), fx=num / den, fy=num / den))
    for (h, w, hn) in ((720, 1280, 256), (480, 640, 368)):
        scale, (H, W), (Hp, Wp), pad = postproc.infer_fast_geometry(h, w, hn)
        assert (Hp, Wp) == ((256, 456) if hn == 256 else (368, 496))
        img = np.random.default_rng(h).integers(0, 256, (h, w, 3), dtype=np.uint8)
        got = postproc.resize_pad_u8(torch.from_numpy(img[None]).cuda(), fx=scale, fy=scale, padded=(Hp, Wp), top=pad[0], left=pad[1])
        want = orc.resize_pad_u8(img, fx=scale, fy=scale, padded=(Hp, Wp), top=pad[0], left=pad[1])
        assert np.array_equal(got[0].cpu().numpy(), want)
