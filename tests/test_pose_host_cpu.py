"""Row f3 host logic (mirror of the reference's modules/pose.py and modules/one_euro_filter.py) against golden
vectors produced by the REAL reference (tests/golden/make_golden.py: gen_pose), and the oracle's pose_convert
restatement of demo.py:101-115 against the same fixtures."""
import numpy as np

import golden_cases as gc


def _pose_mod():
    import lwpose_b200  # noqa: F401
    from lwpose_b200.modules import pose
    return pose


def test_one_euro_filter_trace_matches_reference():
    import lwpose_b200  # noqa: F401
    from lwpose_b200.modules.one_euro_filter import OneEuroFilter
    g = gc.load("pose_golden.npz")
    f = OneEuroFilter(freq=15, beta=0.1)
    got = np.asarray([f(v + (-1) ** (v % 2)) for v in range(12)], np.float64)
    assert np.array_equal(got.view(np.int64), g["one_euro_trace"].view(np.int64))


def test_oracle_pose_convert_matches_reference_golden():
    from oracle import postproc as orc
    g = gc.load("pose_golden.npz")
    pp = gc.load("postproc_golden.npz")
    mg = gc._mg()
    for name, pad, scale in mg.POSE_CASES:
        kp, bbox, conf = orc.pose_convert(pp["pp_%s_demo_poses" % name], pp["pp_%s_demo_allk" % name], 8, 4, pad, scale)
        assert np.array_equal(kp, g["pose_%s_kpts" % name])
        assert np.array_equal(bbox, g["pose_%s_bbox" % name])
        assert np.array_equal(conf.view(np.int64), g["pose_%s_conf" % name].view(np.int64))


def test_pose_bbox_and_tracking_match_reference_golden():
    pose = _pose_mod()
    g = gc.load("pose_golden.npz")
    for smooth in (0, 1):
        pose.Pose.last_id = -1
        prev = []
        for f in range(3):
            kin, conf = g["track_s%d_f%d_in" % (smooth, f)], g["track_s%d_f%d_conf" % (smooth, f)]
            cur = [pose.Pose(kin[i].copy(), conf[i]) for i in range(kin.shape[0])]
            pose.track_poses(prev, cur, smooth=bool(smooth))
            assert [p.id for p in cur] == g["track_s%d_f%d_ids" % (smooth, f)].tolist()
            assert np.array_equal(np.stack([p.keypoints for p in cur]), g["track_s%d_f%d_out" % (smooth, f)])
            assert np.array_equal(np.asarray([p.bbox for p in cur], np.int32), g["track_s%d_f%d_bbox" % (smooth, f)])
            prev = cur
    sim = np.asarray([[pose.get_similarity(a, b) for b in prev] for a in prev], np.int64)
    assert np.array_equal(sim, g["track_similarity"])


def test_pose_class_surface():
    pose = _pose_mod()
    assert pose.Pose.num_kpts == 18 and len(pose.Pose.kpt_names) == 18 and pose.Pose.vars.dtype == np.float32
    k = -np.ones((18, 2), np.int32)
    assert pose.Pose.get_bbox(k) == (0, 0, 0, 0)
    k[3] = (10, 20)
    k[7] = (4, 25)
    p = pose.Pose(k, 1.5)
    assert p.bbox == (4, 20, 7, 6) and p.id is None and len(p.filters) == 18
    pose.Pose.last_id = 4
    p.update_id()
    assert p.id == 5 and pose.Pose.last_id == 5
    p.update_id(2)
    assert p.id == 2


def test_convert_to_coco_format_matches_reference_golden():
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import val
    g = gc.load("pose_golden.npz")
    pp = gc.load("postproc_golden.npz")
    for name in ("p8", "p15n"):
        coco, scores = val.convert_to_coco_format(pp["pp_%s_val_poses" % name], pp["pp_%s_val_allk" % name])
        assert np.array_equal(np.asarray(coco, np.float64), g["coco_%s_kpts" % name])
        assert np.array_equal(np.asarray(scores, np.float64).view(np.int64), g["coco_%s_scores" % name].view(np.int64))
    assert val.convert_to_coco_format(np.asarray([]), np.asarray([])) == ([], [])


def test_val_scale_geometry():
    """Net input sizes and pads of val.infer's four scales for a 480x640 frame."""
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import val
    geo = val.scale_geometry(480, 640, [0.5, 1.0, 1.5, 2.0], 368, 8)
    assert [g[2] for g in geo] == [(368, 368), (368, 496), (552, 736), (736, 984)]
    assert [g[3] for g in geo] == [[92, 61, 92, 62], [0, 2, 0, 3], [0, 0, 0, 0], [0, 1, 0, 2]]
