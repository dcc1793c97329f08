"""CPU-side tests: the C-ABI library loads and exports every symbol the header declares, the module
mirror has the reference's state_dict layout, host helpers behave like the reference's, and the
data-parallel plumbing works over gloo with world_size 2.  No GPU compute is called here."""
import os
import re
import socket

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "lwpose_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lwp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import _lib
    L = _lib.load()
    declared = _header_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), "library does not export %s" % name
        assert name in _lib.SIGNATURES, "ctypes binding misses %s" % name
    assert sorted(_lib.SIGNATURES) == declared
    assert L.lwp_version() >= 100


def test_no_cpu_fallback():
    import torch
    import lwpose_b200  # noqa: F401
    from lwpose_b200.models.with_mobilenet import PoseEstimationWithMobileNet
    from lwpose_b200.modules.keypoints import extract_keypoints
    net = PoseEstimationWithMobileNet(1).eval()
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 64, 64))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            extract_keypoints(np.zeros((8, 8), np.float32), [], 0)


def test_state_dict_layout_and_seeded_init_match_reference_golden():
    import torch
    import golden_cases as gc
    import lwpose_b200  # noqa: F401
    from lwpose_b200.models.with_mobilenet import PoseEstimationWithMobileNet
    g = gc.load("net_golden.npz")
    for name, R in (("r1_64x96", 1), ("r2_48x72", 2)):
        torch.manual_seed(0)
        net = PoseEstimationWithMobileNet(R)
        sd = net.state_dict()
        assert list(sd.keys()) == [str(k) for k in g["net_%s_keys" % name]]
        assert gc.sha(np.concatenate([v.numpy().astype(np.float64).ravel() for v in sd.values()])) == \
            str(g["net_%s_init_sha" % name])
    assert hasattr(net, "model") and hasattr(net, "cpm") and hasattr(net, "initial_stage")
    assert len(net.refinement_stages) == 2


def test_load_state_is_tolerant_like_the_reference(capsys):
    import torch
    import lwpose_b200  # noqa: F401
    from lwpose_b200.models.with_mobilenet import PoseEstimationWithMobileNet
    from lwpose_b200.modules.load_state import load_state
    torch.manual_seed(1)
    src = PoseEstimationWithMobileNet(1)
    ckpt = {"state_dict": {k: v.clone() for k, v in src.state_dict().items()}}
    del ckpt["state_dict"]["cpm.align.0.bias"]
    ckpt["state_dict"]["model.0.0.weight"] = torch.zeros(1)
    torch.manual_seed(2)
    dst = PoseEstimationWithMobileNet(1)
    before = dst.state_dict()["cpm.align.0.bias"].clone()
    load_state(dst, ckpt)
    out = capsys.readouterr().out
    assert out.count("[WARNING] Not found pre-trained parameters") == 2
    sd = dst.state_dict()
    assert torch.equal(sd["cpm.align.0.bias"], before)
    assert torch.equal(sd["model.3.0.weight"], src.state_dict()["model.3.0.weight"])


def test_pad_width_and_normalize_semantics():
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import val
    img = np.full((256, 455, 3), 200, np.uint8)
    n = val.normalize(img, (128, 128, 128), 1 / 256)
    assert n.dtype == np.float64 and np.allclose(n, (200 - 128) / 256)
    dims = [256, 455]
    padded, pad = val.pad_width(n, 8, (0, 0, 0), dims)
    assert padded.shape[:2] == (256, 456) and pad == [0, 0, 0, 1] and dims == [256, 456]
    padded, pad = val.pad_width(np.zeros((184, 100, 3)), 8, (0, 0, 0), [368, 368])
    assert padded.shape[:2] == (368, 368) and pad == [92, 134, 92, 134]


def test_tables_and_synthetic_generator_are_deterministic():
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import synth
    from lwpose_b200.modules import keypoints as kp
    assert len(kp.BODY_PARTS_KPT_IDS) == 19 and len(kp.BODY_PARTS_PAF_IDS) == 19
    assert sorted(c for pair in kp.BODY_PARTS_PAF_IDS for c in pair) == list(range(38))
    a = synth.synthetic_pose_maps(3, 16, 20, seed=4)
    b = synth.synthetic_pose_maps(3, 16, 20, seed=4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == [1, 2, 3]


def test_shard_range_partitions_exactly():
    import lwpose_b200  # noqa: F401
    from lwpose_b200.parallel import shard_range
    for n in (0, 1, 7, 64, 255, 256):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import parallel
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    lo, hi = parallel.shard_range(6, rank, world)
    n_poses = torch.arange(lo, hi, dtype=torch.int32)
    poses = torch.full((hi - lo, 4, 20), float(rank), dtype=torch.float64)
    gn, gp = parallel.gather_pose_tables(n_poses, poses)
    slow = parallel.max_over_ranks(1.0 + rank)
    # unequal shards (7 frames over 2 ranks: 4 + 3): padded for the collective, trimmed afterwards
    lo, hi = parallel.shard_range(7, rank, world)
    un, up = parallel.gather_pose_tables(torch.arange(lo, hi, dtype=torch.int32),
                                         torch.full((hi - lo, 4, 20), float(rank), dtype=torch.float64), total=7)
    q.put((rank, gn.tolist(), gp[:, 0, 0].tolist(), slow, un.tolist(), up[:, 0, 0].tolist()))
    dist.destroy_process_group()


def test_gather_and_timing_over_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, gn, gp, slow, un, up in outs:
        assert gn == [0, 1, 2, 3, 4, 5]
        assert gp == [0.0, 0.0, 0.0, 1.0, 1.0, 1.0]
        assert slow == 2.0
        assert un == [0, 1, 2, 3, 4, 5, 6] and up == [0.0] * 4 + [1.0] * 3


def test_debug_switches_are_refused(monkeypatch):
    """A run with a work-skipping LWP_DEBUG_* switch in the environment is refused (release builds do not even
    contain those code paths: lwp_timing_experiments() == 0)."""
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import _lib
    L = _lib.load()
    assert L.lwp_timing_experiments() == 0
    _lib.refuse_debug_env()
    monkeypatch.setenv("LWP_DEBUG_GEMM", "4")
    with pytest.raises(_lib.LwpError):
        _lib.refuse_debug_env()
    monkeypatch.setenv("LWP_DEBUG_GEMM", "0")
    _lib.refuse_debug_env()


def test_net_load_rejects_garbage_without_a_gpu():
    """lwp_net_load checks the blob header before it touches CUDA."""
    import ctypes
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import _lib
    L = _lib.load()
    h = _lib._c_void_p()
    buf = ctypes.create_string_buffer(b"not a blob at all.." * 4)
    assert L.lwp_net_load(ctypes.cast(buf, ctypes.c_void_p), 64, h) == 1   # LWP_EINVAL
    assert b"LWPB" in L.lwp_last_error()
