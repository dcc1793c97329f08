"""Two-GPU data-parallel check (skipped on a 1-GPU box): the frames of one batch sharded over two ranks
(shard_range, unequal shards), each rank running its own PosePipeline, gathered over NCCL -- must give bit for bit the
pose tables ONE pipeline produces for the whole batch.  Run with `gpurun --gpus 2`."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, H, W = 5, 64, 96   # 5 frames over 2 ranks: shards of 3 and 2


def _inputs():
    from lwpose_b200 import synth
    hm, paf, _ = synth.synthetic_pose_maps(B, H // 8, W // 8, seed=31, max_persons=3)
    inj = np.zeros((B, H // 8, W // 8, 64), np.float32)
    inj[..., :19] = hm.transpose(0, 2, 3, 1)
    inj[..., 19:57] = paf.transpose(0, 2, 3, 1)
    return synth.synthetic_net_input(B, H, W, seed=2), inj


def _net():
    import torch
    from lwpose_b200 import synth
    from lwpose_b200.models.with_mobilenet import PoseEstimationWithMobileNet
    torch.manual_seed(0)
    net = PoseEstimationWithMobileNet(1).eval()
    synth.randomize_bn_(net, seed=7)
    return net.cuda()


def _run(x, inj):
    """(n_poses [b] int32, pose_entries [b, cap, 20] float64) device tensors of one pipeline over frames x."""
    import torch
    from lwpose_b200.pipeline import PosePipeline
    inj_d = torch.from_numpy(inj).cuda()
    pipe = PosePipeline(_net(), x.shape[0], H, W, precision="bf16", heads_hook=lambda t, o: t.add_(inj_d[o:o + t.shape[0]]))
    pipe.run_device(x.cuda())
    pipe.join()
    torch.cuda.synchronize()
    assert pipe.error_flag() == 0
    return pipe.n_poses.clone(), pipe.pose_entries.clone()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import parallel
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    x, inj = _inputs()
    lo, hi = parallel.shard_range(B, rank, world)
    n_d, p_d = _run(x[lo:hi], inj[lo:hi])
    gn, gp = parallel.gather_pose_tables(n_d, p_d, total=B)
    torch.cuda.synchronize()
    q.put((rank, gn.cpu().numpy(), gp.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_single_rank():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import lwpose_b200  # noqa: F401
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted((q.get(timeout=300) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the whole batch on ONE rank, in this process
    x, inj = _inputs()
    n1, p1 = _run(x, inj)
    n1, p1 = n1.cpu().numpy(), p1.cpu().numpy()
    assert n1.shape == (B,) and int(n1.sum()) >= B   # the injected persons were found
    for rank, gn, gp in outs:
        assert gn.shape == (B,) and np.array_equal(gn, n1), (rank, gn, n1)
        for b in range(B):
            k = int(n1[b])
            assert np.array_equal(gp[b, :k].view(np.int64), p1[b, :k].view(np.int64)), (rank, b)
