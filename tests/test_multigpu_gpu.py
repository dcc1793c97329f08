"""Two-GPU data-parallel check (skipped on a 1-GPU box): shards of one batch processed by two ranks and
gathered over NCCL give the same pose tables as one rank processing the whole batch."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import parallel, synth
    from lwpose_b200.models.with_mobilenet import PoseEstimationWithMobileNet
    from lwpose_b200.pipeline import PosePipeline
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    torch.manual_seed(0)
    net = PoseEstimationWithMobileNet(1).eval()
    synth.randomize_bn_(net, seed=7)
    net = net.cuda()
    B, H, W = 4, 64, 96
    hm, paf, _ = synth.synthetic_pose_maps(B, H // 8, W // 8, seed=31, max_persons=2)
    inj = np.zeros((B, H // 8, W // 8, 64), np.float32)
    inj[..., :19] = hm.transpose(0, 2, 3, 1)
    inj[..., 19:57] = paf.transpose(0, 2, 3, 1)
    x = synth.synthetic_net_input(B, H, W, seed=2)
    lo, hi = parallel.shard_range(B, rank, world)
    inj_d = torch.from_numpy(inj[lo:hi]).cuda()
    pipe = PosePipeline(net, hi - lo, H, W, precision="bf16", heads_hook=lambda t, o: t.add_(inj_d[o:o + t.shape[0]]))
    pipe.run_device(x[lo:hi].cuda())
    pipe.join()
    gn, gp = parallel.gather_pose_tables(pipe.n_poses, pipe.pose_entries)
    torch.cuda.synchronize()
    q.put((rank, gn.cpu().tolist(), gp.cpu().numpy().tobytes()))
    dist.destroy_process_group()


def test_two_rank_shards_equal_single_rank():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert outs[0][1] == outs[1][1] and outs[0][2] == outs[1][2]
    assert sum(outs[0][1]) >= 4
