"""__graft_entry__.smoke(): one small pass of the hot path on cuda:0, checked against the oracle
(the oracle is only the checker here)."""
import numpy as np


def run(verbose=True):
    import torch
    from . import _lib, postproc, synth
    from .models.with_mobilenet import PoseEstimationWithMobileNet
    from .pipeline import PosePipeline
    from oracle import net as onet
    from oracle import postproc as orc

    _lib.require_cuda()
    torch.cuda.set_device(0)
    torch.manual_seed(0)
    net = PoseEstimationWithMobileNet(num_refinement_stages=1).eval()
    synth.randomize_bn_(net, seed=7)
    x = synth.synthetic_net_input(2, 64, 96, seed=3)
    ref = onet.forward(net.state_dict(), x)
    net = net.cuda()
    for precision, tol in (("tf32", 1e-3), ("bf16", 3e-3)):
        net.precision = precision
        outs = net(x.cuda())
        torch.cuda.synchronize()
        assert net.engine().plan(precision, 2, 64, 96).error_flag() == 0, "GEMM pipeline wait timed out"
        err = max(float((o.cpu() - r).abs().max()) for o, r in zip(outs, ref))
        assert err < tol, "network %s max abs err %g >= %g" % (precision, err, tol)
        if verbose:
            print("smoke: network %s max abs err vs oracle %.3g (tol %g)" % (precision, err, tol))

    # one batch on the benchmark's 46x82 grid (12 frames x 368x656, bf16): large enough for the plan to pick the
    # production kernels -- CTA-pair 1x1 GEMMs (conv_gemm2_kernel), the 3x3 strip kernel (conv3x3_pair_kernel), fused
    # heads, TMA depthwise -- so that they appear in the driver's launch list; two frames checked against the oracle
    xb = synth.synthetic_net_input(12, 368, 656, seed=4)
    net.precision = "bf16"
    outs = net(xb.cuda())
    torch.cuda.synchronize()
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    err = 0.0
    for b in (0, 11):
        refb = onet.forward(sd, xb[b:b + 1])
        err = max(err, max(float((o[b:b + 1].cpu() - r).abs().max()) for o, r in zip(outs, refb)))
    assert err < 3e-3, "network bf16 @12x368x656 max abs err %g >= 3e-3" % err
    if verbose:
        print("smoke: network bf16 12x3x368x656 max abs err vs oracle %.3g (tol 3e-3)" % err)

    # post-processing: synthetic 2- and 3-person maps through the fused pipeline stages, bit-exact vs oracle
    hm, paf, _ = synth.synthetic_pose_maps(2, 32, 57, seed=5, noise=0.02, persons=None, max_persons=3)
    heads = torch.zeros((2, 32, 57, 64), dtype=torch.float32, device="cuda")
    heads[..., :19] = torch.from_numpy(hm.transpose(0, 2, 3, 1)).cuda()
    heads[..., 19:57] = torch.from_numpy(paf.transpose(0, 2, 3, 1)).cuda()
    heat = postproc.upsample_cubic(heads, channels=19, fx=4, fy=4)
    pafs = postproc.upsample_cubic(heads, channels=38, fx=4, fy=4, channel_offset=19)
    kb = postproc.extract_keypoints_batched(heat)
    poses_d, n_d = postproc.group_keypoints_batched(kb, pafs, demo=True)
    kpts_h, counts_h, start_h, ovf = kb.to_host()
    postproc.raise_on_overflow(ovf)
    poses_h, n_h = poses_d.cpu().numpy(), n_d.cpu().numpy()
    for b in range(2):
        oh = orc.resize_cubic(np.ascontiguousarray(hm[b].transpose(1, 2, 0)), fx=4, fy=4)
        op = orc.resize_cubic(np.ascontiguousarray(paf[b].transpose(1, 2, 0)), fx=4, fy=4)
        assert np.array_equal(heat[b].cpu().numpy().view(np.int32), oh.view(np.int32)), "upsample bits differ"
        total, by_type = 0, []
        for k in range(18):
            total += orc.extract_keypoints(oh[:, :, k], by_type, total)
        ref_poses, _ = orc.group_keypoints(by_type, op, demo=True)
        got = postproc.keypoint_lists(kpts_h, counts_h, start_h, b)
        assert [[(int(a), int(c), float(s), i) for a, c, s, i in l] for l in got] == \
               [[(int(a), int(c), float(s), i) for a, c, s, i in l] for l in by_type], "key-points differ"
        gp = np.asarray(postproc.pose_entries_array(poses_h, n_h, b), np.float64).reshape(-1, 20)
        rp = np.asarray(ref_poses, np.float64).reshape(-1, 20)
        assert gp.shape == rp.shape and np.array_equal(gp.view(np.int64), rp.view(np.int64)), "poses differ"
        if verbose:
            print("smoke: frame %d: %d key-points, %d poses, bit-exact vs oracle" % (b, total, len(rp)))

    # the fused pipeline object end to end (network + post-processing, host in / host out)
    pipe = PosePipeline(net, batch=2, height=64, width=96, precision="bf16")
    res = pipe(x.pin_memory()).check()
    assert pipe.error_flag() == 0
    if verbose:
        print("smoke: pipeline ran, %d poses (random-init network)" % res.total_poses())
