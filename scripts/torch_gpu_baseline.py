"""Context number only (not part of the product, not used by bench.py): the same network run through stock
PyTorch / cuDNN kernels on the GPU (channels_last, bf16 autocast and fp32-TF32), batch 64 @368x656."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def torch_forward(net, x):
    f = net.model(x)
    a = net.cpm.align(f)
    f = net.cpm.conv(a + net.cpm.trunk(a))
    t = net.initial_stage.trunk(f)
    outs = [net.initial_stage.heatmaps(t), net.initial_stage.pafs(t)]
    for st in net.refinement_stages:
        t = torch.cat([f, outs[-2], outs[-1]], 1)
        for blk in st.trunk:
            i = blk.initial(t)
            t = i + blk.trunk(i)
        outs += [st.heatmaps(t), st.pafs(t)]
    return outs


def main():
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import synth
    net = bench.make_net().cuda().to(memory_format=torch.channels_last)
    x = synth.synthetic_net_input(64, bench.HEIGHT, bench.WIDTH, seed=1).cuda().contiguous(memory_format=torch.channels_last)
    torch.backends.cudnn.benchmark = True
    for name, ctx in (("bf16 autocast", torch.autocast("cuda", dtype=torch.bfloat16)),
                      ("fp32 (TF32 allowed)", torch.autocast("cuda", enabled=False))):
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        with torch.no_grad(), ctx:
            for _ in range(5):
                torch_forward(net, x)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                torch_forward(net, x)
            b.record()
            torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print("torch/cuDNN eager %s: %.3f ms per 64 frames (network only) = %.0f frames/s" % (name, ms, 64 / ms * 1000))


if __name__ == "__main__":
    main()
