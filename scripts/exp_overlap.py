"""How much of the post-processing is hidden behind the next batch's network: the 64-frame step timed as
(a) network only, (b) network + post-processing on its own stream (the product), (c) both on one stream,
(d) post-processing only (heads resident)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def timed(fn, steps=20):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn(steps)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import synth
    from lwpose_b200.pipeline import PosePipeline
    net = bench.make_net().cuda()
    inject_h, _ = bench.person_maps(64, 0, 30)
    inject = torch.from_numpy(inject_h).cuda()
    hook = lambda heads, lo: heads.add_(inject[lo:lo + heads.shape[0]])  # noqa: E731
    x = synth.synthetic_net_input(64, bench.HEIGHT, bench.WIDTH, seed=1).cuda()
    po = PosePipeline(net, 64, bench.HEIGHT, bench.WIDTH, precision="bf16", demo=True, heads_hook=hook)
    ps = PosePipeline(net, 64, bench.HEIGHT, bench.WIDTH, precision="bf16", demo=True, heads_hook=hook, overlap_postproc=False)
    plan = po.chunks[0].plan
    n = plan.num_compute_ops

    def net_only(k):
        for _ in range(k):
            plan.run(x, 0, n)

    def both(p):
        def run(k):
            for _ in range(k):
                p.run_device(x)
            p.join()
        return run

    def post_only(k):
        c = po.chunks[0]
        for _ in range(k):
            c.enqueue_postproc(c.heads)
    for name, fn in (("a network only", net_only), ("b product (post-processing on its own stream)", both(po)),
                     ("c one stream", both(ps)), ("d post-processing only", post_only)):
        fn(3)
        print(name, "%.3f ms" % min(timed(fn) for _ in range(3)))


if __name__ == "__main__":
    main()
