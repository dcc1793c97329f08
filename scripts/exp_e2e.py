"""Where the gap between the device-resident loop and the host-to-host loop goes: the same 64-frame step timed as
(a) run_device, float32 frames on the device; (b) run_device, uint8 frames on the device; (c) submit/collect, uint8
frames on the device (no H2D); (d) submit/collect, pinned uint8 host frames (the bench's `e2e`)."""
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
    depth = int(os.environ.get("DEPTH", "2"))
    pf = PosePipeline(net, 64, bench.HEIGHT, bench.WIDTH, precision="bf16", demo=True, heads_hook=hook, depth=depth)
    p8 = PosePipeline(net, 64, bench.HEIGHT, bench.WIDTH, precision="bf16", demo=True, heads_hook=hook, input_format="u8_nhwc",
                      depth=depth)
    xf = synth.synthetic_net_input(64, bench.HEIGHT, bench.WIDTH, seed=1).cuda()
    x8h = torch.from_numpy(synth.synthetic_frames(64, bench.HEIGHT, bench.WIDTH, seed=1)).pin_memory()
    x8 = x8h.cuda()

    def dev_loop(p, x):
        def run(k):
            for _ in range(k):
                p.run_device(x)
            p.join()
        return run

    def stream_loop(p, x):
        return lambda k: bench.stream_steps(p, x, k)
    for name, fn in (("a run_device f32", dev_loop(pf, xf)), ("b run_device u8", dev_loop(p8, x8)),
                     ("c submit/collect u8 on device", stream_loop(p8, x8)), ("d submit/collect u8 pinned host", stream_loop(p8, x8h))):
        fn(3)
        print(name, "%.3f ms" % min(timed(fn) for _ in range(3)))


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


def per_op():
    """Per-op event times of one pass of the float32-input and the uint8-input plan (ops back to back, no post-processing)."""
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import synth
    from lwpose_b200.pipeline import PosePipeline
    net = bench.make_net().cuda()
    pf = PosePipeline(net, 64, bench.HEIGHT, bench.WIDTH, precision="bf16", demo=True)
    p8 = PosePipeline(net, 64, bench.HEIGHT, bench.WIDTH, precision="bf16", demo=True, input_format="u8_nhwc")
    xf = synth.synthetic_net_input(64, bench.HEIGHT, bench.WIDTH, seed=1).cuda()
    x8 = torch.from_numpy(synth.synthetic_frames(64, bench.HEIGHT, bench.WIDTH, seed=1)).cuda()
    for name, p, x in (("f32", pf, xf), ("u8", p8, x8)):
        plan = p.chunks[0].plan
        n = plan.num_compute_ops
        acc = [[] for _ in range(n)]
        for _ in range(7):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
            ev[0].record()
            for i in range(n):
                plan.run(x, i, i + 1)
                ev[i + 1].record()
            torch.cuda.synchronize()
            for i in range(n):
                acc[i].append(ev[i].elapsed_time(ev[i + 1]))
        med = [sorted(a)[len(a) // 2] for a in acc]
        print(name, "sum %.3f ms" % sum(med), "ops:", [(plan.op_names[i], round(med[i] * 1000)) for i in range(n)])
        t = []
        for _ in range(7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); plan.run(x, 0, n); b.record(); torch.cuda.synchronize()
            t.append(a.elapsed_time(b))
        print(name, "network in one call %.3f ms" % sorted(t)[3])


if len(sys.argv) > 1 and sys.argv[1] == "ops":
    per_op()
