"""Where a bench step's device time goes: network alone, network + hook + head copy, full step with the
post-processing on its own stream / on the network's stream.  python scripts/step_breakdown.py [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def timed(fn, reps):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return round(a.elapsed_time(b) / reps, 4)


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import synth
    from lwpose_b200.pipeline import PosePipeline
    net = bench.make_net().cuda()
    inject_h, _ = bench.person_maps(64, 0, 30)
    inject = torch.from_numpy(inject_h).cuda()
    x = synth.synthetic_net_input(64, bench.HEIGHT, bench.WIDTH, seed=1).cuda()
    out = {}
    for overlap in (True, False):
        pipe = PosePipeline(net, 64, bench.HEIGHT, bench.WIDTH, precision="bf16", demo=True,
                            heads_hook=lambda heads, lo: heads.add_(inject[lo:lo + heads.shape[0]]),
                            overlap_postproc=overlap)
        c = pipe.chunks[0]
        if overlap:
            out["net_only_ms"] = timed(lambda: c.plan.run_compute(x), reps)
            out["postproc_stage_ms"] = bench.postproc_stage_ms(pipe, x, reps)   # on heads with the injected persons

        def step():
            pipe.run_device(x)
        t = timed(step, reps)
        pipe.join()
        out["step_overlap_ms" if overlap else "step_serial_ms"] = t
        del pipe
    print(json.dumps(out))


if __name__ == "__main__":
    main()
