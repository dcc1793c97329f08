"""Run selected layers / post-processing stages of one bench step once, between cudaProfilerStart/Stop,
for `ncu --profile-from-start off`.  Usage: python scripts/prof_layers.py [op names or 'postproc' ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    names = sys.argv[1:] or ["model.0", "model.1.pw", "model.8.dw", "model.8.pw", "initial_stage.trunk.0",
                             "initial_stage.heads.fused", "refinement_stages.0.trunk.0.trunk.1", "postproc"]
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import synth
    from lwpose_b200.pipeline import PosePipeline
    net = bench.make_net().cuda()
    inject_h, _ = bench.person_maps(64, 0, 30)
    inject = torch.from_numpy(inject_h).cuda()
    pipe = PosePipeline(net, 64, bench.HEIGHT, bench.WIDTH, precision="bf16", demo=True,
                        heads_hook=lambda heads, lo: heads.add_(inject[lo:lo + heads.shape[0]]))
    x = synth.synthetic_net_input(64, bench.HEIGHT, bench.WIDTH, seed=1).cuda()
    for _ in range(2):
        pipe.run_device(x)
    torch.cuda.synchronize()
    c = pipe.chunks[0]
    plan = c.plan
    layer_names = [nm for nm in names if nm != "postproc"]
    torch.cuda.profiler.start()
    for nm in names:
        if nm == "postproc":
            if layer_names:   # the single-layer runs above overwrote the heads: restore a full pass with the person maps
                torch.cuda.profiler.stop()
                pipe.run_device(x)
                torch.cuda.synchronize()
                torch.cuda.profiler.start()
            c.enqueue_postproc(c.heads)
        else:
            i = plan.op_names.index(nm)
            plan.run(x, i, i + 1)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled:", names)


if __name__ == "__main__":
    main()
