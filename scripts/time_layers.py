"""Device time of selected layers of one bench step (CUDA events on the launching stream, each layer alone,
inputs of the step resident).  Usage: python scripts/time_layers.py [--reps R] [--precision bf16|tf32] [--u8] [substr ...]
(a layer is timed when any substring matches its op name; no substrings = all layers)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    args = sys.argv[1:]
    reps, precision, subs, u8 = 20, "bf16", [], False
    while args:
        a = args.pop(0)
        if a == "--reps":
            reps = int(args.pop(0))
        elif a == "--precision":
            precision = args.pop(0)
        elif a == "--u8":      # uint8 NHWC frames (val.normalize fused into the stem)
            u8 = True
        else:
            subs.append(a)
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import synth
    from lwpose_b200.pipeline import PosePipeline
    net = bench.make_net().cuda()
    pipe = PosePipeline(net, 64, bench.HEIGHT, bench.WIDTH, precision=precision, demo=True,
                        **({"input_format": "u8_nhwc"} if u8 else {}))
    if u8:
        x = torch.from_numpy(synth.synthetic_frames(64, bench.HEIGHT, bench.WIDTH, seed=1)).cuda()
    else:
        x = synth.synthetic_net_input(64, bench.HEIGHT, bench.WIDTH, seed=1).cuda()
    for _ in range(2):
        pipe.run_device(x)
    torch.cuda.synchronize()
    plan = pipe.chunks[0].plan
    out = {}
    for i, nm in enumerate(plan.op_names[:plan.num_compute_ops]):
        if subs and not any(s in nm for s in subs):
            continue
        plan.run(x, i, i + 1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            plan.run(x, i, i + 1)
        b.record()
        torch.cuda.synchronize()
        out[nm] = round(a.elapsed_time(b) / reps * 1000.0, 1)
    if pipe.error_flag() != 0:
        out["ERROR_FLAG"] = pipe.error_flag()
    tag = {k: v for k, v in os.environ.items() if k.startswith("LWP_")}
    print(json.dumps({"env": tag, "us": out, "total_us": round(sum(v for v in out.values() if isinstance(v, float)), 1)}))


if __name__ == "__main__":
    main()
