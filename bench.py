"""bench.py -- frames/s of the Lightweight OpenPose inference hot path (network + cubic up-sample +
key-point extraction + PAF grouping) at 368x656, on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|tf32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one batch of 64 synthetic frames per GPU (BASELINE.json
configs[1]: PoseEstimationWithMobileNet, 1 refinement stage, random-init weights, 64x3x368x656).
Random-init weights cannot produce person-like maps, so seeded synthetic person maps (configs[2]
generator, 1..30 persons per frame) are added to the network's head output between the network and
the post-processing; both arms do the same.  Frames are sharded data-parallel (weak scaling, no
collective on the compute path).  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HEIGHT, WIDTH, REFINE = 368, 656, 1
METRIC = "frames/s end-to-end (net+grouping) @368x656"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"])
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU per step")
    ap.add_argument("--max-persons", type=int, default=30)
    ap.add_argument("--cpu-frames", type=int, default=8, help="frames of the workload timed for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-u8", action="store_true", help="skip the uint8-frame end-to-end run")
    ap.add_argument("--no-other-precision", action="store_true", help="skip the device-resident run on the other precision")
    ap.add_argument("--chunk", type=int, default=0, help="frames per pipeline chunk (0 = whole batch)")
    ap.add_argument("--e2e-depth", type=int, default=3, help="batches in flight in the end-to-end (submit / collect) runs")
    ap.add_argument("--no-overlap-postproc", action="store_true", help="post-processing on the network's stream")
    ap.add_argument("--unfused-postproc", action="store_true", help="materialise the up-sampled maps like the reference")
    ap.add_argument("--steady-seconds", type=float, default=2.0, help="length of the extra steady-state run (0 = skip)")
    ap.add_argument("--no-configs", action="store_true", help="skip the config0 / config3 / config4 extra keys")
    ap.add_argument("--config3-batch", type=int, default=256, help="GLOBAL batch of configs[3] (strong scaling over the ranks)")
    ap.add_argument("--config4-frames", type=int, default=64, help="GLOBAL number of 480x640 frames of configs[4] per step (sharded over the ranks)")
    return ap.parse_args()


NET_TOL = {"bf16": 3e-3, "tf32": 1e-3}   # max abs error of the heads vs the CPU fp32 oracle (outputs are O(0.1))


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], bf16=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm=6650.0, bf16=1400.0, source="fallback")


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML in a background thread DURING the timed region."""

    def __init__(self, index, period_s=0.002):
        import threading
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
            return
        self.period = period_s
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for nm, bit in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        if self._thread is None:
            return out
        self._stop.set()
        self._thread.join(timeout=2)
        if self.samples:
            sm = sorted(self.samples)
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                   "samples": len(sm)}
        return out


def make_net(seed=0, refine=REFINE):
    import torch
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import synth
    from lwpose_b200.models.with_mobilenet import PoseEstimationWithMobileNet
    torch.manual_seed(seed)
    net = PoseEstimationWithMobileNet(num_refinement_stages=refine).eval()
    synth.randomize_bn_(net, seed=7)
    return net


def person_maps(batch, rank, max_persons, height=HEIGHT, width=WIDTH):
    """[batch, h/8, w/8, 64] float32: synthetic person heat-maps / PAFs in head-buffer layout."""
    import numpy as np
    from lwpose_b200 import synth
    hm, paf, counts = synth.synthetic_pose_maps(batch, height // 8, width // 8, seed=100 + rank, max_persons=max_persons)
    m = np.zeros((batch, height // 8, width // 8, 64), np.float32)
    m[..., :19] = hm.transpose(0, 2, 3, 1)
    m[..., 19:57] = paf.transpose(0, 2, 3, 1)
    return m, counts


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's CPU path (reference itself is not on the GPU box)
# ----------------------------------------------------------------------------------------------
def cpu_frame(state_dict, x1, inject1, keep=None):
    """One frame through the CPU path: torch-CPU forward, cubic x4, 18 x extract, group (demo=True).
    keep: optional list that receives the frame's raw heads [h, w, 57] (network output before the person maps)."""
    import numpy as np
    from oracle import net as onet
    from oracle import postproc as orc
    outs = onet.forward(state_dict, x1)
    hm0 = outs[-2][0].numpy().transpose(1, 2, 0)
    paf0 = outs[-1][0].numpy().transpose(1, 2, 0)
    if keep is not None:
        keep.append(np.concatenate([hm0, paf0], 2))
    hm = hm0 + inject1[..., :19] if inject1 is not None else hm0
    paf = paf0 + inject1[..., 19:57] if inject1 is not None else paf0
    heat = orc.resize_cubic(np.ascontiguousarray(hm), fx=4, fy=4)
    pafs = orc.resize_cubic(np.ascontiguousarray(paf), fx=4, fy=4)
    total, by_type = 0, []
    for k in range(18):
        total += orc.extract_keypoints(heat[:, :, k], by_type, total)
    poses, _ = orc.group_keypoints(by_type, pafs, demo=True)
    return len(poses)


def cpu_time_frames(state_dict, x, inject, frames, warm=1, keep=None):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    for i in range(warm):
        cpu_frame(state_dict, x[i:i + 1], inject[i])
    t0 = time.perf_counter()
    for i in range(frames):
        cpu_frame(state_dict, x[i % x.shape[0]:i % x.shape[0] + 1], inject[i % inject.shape[0]], keep)
    return time.perf_counter() - t0


def raw_heads(pipe, x_dev, frames):
    """The network's own head output (no person maps) of the first `frames` frames, float32 [frames, h, w, 57] on the host."""
    import torch
    hook, pipe.heads_hook = pipe.heads_hook, None
    try:
        pipe.run_device(x_dev)
        pipe.join()
        torch.cuda.synchronize()
        return pipe.heads[:frames, :, :, :57].cpu().numpy()
    finally:
        pipe.heads_hook = hook


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from lwpose_b200 import synth
    net = make_net()
    sd = net.state_dict()
    frames = max(1, min(args.cpu_frames, 4))
    x = synth.synthetic_net_input(frames, HEIGHT, WIDTH, seed=1)
    inject, _ = person_maps(frames, 0, args.max_persons)
    for _ in range(max(0, min(args.warmup, 2))):
        cpu_time_frames(sd, x, inject, 1, warm=0)
    times = [cpu_time_frames(sd, x, inject, frames, warm=0) for _ in range(args.steps)]
    total = sum(times)
    value = frames * args.steps / total
    cores = torch.get_num_threads()
    sample = "%d frames/step of the %dx3x%dx%d workload, batch-1 loop (torch CPU fp32 + C oracle post-processing)" % (
        frames, args.batch, HEIGHT, WIDTH)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, world):
    return {"workload": "configs[1]: PoseEstimationWithMobileNet R=%d random-init (seed 0, randomised BN), %d frames/GPU "
                        "x 3x%dx%d, + synthetic 1..%d-person maps added to the head output, cubic x4, extract, group "
                        "(demo=True)" % (REFINE, args.batch, HEIGHT, WIDTH, args.max_persons),
            "frames_per_gpu_per_step": args.batch, "global_batch": args.batch * world,
            "parallelism": "dp%d (frames sharded, no collective on the compute path)" % world,
            "chunk": args.chunk,
            "l2": "inputs larger than L2 (%.0f MB of frames per step per GPU)" % (args.batch * 3 * HEIGHT * WIDTH * 4 / 1e6)}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line (NCCL prints its version there)
        dist.init_process_group("nccl", device_id=dev)
    import lwpose_b200  # noqa: F401
    from lwpose_b200 import _lib, parallel, synth
    from lwpose_b200.pipeline import PosePipeline
    _lib.load()
    _lib.refuse_debug_env()   # no timed run with a work-skipping LWP_DEBUG_* switch in the environment

    net = make_net().to(dev)
    inject_h, persons = person_maps(args.batch, rank, args.max_persons)
    inject = torch.from_numpy(inject_h).to(dev)
    pipe = PosePipeline(net, args.batch, HEIGHT, WIDTH, precision=args.precision, demo=True,
                        heads_hook=lambda heads, lo: heads.add_(inject[lo:lo + heads.shape[0]]),
                        fused=not args.unfused_postproc, chunk=args.chunk or None,
                        overlap_postproc=not args.no_overlap_postproc, depth=args.e2e_depth)
    x_host = synth.synthetic_net_input(args.batch, HEIGHT, WIDTH, seed=1 + rank).pin_memory()
    x_dev = x_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput -------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        pipe.run_device(x_dev)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        pipe.run_device(x_dev)
    pipe.join()   # post-processing runs on its own stream: the timed region ends when the last step's grouping ends
    e1.record()
    barrier()
    ms = parallel.max_over_ranks(e0.elapsed_time(e1), device=dev) / args.steps
    clocks = sampler.stop() if sampler else None
    if pipe.error_flag() != 0:
        raise RuntimeError("GEMM pipeline wait timed out (error flag %d)" % pipe.error_flag())
    value = world * args.batch / (ms / 1000.0)

    # ---- steady state: the same loop for >= --steady-seconds (power-cap / clock behaviour of a long run) ----------
    steady = None
    if args.steady_seconds > 0:
        n_steady = max(args.steps, int(args.steady_seconds * 1000.0 / ms) + 1)
        sampler2 = ClockSampler(local_rank, period_s=0.02) if rank == 0 else None
        barrier()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(n_steady):
            pipe.run_device(x_dev)
        pipe.join()
        q1.record()
        barrier()
        st_ms = parallel.max_over_ranks(q0.elapsed_time(q1), device=dev) / n_steady
        steady = {"value": world * args.batch / (st_ms / 1000.0), "unit": "frames/s", "ms_per_step": st_ms,
                  "steps": n_steady, "seconds": st_ms * n_steady / 1000.0,
                  "clocks": sampler2.stop() if sampler2 else None}
    cpu_n = max(1, min(args.cpu_frames, args.batch))
    gpu_heads = {}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gpu_heads[args.precision] = raw_heads(pipe, x_dev, cpu_n)

    # ---- end to end through the public API: pinned host frames in, host pose tables out --------------
    # streaming use of PosePipeline (submit / collect, --e2e-depth batches in flight): every step's H2D copy and
    # result read-back are inside the timed region; the copy of step i+1 overlaps the kernels of step i
    for _ in range(2):
        pipe(x_host)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    res = stream_steps(pipe, x_host, args.steps)
    s1.record()
    s1.synchronize()
    barrier()
    res.check()
    e2e_ms = parallel.max_over_ranks(s0.elapsed_time(s1) / args.steps, device=dev)
    e2e_value = world * args.batch / (e2e_ms / 1000.0)
    total_poses = res.total_poses()
    # the same streaming run fed with raw uint8 BGR frames (normalisation fused into the stem kernel): 4x fewer H2D bytes
    e2e_u8 = None
    if not args.no_u8:
        pipe8 = PosePipeline(net, args.batch, HEIGHT, WIDTH, precision=args.precision, demo=True,
                             heads_hook=lambda heads, lo: heads.add_(inject[lo:lo + heads.shape[0]]),
                             fused=not args.unfused_postproc, chunk=args.chunk or None,
                             overlap_postproc=not args.no_overlap_postproc, input_format="u8_nhwc", depth=args.e2e_depth)
        x8 = torch.from_numpy(synth.synthetic_frames(args.batch, HEIGHT, WIDTH, seed=1 + rank)).pin_memory()
        for _ in range(2):
            pipe8(x8)
        barrier()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        u0.record()
        stream_steps(pipe8, x8, args.steps).check()
        u1.record()
        u1.synchronize()
        barrier()
        u8_ms = parallel.max_over_ranks(u0.elapsed_time(u1) / args.steps, device=dev)
        e2e_u8 = {"value": world * args.batch / (u8_ms / 1000.0), "unit": "frames/s", "ms_per_step": u8_ms,
                  "h2d_bytes_per_step": pipe8.h2d_bytes, "d2h_bytes_per_step": pipe8.d2h_bytes,
                  "input": "uint8 BGR frames [n,368,656,3], normalisation fused into the stem"}
        del pipe8
    # the same streaming run from RAW camera frames (720x1280 uint8, what demo.py:91-93 hands to infer_fast): cubic resize + pad
    # on the GPU (row f1), result post-conversion to per-pose key-points / boxes on the GPU too (row f3)
    e2e_raw = None
    if not args.no_u8:
        from lwpose_b200 import postproc as _pp
        rscale, _, rpadded, rpad = _pp.infer_fast_geometry(720, 1280, HEIGHT)
        assert tuple(rpadded) == (HEIGHT, WIDTH), rpadded
        pipe_r = PosePipeline(net, args.batch, HEIGHT, WIDTH, precision=args.precision, demo=True,
                              heads_hook=lambda heads, lo: heads.add_(inject[lo:lo + heads.shape[0]]),
                              fused=not args.unfused_postproc, chunk=args.chunk or None,
                              overlap_postproc=not args.no_overlap_postproc, input_format="u8_raw", raw_size=(720, 1280),
                              convert=dict(pad=rpad, scale=rscale), depth=args.e2e_depth)
        xr = torch.from_numpy(synth.synthetic_frames(args.batch, 720, 1280, seed=1 + rank)).pin_memory()
        for _ in range(2):
            pipe_r(xr)
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        rres = stream_steps(pipe_r, xr, args.steps).check()
        r1.record()
        r1.synchronize()
        barrier()
        raw_ms = parallel.max_over_ranks(r0.elapsed_time(r1) / args.steps, device=dev)
        e2e_raw = {"value": world * args.batch / (raw_ms / 1000.0), "unit": "frames/s", "ms_per_step": raw_ms,
                   "h2d_bytes_per_step": pipe_r.h2d_bytes, "d2h_bytes_per_step": pipe_r.d2h_bytes,
                   "input": "raw uint8 BGR camera frames [n,720,1280,3] (pinned); cubic resize to 368x655 + pad to 368x656 + normalisation "
                            "on the GPU; output: pose tables + per-pose key-points / boxes in frame coordinates",
                   "pose_objects_frame0": len(rres.poses(0))}
        del pipe_r, xr
    # the other arithmetic of BASELINE.json configs[1] ("fp32 and bf16"): the same device-resident measurement on the
    # other precision's plan (fp32 storage + TF32 tensor-core products when the headline is bf16, and vice versa)
    other = None
    if not args.no_other_precision:
        oprec = "tf32" if args.precision == "bf16" else "bf16"
        pipe_o = PosePipeline(net, args.batch, HEIGHT, WIDTH, precision=oprec, demo=True,
                              heads_hook=lambda heads, lo: heads.add_(inject[lo:lo + heads.shape[0]]),
                              fused=not args.unfused_postproc, chunk=args.chunk or None,
                              overlap_postproc=not args.no_overlap_postproc)
        for _ in range(3):
            pipe_o.run_device(x_dev)
        barrier()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record()
        for _ in range(args.steps):
            pipe_o.run_device(x_dev)
        pipe_o.join()
        o1.record()
        barrier()
        o_ms = parallel.max_over_ranks(o0.elapsed_time(o1), device=dev) / args.steps
        other = {"dtype": oprec, "value": world * args.batch / (o_ms / 1000.0), "unit": "frames/s", "ms_per_step": o_ms}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            gpu_heads[oprec] = raw_heads(pipe_o, x_dev, cpu_n)
        if rank == 0 and not args.no_roofline:
            other["roofline"] = roofline_pass(pipe_o, x_dev, args, precision=oprec, brief=True)
        del pipe_o
    # latency of one synchronous call (H2D -> kernels -> D2H, nothing overlapped)
    t0 = time.perf_counter()
    for _ in range(3):
        pipe(x_host)
    sync_ms = (time.perf_counter() - t0) / 3 * 1000.0

    e2e_f32 = {"value": e2e_value, "unit": "frames/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": pipe.h2d_bytes,
               "d2h_bytes_per_step": pipe.d2h_bytes, "mode": "PosePipeline.submit/collect, %d batches in flight" % len(pipe.slots),
               "input": "float32 NCHW frames [n,3,368,656], already normalised on the host", "sync_call_ms": sync_ms}
    out = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": workload_config(args, world),
        # headline end-to-end number: what the reference's entry point is handed -- raw uint8 BGR frames
        # (demo.py:54 infer_fast(net, img, ...)) -- in pinned host memory, pose tables back in host memory.  The float32
        # NCHW variant (the tensor the reference builds on the host before its own H2D copy, 4x the bytes: 185 MB per
        # step, which is PCIe time comparable to the whole step) is reported next to it.
        "e2e": dict(e2e_u8, mode="PosePipeline.submit/collect, %d batches in flight" % args.e2e_depth) if e2e_u8 is not None else e2e_f32,
        "e2e_f32": e2e_f32,
        "e2e_raw_frames": e2e_raw,
        "steady_state": steady,
        "other_precision": other,
        "gpu_launches": pipe.launches_per_step * args.steps,
        "clocks": clocks,
        "poses_per_step_rank0": total_poses, "persons_injected_rank0": int(sum(persons)),
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # before the roofline pass re-runs single layers
        out["cpu_baseline"], out["postproc_parity"], out["net_parity"] = cpu_baseline(net, pipe, x_host, inject_h, res,
                                                                                      args, gpu_heads)
    if rank == 0 and not args.no_roofline:
        out.update(roofline_pass(pipe, x_dev, args))
    del pipe
    if not args.no_configs:
        out.update(extra_configs(args, net, dev, rank, world, barrier))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def extra_configs(args, net, dev, rank, world, barrier):
    """The other configurations BASELINE.json names, as extra keys of the same JSON line (the headline stays configs[1])."""
    out = {}
    out["config3"] = config3(args, dev, rank, world, barrier)
    out["config4"] = config4(args, net, dev, rank, world, barrier)
    if rank == 0:
        out["config0_latency"] = config0_latency(args, net, dev)
    return out


def config3(args, dev, rank, world, barrier):
    """configs[3]: 3 refinement stages, GLOBAL batch 256 at 368x656 split over the ranks (strong scaling: 256 / N
    frames per GPU), network + grouping, device-resident like `value`."""
    import numpy as np
    import torch
    from lwpose_b200 import parallel, synth
    from lwpose_b200.pipeline import PosePipeline
    lo, hi = parallel.shard_range(args.config3_batch, rank, world)
    per = hi - lo
    net3 = make_net(refine=3).to(dev)
    inj_h, _ = person_maps(per, rank, args.max_persons)
    inject = torch.from_numpy(inj_h).to(dev)
    base = synth.synthetic_net_input(min(per, 32), HEIGHT, WIDTH, seed=11 + rank).to(dev)
    x_dev = base.repeat((per + base.shape[0] - 1) // base.shape[0], 1, 1, 1)[:per].contiguous()
    pipe = PosePipeline(net3, per, HEIGHT, WIDTH, precision=args.precision, demo=True,
                        heads_hook=lambda heads, l: heads.add_(inject[l:l + heads.shape[0]]))
    for _ in range(3):
        pipe.run_device(x_dev)
    barrier()
    steps = max(3, min(args.steps, 10))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        pipe.run_device(x_dev)
    pipe.join()
    b.record()
    barrier()
    ms = parallel.max_over_ranks(a.elapsed_time(b), device=dev) / steps
    res = {"workload": "configs[3]: R=3 network, global batch %d @368x656 split %d per GPU over %d GPU(s), + person maps, "
                       "cubic x4, extract, group" % (args.config3_batch, per, world),
           "value": args.config3_batch / (ms / 1000.0), "unit": "frames/s", "ms_per_step": ms, "steps": steps,
           "scaling": "strong", "dtype": args.precision, "global_batch": args.config3_batch, "frames_per_gpu": per,
           "poses_rank0": int(pipe.n_poses.sum().item())}
    if rank == 0 and not args.no_roofline:
        rf = roofline_pass(pipe, x_dev, args, brief=True)
        res["roofline_gemm3x3"] = rf.get("roofline_gemm3x3")
        res["roofline"] = rf.get("roofline")
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import net as onet
        sd = {k: v.detach().cpu() for k, v in net3.state_dict().items()}
        gh = raw_heads(pipe, x_dev, 2)
        ref = onet.forward(sd, x_dev[:2].cpu())
        rh = np.concatenate([ref[-2].numpy(), ref[-1].numpy()], 1).transpose(0, 2, 3, 1)
        err = float(np.abs(gh - rh).max())
        res["net_parity"] = {"max_abs_err": err, "tol": NET_TOL[args.precision], "frames": 2}
        if not err < NET_TOL[args.precision]:
            raise RuntimeError("config3 net_parity FAILED: %g" % err)
    del pipe
    return res


def config4(args, net, dev, rank, world, barrier):
    """configs[4]: val.py-style multi-scale inference (scales 0.5 / 1 / 1.5 / 2 x 368: net inputs 368x368, 368x496, 552x736,
    736x984 for a 480x640 frame), maps resized x8, cropped, resized to the frame size and averaged, 18 x extract +
    group(demo=False) at 480x640, convert_to_coco_format -- end to end from pinned uint8 frames to COCO lists on the host,
    frames sharded over the ranks (each frame's four scales stay on one GPU)."""
    import numpy as np
    import torch
    from lwpose_b200 import parallel, postproc, synth, val
    Hf, Wf, chunk = 480, 640, int(os.environ.get("LWP_CONFIG4_CHUNK", "16"))
    scales = [0.5, 1.0, 1.5, 2.0]
    lo, hi = parallel.shard_range(args.config4_frames, rank, world)
    per = hi - lo
    frames = torch.from_numpy(synth.synthetic_frames(max(per, 1), Hf, Wf, seed=300 + rank)).pin_memory()
    # synthetic persons at the frame resolution: stride-8 maps (1..10 persons) up-sampled x8 once, added to the averaged maps
    hm, paf, persons = synth.synthetic_pose_maps(chunk, Hf // 8, Wf // 8, seed=400 + rank, max_persons=10)
    m = np.zeros((chunk, Hf // 8, Wf // 8, 64), np.float32)
    m[..., :19] = hm.transpose(0, 2, 3, 1)
    m[..., 19:57] = paf.transpose(0, 2, 3, 1)
    md = torch.from_numpy(m).to(dev)
    inj_h = postproc.upsample_cubic(md, channels=19, fx=8, fy=8, channel_offset=0)
    inj_p = postproc.upsample_cubic(md, channels=38, fx=8, fy=8, channel_offset=19)
    net.precision = args.precision

    def hook(avg_h, avg_p):
        avg_h.add_(inj_h[:avg_h.shape[0]])
        avg_p.add_(inj_p[:avg_p.shape[0]])

    def one_pass():
        n = 0
        for c0 in range(0, per, chunk):
            res = val.evaluate_batch(net, frames[c0:c0 + chunk], scales=scales, base_height=368, maps_hook=hook)
            n += sum(len(r[0]) for r in res)
        return n
    one_pass()
    torch.cuda.synchronize()
    barrier()
    steps = 2
    t0 = time.perf_counter()
    for _ in range(steps):
        n_poses = one_pass()
    torch.cuda.synchronize()
    dt = parallel.max_over_ranks((time.perf_counter() - t0) / steps, device=dev)
    barrier()
    res = {"workload": "configs[4]: %d frames 480x640 uint8 (global, %d per GPU), scales %s x 368, R=1 network, averaged maps + 1..10 "
                       "synthetic persons, extract / group(demo=False) at 480x640, COCO conversion; host frames in, host lists out"
                       % (args.config4_frames, per, scales),
           "value": args.config4_frames / dt, "unit": "frames/s", "ms_per_step": dt * 1000.0, "steps": steps, "scaling": "strong",
           "dtype": args.precision, "coco_poses_rank0": n_poses, "timing": "wall clock around the public call, device synchronised"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res["cpu_port"] = config4_cpu(net, frames[0].numpy(), scales, hm[0], paf[0])
    return res


def config4_cpu(net, frame, scales, hm8, paf8):
    """Oracle port of reference val.py:81-110,129-136 on ONE frame (host cv2 float64 input resize, torch CPU fp32 network at the
    four sizes, OpenCV-exact output resizes, averaging, extract / group(demo=False), COCO conversion)."""
    import cv2
    import numpy as np
    import torch
    from lwpose_b200 import val
    from oracle import net as onet
    from oracle import postproc as orc
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    torch.set_num_threads(os.cpu_count() or 1)
    inj_h = orc.resize_cubic(np.ascontiguousarray(hm8.transpose(1, 2, 0)), fx=8, fy=8)
    inj_p = orc.resize_cubic(np.ascontiguousarray(paf8.transpose(1, 2, 0)), fx=8, fy=8)
    t0 = time.perf_counter()
    height, width = frame.shape[:2]
    normed = val.normalize(frame, (128, 128, 128), 1 / 256)
    avg_h = np.zeros((height, width, 19), np.float32)
    avg_p = np.zeros((height, width, 38), np.float32)
    for s in scales:
        ratio = s * 368 / float(height)
        scaled = cv2.resize(normed, (0, 0), fx=ratio, fy=ratio, interpolation=cv2.INTER_CUBIC)
        padded, pad = val.pad_width(scaled, 8, (0, 0, 0), [368, max(scaled.shape[1], 368)])
        outs = onet.forward(sd, torch.from_numpy(padded).permute(2, 0, 1).unsqueeze(0).float())
        for o, avg in ((outs[-2], avg_h), (outs[-1], avg_p)):
            mm = orc.resize_cubic(np.ascontiguousarray(o[0].numpy().transpose(1, 2, 0)), fx=8, fy=8)
            mm = mm[pad[0]:mm.shape[0] - pad[2], pad[1]:mm.shape[1] - pad[3], :]
            avg += orc.resize_cubic(np.ascontiguousarray(mm), dsize=(width, height)) / len(scales)
    avg_h += inj_h
    avg_p += inj_p
    total, by_type = 0, []
    for k in range(18):
        total += orc.extract_keypoints(avg_h[:, :, k], by_type, total)
    poses, allk = orc.group_keypoints(by_type, avg_p, demo=False)
    coco, _ = val.convert_to_coco_format(poses, allk)
    dt = time.perf_counter() - t0
    return {"value": 1.0 / dt, "unit": "frames/s", "seconds_per_frame": dt, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "1 frame, 4 scales", "coco_poses": len(coco)}


def config0_latency(args, net, dev):
    """configs[0]: ONE 1x3x256x456 frame end to end -- the only thing the reference itself ever runs (demo.py:91-100).
    (a) the drop-in call pattern infer_fast + 18 x extract_keypoints + group_keypoints on a 720p frame, (b) the
    batch-1 PosePipeline replaying one CUDA graph per pass, (c) the CPU port of the reference on the same frame."""
    import cv2
    import numpy as np
    import torch
    from lwpose_b200 import demo, synth, val
    from lwpose_b200.pipeline import PosePipeline
    H0, W0 = 256, 456
    img = np.random.default_rng(0).integers(0, 256, (720, 1280, 3), dtype=np.uint8)
    scale = H0 / img.shape[0]
    scaled = cv2.resize(img, (0, 0), fx=scale, fy=scale, interpolation=cv2.INTER_CUBIC)
    # pad with the mean: the stem normalises (128 - 128) / 256 = 0, the reference's pad value in the normalised image
    padded_u8, pad = val.pad_width(scaled, 8, (128, 128, 128), [H0, max(scaled.shape[1], H0)])
    assert padded_u8.shape == (H0, W0, 3), padded_u8.shape
    hm, paf, _ = synth.synthetic_pose_maps(1, H0 // 8, W0 // 8, seed=5, persons=5)
    inj_h = np.zeros((1, H0 // 8, W0 // 8, 64), np.float32)
    inj_h[..., :19] = hm.transpose(0, 2, 3, 1)
    inj_h[..., 19:57] = paf.transpose(0, 2, 3, 1)
    inject = torch.from_numpy(inj_h).to(dev)
    out = {"workload": "configs[0]: one 720x1280 BGR frame -> 1x3x256x456, R=1 network, 5 synthetic persons added to the "
                       "heads, cubic x4, extract, group (demo=True)", "dtype": args.precision, "unit": "ms per frame"}
    # (b) batch-1 pipeline, one CUDA graph per pass; frame already resized + padded on the host (uint8)
    x8 = torch.from_numpy(np.ascontiguousarray(padded_u8[None])).pin_memory()
    for graph in (True, False):
        pipe = PosePipeline(net, 1, H0, W0, precision=args.precision, demo=True, input_format="u8_nhwc", depth=1,
                            graph=graph, heads_hook=lambda heads, l: heads.add_(inject[l:l + heads.shape[0]]))
        for _ in range(5):
            r = pipe(x8)
        ts = []
        for _ in range(200):
            t0 = time.perf_counter()
            r = pipe(x8)
            ts.append((time.perf_counter() - t0) * 1000.0)
        ts.sort()
        key = "pipeline_graph" if graph else "pipeline_stream_launches"
        out[key] = {"median_ms": ts[len(ts) // 2], "p90_ms": ts[int(len(ts) * 0.9)], "frames_per_s": 1000.0 / ts[len(ts) // 2],
                    "poses": r.check().total_poses(), "h2d_bytes": pipe.h2d_bytes, "d2h_bytes": pipe.d2h_bytes}
        del pipe
    # (a) the reference's own call pattern through the drop-in functions (host cv2 resize, maps back on the host,
    # 18 + 1 synchronous post-processing calls)
    net.precision = args.precision
    for _ in range(2):
        demo.run_frame(net, img, H0)
    t0 = time.perf_counter()
    for _ in range(10):
        demo.run_frame(net, img, H0)
    out["dropin_run_frame_ms"] = (time.perf_counter() - t0) / 10 * 1000.0
    # (c) CPU port of the reference on the same frame (host resize + normalise + pad + fp32 forward + post-processing)
    if not args.no_cpu_baseline:
        sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}

        def cpu_once():
            sc = cv2.resize(img, (0, 0), fx=scale, fy=scale, interpolation=cv2.INTER_CUBIC)
            pd, _ = val.pad_width(val.normalize(sc, (128, 128, 128), 1 / 256), 8, (0, 0, 0), [H0, max(sc.shape[1], H0)])
            x = torch.from_numpy(pd).permute(2, 0, 1).unsqueeze(0).float()
            return cpu_frame(sd, x, inj_h[0])
        torch.set_num_threads(os.cpu_count() or 1)
        cpu_once()
        t0 = time.perf_counter()
        for _ in range(5):
            n_cpu = cpu_once()
        out["cpu_port_ms"] = (time.perf_counter() - t0) / 5 * 1000.0
        out["cpu_port_poses"] = n_cpu
        out["cpu_cores"] = torch.get_num_threads()
    return out


def ncu_traffic(prefixes):
    """Average dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernels from the newest committed
    ncu --set full capture (profiles/r*_dram_traffic_bytes.json); None when there is none."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_dram_traffic_bytes.json")))
    if not files:
        return None, None
    with open(files[-1]) as f:
        data = json.load(f)
    vals = [b for k, layers in data.items() if any(k.startswith(p) for p in prefixes) for b in layers.values()]
    if not vals:
        return None, None
    return sum(vals) / len(vals), "%s (mean over the %d captured launches)" % (
        os.path.relpath(files[-1], ROOT).replace("dram_traffic_bytes.json", "ncu_full_layers.csv"), len(vals))


def stream_steps(pipe, frames, steps):
    """`steps` batches through submit()/collect() with as many in flight as the pipeline has slots; returns the last
    PoseResult.  Every batch's H2D copy and result read-back happen inside the caller's timed region."""
    depth = len(pipe.slots)
    res, submitted, collected = None, 0, 0
    while collected < steps:
        while submitted < steps and submitted - collected < depth:
            pipe.submit(frames)
            submitted += 1
        res = pipe.collect()
        collected += 1
    return res


def roofline_pass(pipe, x_dev, args, precision=None, brief=False):
    """Per-kernel device time (CUDA events on the launching stream, op by op, after the timed region) ->
    achieved TFLOP/s of the tcgen05 GEMM kernel and GB/s of the depthwise kernel vs the measured peaks.
    Times are summed over the chunks of one step."""
    import torch
    pk = peaks()
    reps = max(3, min(args.steps, 10))
    plan0 = pipe.chunks[0].plan
    nops = plan0.num_compute_ops
    times = [0.0] * nops
    # per op: MEDIAN over the repetitions (a host-side hiccup between two launches idles the GPU inside one event
    # interval; one such outlier in a mean of <= 10 would distort that layer's -- and its roofline object's -- number)
    for c in pipe.chunks:
        plan, xc = c.plan, x_dev[c.lo:c.lo + c.n]
        for i in range(nops):
            plan.run(xc, i, i + 1)  # warm
        torch.cuda.synchronize()
        samples = [[] for _ in range(nops)]
        for _ in range(reps):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(nops + 1)]
            evs[0].record()
            for i in range(nops):
                plan.run(xc, i, i + 1)
                evs[i + 1].record()
            torch.cuda.synchronize()
            for i in range(nops):
                samples[i].append(evs[i].elapsed_time(evs[i + 1]))
        for i in range(nops):
            times[i] += sorted(samples[i])[len(samples[i]) // 2]
    agg = {}
    for i, t in enumerate(times):
        kind = plan0.op_meta[i]["kind"]
        a = agg.setdefault(kind, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
        a["ms"] += t
        for c in pipe.chunks:
            a["flops"] += c.plan.op_meta[i]["flops"]; a["bytes"] += c.plan.op_meta[i]["bytes"]; a["launches"] += 1
    tc_kinds = [k for k in agg if k.startswith("gemm") or k == "dwpw"]   # the tcgen05 implicit-GEMM kernels (the fused
    # depthwise+1x1 blocks of the thin layers are HBM-bound: they get their own object below)
    gemm_ms = sum(agg[k]["ms"] for k in tc_kinds)
    gemm_flops = sum(agg[k]["flops"] for k in tc_kinds)
    gemm_launches = sum(agg[k]["launches"] for k in tc_kinds)
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    precision = precision or args.precision
    peak = pk["bf16"] if precision == "bf16" else pk["bf16"] / 2.0
    traffic, traffic_src = ncu_traffic(("conv_gemm", "conv3x3_pair", "dwpw_gemm_kernel"))
    res = {"roofline": {"kernel": "tcgen05 implicit-GEMM kernels conv_gemm_kernel / conv_gemm_wres_kernel / conv_gemm2_kernel / conv3x3_pair_kernel%s (all %d launches of a "
                                  "step; %.0f %% of the network time)" % (" / dwpw_gemm_kernel" if "dwpw" in agg else "", gemm_launches,
                                                                         100.0 * gemm_ms / max(sum(times), 1e-9)),
                        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                        "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                        "peak_source": pk["source"] + (" bf16 sustained" if precision == "bf16"
                                                       else " bf16 sustained / 2 (tf32 nominal ratio)"),
                        "algorithmic_flops_per_launch": gemm_flops / max(gemm_launches, 1),
                        "avg_launch_ms": gemm_ms / max(gemm_launches, 1), "ms_per_step": gemm_ms}}
    if "gemm3x3" in agg:
        d = agg["gemm3x3"]
        tf = d["flops"] / (d["ms"] * 1e-3) / 1e12
        res["roofline_gemm3x3"] = {"kernel": "conv3x3_pair_kernel (CTA pairs + activation strips), dense 3x3 layers (%d launches)" % d["launches"],
                                   "bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                                   "ms_per_step": d["ms"]}
    if "gemm1x1" in agg:
        d = agg["gemm1x1"]
        tf = d["flops"] / (d["ms"] * 1e-3) / 1e12
        res["roofline_pointwise"] = {"kernel": "conv_gemm_kernel / conv_gemm_wres_kernel (thin layers) / conv_gemm2_kernel (CTA pairs), 1x1 layers (%d launches)" % d["launches"],
                                     "bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                                     "ms_per_step": d["ms"]}
        big = [(plan0.op_meta[i]["flops"] * len(pipe.chunks), times[i]) for i in range(nops)
               if plan0.op_meta[i]["kind"] == "gemm1x1" and plan0.op_names[i] in ("model.%d.pw" % k for k in range(7, 12))]
        if big:
            tfb = sum(f for f, _ in big) / (sum(t for _, t in big) * 1e-3) / 1e12
            res["roofline_pointwise"]["backbone_512x512"] = {"achieved": tfb, "frac": tfb / peak, "launches": len(big)}
    if "dwpw" in agg:
        d = agg["dwpw"]
        gbs = d["bytes"] / (d["ms"] * 1e-3) / 1e9
        res["roofline_dwpw_hbm"] = {"kernel": "dwpw_gemm_kernel as an HBM-bound fused block (%d launches)" % d["launches"],
                                    "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
                                    "frac": gbs / pk["hbm"], "ms_per_step": d["ms"]}
    if "sepconv" in agg:
        d = agg["sepconv"]
        gbs = d["bytes"] / (d["ms"] * 1e-3) / 1e9
        res["roofline_sepconv"] = {"kernel": "sepconv_kernel: weight-resident fused depthwise 3x3 + 1x1 blocks of the thin layers (%d launches); "
                                             "bytes = block input + block output (+ weights)" % d["launches"],
                                   "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
                                   "tflops": d["flops"] / (d["ms"] * 1e-3) / 1e12, "ms_per_step": d["ms"]}
    if "depthwise" in agg:
        d = agg["depthwise"]
        gbs = d["bytes"] / (d["ms"] * 1e-3) / 1e9
        traffic_d, src_d = ncu_traffic(("depthwise3x3_tma_kernel",))
        res["roofline_depthwise"] = {"kernel": "depthwise3x3_tma_kernel (%d launches)" % d["launches"], "bound": "hbm",
                                     "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
                                     "traffic": traffic_d, "traffic_source": src_d, "ms_per_step": d["ms"]}
    res["kernel_ms_per_step"] = {k: round(v["ms"], 4) for k, v in agg.items()}
    if brief:   # the other precision of configs[1]: the three roofline objects only
        return {k: {kk: vv for kk, vv in v.items() if kk in ("bound", "achieved", "peak", "unit", "frac", "ms_per_step")}
                for k, v in res.items() if k.startswith("roofline")}
    res["kernel_ms_per_step"].update(postproc_stage_ms(pipe, x_dev, reps))
    res["layer_ms"] = {n: round(t, 4) for n, t in zip(plan0.op_names[:nops], times)}
    return res


def postproc_stage_ms(pipe, x_dev, reps):
    """Device time of the post-processing stages of one step (events on the launching stream)."""
    import torch
    pipe.run_device(x_dev)  # leaves net output + injected persons in the head buffers
    names = ["extract_fused", "group_fused"] if pipe.fused else ["upsample", "extract", "group"]
    out = {}
    for name in names:
        for c in pipe.chunks:
            c.enqueue_postproc(c.heads, stage=name)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            for c in pipe.chunks:
                c.enqueue_postproc(c.heads, stage=name)
        b.record()
        torch.cuda.synchronize()
        out[name] = round(a.elapsed_time(b) / reps, 4)
    return out


def cpu_baseline(net, pipe, x_host, inject_h, res, args, gpu_heads):
    """The oracle port of the reference's CPU path timed on this box's host cores on a bounded sample; the heads the
    GPU network produced for those very frames compared with the CPU fp32 forward (net_parity: the run FAILS above the
    stated tolerance -- a wrong network cannot print a number); and a bit-exactness check of the GPU post-processing
    against the oracle on the GPU's own maps."""
    import numpy as np
    import torch
    from lwpose_b200 import postproc
    from oracle import postproc as orc
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    frames = max(1, min(args.cpu_frames, args.batch))
    cpu_heads = []
    dt = cpu_time_frames(sd, x_host[:frames], inject_h[:frames], frames, keep=cpu_heads)
    base = {"value": frames / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d of the %d frames of one step, batch-1 loop: torch CPU fp32 forward + C oracle cubic x4 / "
                      "extract / group" % (frames, args.batch)}
    cpu_heads = np.stack(cpu_heads)
    net_parity = {}
    for prec, gh in gpu_heads.items():
        err = float(np.abs(gh[:frames] - cpu_heads).max())
        net_parity[prec] = {"max_abs_err": err, "tol": NET_TOL[prec], "frames": frames,
                            "ref_max_abs": float(np.abs(cpu_heads).max()),
                            "what": "last-stage heat-maps + PAFs [frames,46,82,57] of the timed batch, GPU %s vs torch CPU fp32 oracle" % prec}
        if not err < NET_TOL[prec]:
            raise RuntimeError("net_parity FAILED: %s heads differ from the CPU fp32 oracle by %g (tol %g)" % (prec, err, NET_TOL[prec]))
    # parity: oracle post-processing on the maps the GPU actually produced (heads + injected persons)
    heads = pipe.heads.cpu().numpy()
    checked = 0
    for b in range(min(4, args.batch)):
        heat = orc.resize_cubic(np.ascontiguousarray(heads[b, :, :, :19]), fx=4, fy=4)
        pafs = orc.resize_cubic(np.ascontiguousarray(heads[b, :, :, 19:57]), fx=4, fy=4)
        total, by_type = 0, []
        for k in range(18):
            total += orc.extract_keypoints(heat[:, :, k], by_type, total)
        ref_poses, _ = orc.group_keypoints(by_type, pafs, demo=True)
        got_poses, _ = res.frame(b)
        rp = np.asarray(ref_poses, np.float64).reshape(-1, 20)
        gp = np.asarray(got_poses, np.float64).reshape(-1, 20)
        if rp.shape != gp.shape or not np.array_equal(rp.view(np.int64), gp.view(np.int64)):
            raise RuntimeError("postproc_parity FAILED: pose table of frame %d differs from the oracle" % b)
        checked += 1
    return base, "bit-exact pose tables vs oracle on %d frames" % checked, net_parity


if __name__ == "__main__":
    main()
