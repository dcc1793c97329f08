"""Cycles per phase of the fused front-end kernel (frontend_fused.cu), CTA 0, summed over its tiles.
Needs a library built with LWP_NVCC_EXTRA=-DLWP_TIMING_EXPERIMENTS and LWP_ALLOW_TIMING_EXPERIMENTS=1 LWP_FRONTEND_FUSION=1.
U8=1 feeds uint8 frames."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import lwpose_b200  # noqa: E402,F401
from lwpose_b200 import synth, _lib  # noqa: E402
from lwpose_b200.pipeline import PosePipeline  # noqa: E402

net = bench.make_net().cuda()
u8 = os.environ.get("U8") == "1"
pipe = PosePipeline(net, 64, bench.HEIGHT, bench.WIDTH, precision="bf16", demo=True, **({"input_format": "u8_nhwc"} if u8 else {}))
x = (torch.from_numpy(synth.synthetic_frames(64, bench.HEIGHT, bench.WIDTH, seed=1)).cuda() if u8
     else synth.synthetic_net_input(64, bench.HEIGHT, bench.WIDTH, seed=1).cuda())
plan = pipe.chunks[0].plan
assert plan.op_names[0] == "model.0-2.frontend", "set LWP_FRONTEND_FUSION=1"
plan.run(x, 0, 1)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    plan.run(x, 0, 1)
b.record()
torch.cuda.synchronize()
print("frontend us:", a.elapsed_time(b) * 100)
L = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_longlong * 16)()
fn = L.lwp_debug_frontend_prof
fn.argtypes = [ctypes.c_void_p]
assert fn(buf) == 0
names = ["0 patch->smem", "1 prefetch issue", "2 im2col", "3 stem mma", "4 stem epi", "5 dw1", "6 pw mma", "7 pw epi", "8 dw2"]
tot = sum(buf[:9])
for k, nme in enumerate(names):
    print("%-18s %9d cycles  %5.1f %%" % (nme, buf[k], 100.0 * buf[k] / tot))
print("tiles of CTA 0: ~57; cycles per tile:", tot / 57)
