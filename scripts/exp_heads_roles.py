"""Per-role cycle accounting of heads_fused_kernel (total / waiting cycles per role, mean over CTAs).
Needs a library built with LWP_NVCC_EXTRA=-DLWP_TIMING_EXPERIMENTS and LWP_ALLOW_TIMING_EXPERIMENTS=1."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, numpy as np
import bench, lwpose_b200
from lwpose_b200 import synth, _lib
from lwpose_b200.pipeline import PosePipeline
net = bench.make_net().cuda()
pipe = PosePipeline(net, 64, bench.HEIGHT, bench.WIDTH, precision="bf16", demo=True)
x = synth.synthetic_net_input(64, bench.HEIGHT, bench.WIDTH, seed=1).cuda()
pipe.run_device(x); torch.cuda.synchronize()
plan = pipe.chunks[0].plan
L = ctypes.CDLL(_lib.LIB_PATH)
roles = ["Xprod", "GEMM1", "Wprod", "GEMM2", "conv00", "epi0", "-", "-"]
for i, nm in enumerate(plan.op_names[:plan.num_compute_ops]):
    if "heads" not in nm: continue
    plan.run(x, i, i + 1); torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (160 * 16))()
    fn = L.lwp_debug_heads_prof; fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
    assert fn(buf, 160 * 16) == 0
    a = np.array(buf[:148 * 16]).reshape(148, 8, 2)
    print(nm, {r: (int(a[:, k, 0].mean()), int(a[:, k, 1].mean())) for k, r in enumerate(roles)})
