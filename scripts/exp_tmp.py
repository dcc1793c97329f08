import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import lwpose_b200
from lwpose_b200 import synth
from lwpose_b200.pipeline import PosePipeline
net = bench.make_net().cuda()
pipe = PosePipeline(net, 64, bench.HEIGHT, bench.WIDTH, precision="bf16", demo=True)
x = synth.synthetic_net_input(64, bench.HEIGHT, bench.WIDTH, seed=1).cuda()
pipe.run_device(x); torch.cuda.synchronize()
h = pipe.chunks[0].heads[..., :18].abs()
print("net heat-map |v|: max %.4f mean %.4f  frac>0.0529: %.4f  frac>0.1: %.5f" % (h.max().item(), h.mean().item(), (h > 0.0529).float().mean().item(), (h > 0.1).float().mean().item()))
# per (tile of 8x8 source px, channel) fraction with max > bound
hm = torch.nn.functional.max_pool2d(h.permute(0, 3, 1, 2), kernel_size=11, stride=8, padding=2)
print("windows (11x11 src) with max > 0.0529: %.4f" % (hm > 0.0529).float().mean().item())
