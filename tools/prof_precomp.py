import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gpt-pinn_b200"), os.path.join(ROOT, "gpt-pinn_b200", "dropin")):
    sys.path.insert(0, p)
import bench
from gptpinn import precomp
dev = torch.device("cuda", 0)
nc = 1000
x = torch.linspace(-1.0, 1.0, nc, device=dev); t = torch.linspace(0.0, 5.0, nc, device=dev)
xt = torch.stack((x.repeat(nc), t.repeat_interleave(nc)), dim=1).contiguous()
layers = [(w.to(dev), b.to(dev)) for w, b in bench.synthetic_weights(1000)]
bufs = [torch.empty((nc * nc, 15), device=dev) for _ in range(3)]
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    precomp.mlp_channels(layers, "cos", xt, v=bufs[0][:, 3], q=bufs[1][:, 3], r=bufs[2][:, 3])
    e1.record(); torch.cuda.synchronize()
    print("1e6 points, 5 channels:", e0.elapsed_time(e1), "ms")
