"""Plain forward passes of the full model for profiling.  usage: run_forward.py [B] [precision] [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from poisson_cnn_b200.synthetic import make_problem

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else "tc2"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda", 0)
model, _ = bench.build_model(dev, prec)
p = make_problem(min(B, 16), 256, 256, seed=1001)
inp = [p[k].repeat(-(-B // p[k].shape[0]), *([1] * (p[k].dim() - 1)))[:B].contiguous().cuda() for k in bench.KEYS]
for _ in range(iters):
    out = model(inp)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
