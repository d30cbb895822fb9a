"""Forward passes of the full model through the model-level C ABI, for `ncu` launch lists.
usage: profile_forward.py [B] [precision] [nx] [ny] [passes]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from poisson_cnn_b200.synthetic import make_problem
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prec = sys.argv[2] if len(sys.argv) > 2 else "mixed"
nx = int(sys.argv[3]) if len(sys.argv) > 3 else 256
ny = int(sys.argv[4]) if len(sys.argv) > 4 else nx
passes = int(sys.argv[5]) if len(sys.argv) > 5 else 2
dev = torch.device("cuda", 0)
model, _ = bench.build_model(dev, prec)
p = make_problem(min(B, 8), nx, ny, seed=1001)
inp = [p[k].repeat(-(-B // p[k].shape[0]), *([1] * (p[k].dim() - 1)))[:B].contiguous().cuda() for k in bench.KEYS]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(passes):
    if i == passes - 1:
        e0.record()
    out = model(inp)
e1.record(); torch.cuda.synchronize()
print("B=%d %s %dx%d: last forward %.2f ms (%.1f solutions/s), finite=%s" % (B, prec, nx, ny, e0.elapsed_time(e1), B * 1e3 / e0.elapsed_time(e1), bool(torch.isfinite(out).all())))
