"""Forward time of a 256-sample batch for several micro-batch sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from poisson_cnn_b200.synthetic import make_problem
dev = torch.device("cuda", 0)
model, _ = bench.build_model(dev, sys.argv[1] if len(sys.argv) > 1 else "mixed")
B = 256
p = make_problem(16, 256, 256, seed=1001)
inp = [p[k].repeat(B // 16, *([1] * (p[k].dim() - 1))).contiguous().cuda() for k in bench.KEYS]
from poisson_cnn_b200 import ops
for mb in (32, 64, 128):
    ops.blk8_pool_clear(); torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
    model.max_microbatch = mb
    for _ in range(2): model(inp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): model(inp)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("microbatch %3d: %.1f ms per 256 -> %.0f sol/s, peak mem %.1f GB" % (mb, ms, 256e3 / ms, torch.cuda.max_memory_allocated() / 1e9), flush=True)
