"""BASELINE configs 3/4: throughput of the full model on other grid shapes (precision mixed), checked against the FP32 path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from poisson_cnn_b200.synthetic import make_problem
dev = torch.device("cuda", 0)
model, _ = bench.build_model(dev, "mixed")
for nx, ny, B in ((384, 128, 128), (512, 256, 128), (200, 300, 128), (1024, 1024, 16), (2048, 2048, 16)):
    p = make_problem(2, nx, ny, seed=1003)
    inp = [p[k].repeat(B // 2, *([1] * (p[k].dim() - 1))).contiguous().cuda() for k in bench.KEYS]
    model.set_precision("mixed")
    for _ in range(2):                      # the first pass of a new shape allocates the activation pool
        out = model(inp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = model(inp); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ref = model.set_precision("fp32")([t[:2] for t in inp])
    err = float((out[:2].double() - ref.double()).norm() / ref.double().norm())
    print("%4dx%-4d batch %3d: %8.1f ms -> %7.1f sol/s (%.2f Mpixel/s), mixed vs fp32 path rel-L2 %.2e, peak mem %.0f GB" % (
        nx, ny, B, ms, B * 1e3 / ms, B * nx * ny / ms / 1e3, err, torch.cuda.max_memory_allocated() / 1e9), flush=True)
