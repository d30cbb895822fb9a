"""Host-side cost of enqueueing one forward pass: wall time of a tiny batch (GPU time negligible) vs its CUDA-event time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from poisson_cnn_b200.synthetic import make_problem
dev = torch.device("cuda", 0)
model, _ = bench.build_model(dev, sys.argv[2] if len(sys.argv) > 2 else "tc2")
for B in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1,16,64").split(",")]:
    p = make_problem(min(B, 16), 256, 256, seed=1001)
    inp = [p[k].repeat(-(-B // p[k].shape[0]), *([1] * (p[k].dim() - 1)))[:B].contiguous().cuda() for k in bench.KEYS]
    for _ in range(3): model(inp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(5): model(inp)
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("B=%d eager: enqueue %.2f ms/forward, GPU %.2f ms/forward, wall %.2f ms/forward" % (B, (t1 - t0) * 200, e0.elapsed_time(e1) / 5, (t2 - t0) * 200))
    if B <= 128:
        g = model.capture(inp)
        for _ in range(2): g(inp)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); e0.record()
        for _ in range(5): g(inp)
        e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print("B=%d graph: enqueue %.2f ms/forward, GPU %.2f ms/forward, wall %.2f ms/forward" % (B, (t1 - t0) * 200, e0.elapsed_time(e1) / 5, (t2 - t0) * 200))
