"""Warm, in-pipeline time per C-ABI entry point: wraps every libpcnn call in CUDA events (same stream) and sums
the elapsed time per function over one forward pass.  usage: profile_ops.py [B] [precision]"""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from poisson_cnn_b200 import ops, _lib
from poisson_cnn_b200.synthetic import make_problem

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prec = sys.argv[2] if len(sys.argv) > 2 else "tc2"
dev = torch.device("cuda", 0)
model, _ = bench.build_model(dev, prec)
p = make_problem(min(B, 16), 256, 256, seed=1001)
inp = [p[k].repeat(-(-B // p[k].shape[0]), *([1] * (p[k].dim() - 1)))[:B].contiguous().cuda() for k in bench.KEYS]
for _ in range(2):
    model(inp)
torch.cuda.synchronize()

records = []
class Prof:
    def __init__(self, lib): self._lib = lib
    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.startswith("pcnn_") or name in ("pcnn_last_error", "pcnn_launch_count", "pcnn_blk8_bytes", "pcnn_conv_tc_packed_weight_bytes", "pcnn_version", "pcnn_dst_workspace_bytes"):
            return fn
        def wrapped(*a):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); r = fn(*a); e1.record()
            records.append((name, e0, e1))
            return r
        return wrapped
ops.lib = Prof(_lib.lib)
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); model(inp); t1.record(); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for name, a, b in records:
    agg[name][0] += 1; agg[name][1] += a.elapsed_time(b)
total = t0.elapsed_time(t1); s = sum(v[1] for v in agg.values())
print("B=%d %s: forward %.2f ms (%.3f ms/sample); sum over ops %.2f ms; torch/other %.2f ms" % (B, prec, total, total / B, s, total - s))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("  %-34s n=%4d %8.2f ms %5.1f%%" % (k, n, t, 100 * t / total))
