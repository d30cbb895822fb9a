"""Warm, in-pipeline time per C-ABI entry point: wraps every libpcnn call in CUDA events (same stream) and sums
the elapsed time per function over one forward pass.  usage: profile_ops.py [B] [precision] [nx] [ny]"""
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
NX = int(sys.argv[3]) if len(sys.argv) > 3 else 256
NY = int(sys.argv[4]) if len(sys.argv) > 4 else NX
p = make_problem(min(B, 16), NX, NY, seed=1001)
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
            extra = None
            if name == "pcnn_conv2d_tc": extra = a[11:21]
            elif name == "pcnn_conv2d_f32": extra = ("f32",) + tuple(a[8:15])
            elif name == "pcnn_deconv_same_f32": extra = ("deconv",) + tuple(a[4:14])
            elif name == "pcnn_avgpool_same_f32": extra = ("pool",) + tuple(a[2:7])
            records.append((name, e0, e1, extra))
            return r
        return wrapped
ops.lib = Prof(_lib.lib)
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); model(inp); t1.record(); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
tc = collections.OrderedDict(); other = collections.OrderedDict()
for name, a, b, args in records:
    agg[name][0] += 1; agg[name][1] += a.elapsed_time(b)
    if args is not None and isinstance(args[0], str):
        other.setdefault(args, [0, 0.0]); other[args][0] += 1; other[args][1] += a.elapsed_time(b)
    elif args is not None:
        Bn, cin, cout, _, _, H, W, k, act, ns = args
        key = (Bn, cin, cout, H, W, k)
        tc.setdefault(key, [0, 0.0, ns])
        tc[key][0] += 1; tc[key][1] += a.elapsed_time(b)
total = t0.elapsed_time(t1); s = sum(v[1] for v in agg.values())
print("B=%d %s: forward %.2f ms (%.3f ms/sample); sum over ops %.2f ms; torch/other %.2f ms" % (B, prec, total, total / B, s, total - s))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("  %-34s n=%4d %8.2f ms %5.1f%%" % (k, n, t, 100 * t / total))

print("conv2d_tc launches by shape (B, Cin_total, Cout, H, W, k): n, total ms, fraction of the MMA floor (128 cyc/MMA @1.9 GHz, 148 SMs)")
for (Bn, cin, cout, H, W, k), (n, t, ns) in sorted(tc.items(), key=lambda kv: -kv[1][1]):
    ntile = 256 if W >= 256 else -(-W // 16) * 16
    tiles = Bn * -(-H // 4) * -(-W // ntile)
    nv = -(-cin // 16) * {1: 1, 2: 3, 3: 2}[ns]
    cp = _lib.lib.pcnn_conv_tc_channel_slots(cout, k); rt = 128 // cp
    tiles = Bn * (-(-H // rt)) * (-(-W // ntile))
    floor_ms = -(-tiles // 148) * nv * k * (k + rt - 1) * 128 * (ntile / 256.0) / 1.9e6
    print("  %-32s n=%2d %8.2f ms  floor %.2f ms  -> %.0f%%" % ((Bn, cin, cout, H, W, k), n, t, floor_ms * n, 100 * floor_ms * n / t))

print("other launches by shape:")
for k, (n, t) in sorted(other.items(), key=lambda kv: -kv[1][1])[:24]:
    print("  %-60s n=%2d %7.3f ms" % (k, n, t))
