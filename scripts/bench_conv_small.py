"""Per-tile cost of the tcgen05 conv on small-K layers: separates MMA floor from epilogue cost (activation / mode)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poisson_cnn_b200 import ops

def bench(B, Cin, Cout, H, W, k, mode, act, iters=5):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, Cin, H, W, generator=g).cuda()
    kern = (torch.randn(k, k, Cin, Cout, generator=g) / (k * Cin ** 0.5)).cuda()
    bias = torch.zeros(Cout).cuda()
    t = ops.to_blk8(x, split=mode); wp = ops.pack_conv_weights_tc(kern, nsplit=mode)
    out = ops.Blk8(B, Cout, H, W, x.device, split=mode)
    for _ in range(2): ops.conv2d_tc(t, wp, bias, act, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.conv2d_tc(t, wp, bias, act, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    ntile = 256 if W >= 256 else -(-W // 16) * 16
    tiles = B * -(-H // 4) * -(-W // ntile)
    per_cta = -(-tiles // 148)
    nv = -(-Cin // 16) * {1: 1, 2: 3, 3: 2}[mode]
    print("B%d %d->%d %dx%d k%d mode%d act%d: %.3f ms, %.1f us/tile, MMA floor %.1f us/tile" % (
        B, Cin, Cout, H, W, k, mode, act, ms, ms * 1e3 / per_cta, nv * k * (k + 3) * 128 * (ntile / 256) / 1.9e3), flush=True)

for mode in (1, 3):
    for act in (0, 1, 2):
        bench(256, 15, 15, 256, 256, 5, mode, act)
bench(256, 5, 5, 256, 256, 3, 3, 2)
bench(64, 32, 32, 64, 64, 7, 3, 1)
bench(64, 32, 32, 32, 32, 7, 3, 1)
bench(64, 32, 32, 16, 16, 7, 3, 1)
