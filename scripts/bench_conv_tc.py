"""Micro-benchmark of the tcgen05 convolution on the dominant layer shapes (CUDA events, L2-exceeding inputs)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poisson_cnn_b200 import ops

def bench(B, Cin, Cout, H, W, k, iters=5, f32=False):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, Cin, H, W, generator=g).cuda()
    kern = (torch.randn(k, k, Cin, Cout, generator=g) / (k * Cin ** 0.5)).cuda()
    bias = torch.zeros(Cout).cuda()
    flops = 2.0 * B * H * W * k * k * Cin * Cout
    t = ops.to_blk8(x); wp = ops.pack_conv_weights_tc(kern)
    out = ops.Blk8(B, Cout, H, W, x.device)
    for _ in range(2): ops.conv2d_tc(t, wp, bias, 1, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.conv2d_tc(t, wp, bias, 1, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    r = {"shape": [B, Cin, Cout, H, W, k], "tc_ms": ms, "tc_tflops_useful": flops / ms / 1e9}
    if f32:
        ops.conv2d(x, kern, bias, 1)
        torch.cuda.synchronize(); e0.record()
        for _ in range(2): o32 = ops.conv2d(x, kern, bias, 1)
        e1.record(); torch.cuda.synchronize()
        r["f32_ms"] = e0.elapsed_time(e1) / 2
        got = ops.from_blk8(out)
        r["rel_l2_vs_f32"] = float((got - o32).norm() / o32.norm())
    print(json.dumps(r), flush=True)

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    bench(B, 32, 32, 256, 256, 15, f32=True)
    bench(B, 32, 28, 256, 256, 13)
    bench(B, 64, 32, 256, 256, 7)
    bench(B, 32, 32, 256, 256, 7)
    bench(B, 32, 32, 128, 128, 11)
    bench(B, 24, 24, 256, 256, 9)
    bench(B, 16, 16, 256, 256, 5)
    bench(B, 12, 12, 256, 256, 3)
    bench(4 * B, 29, 23, 256, 256, 7)
