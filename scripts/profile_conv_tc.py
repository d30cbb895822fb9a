"""One conv_tc launch set for ncu / timing.  usage: profile_conv_tc.py [B] [mode 1|2|3] [Cin] [Cout] [k] [residual 0|1] [H] [W]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poisson_cnn_b200 import ops
a = [int(v) for v in sys.argv[1:]] + [None] * 8
B, mode, Cin, Cout, k, res, H, W = (a[0] or 16), (a[1] or 1), (a[2] or 32), (a[3] or 32), (a[4] or 15), (a[5] or 0), (a[6] or 256), (a[7] or 256)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, Cin, H, W, generator=g).cuda()
kern = (torch.randn(k, k, Cin, Cout, generator=g) / (k * Cin ** 0.5)).cuda()
t = ops.to_blk8(x, split=mode); wp = ops.pack_conv_weights_tc(kern, nsplit=mode); out = ops.Blk8(B, Cout, H, W, x.device, split=mode)
r = ops.to_blk8(torch.randn(B, Cout, H, W, generator=g).cuda(), split=mode) if res else None
bias = torch.zeros(Cout).cuda()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(3):
    if i == 2: e0.record()
    ops.conv2d_tc(t, wp, bias, 1, out=out, residual=r)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("mode %d B %d %d->%d k%d res %d %dx%d: %.3f ms, %.1f TFLOP/s algorithmic" % (mode, B, Cin, Cout, k, res, H, W, ms, 2.0 * B * H * W * k * k * Cin * Cout / ms / 1e9))
