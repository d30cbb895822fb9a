"""One dominant-layer launch set for ncu: 32->32 k15 tcgen05 conv, 256x256.  usage: profile_conv_tc.py [B] [mode 1|2|3]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poisson_cnn_b200 import ops
g = torch.Generator().manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
x = torch.randn(B, 32, 256, 256, generator=g).cuda()
kern = (torch.randn(15, 15, 32, 32, generator=g) / (15 * 32 ** 0.5)).cuda()
t = ops.to_blk8(x, split=mode); wp = ops.pack_conv_weights_tc(kern, nsplit=mode); out = ops.Blk8(B, 32, 256, 256, x.device, split=mode)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(3):
    if i == 2: e0.record()
    ops.conv2d_tc(t, wp, torch.zeros(32).cuda(), 1, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("mode %d B %d: %.3f ms, %.1f TFLOP/s algorithmic" % (mode, B, ms, 2.0 * B * 65536 * 225 * 1024 / ms / 1e9))
