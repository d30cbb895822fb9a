"""One dominant-layer launch set for ncu: 32->32 k15 tcgen05 conv, B=16, 256x256."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poisson_cnn_b200 import ops
g = torch.Generator().manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
x = torch.randn(B, 32, 256, 256, generator=g).cuda()
kern = (torch.randn(15, 15, 32, 32, generator=g) / (15 * 32 ** 0.5)).cuda()
t = ops.to_blk8(x); wp = ops.pack_conv_weights_tc(kern); out = ops.Blk8(B, 32, 256, 256, x.device)
for _ in range(3):
    ops.conv2d_tc(t, wp, torch.zeros(32).cuda(), 1, out=out)
torch.cuda.synchronize()
print("ok")
