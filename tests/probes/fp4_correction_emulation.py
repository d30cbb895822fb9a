"""CPU emulation (float64 oracle with rounded operands) of the HPNN trunk's compensated tensor-core product:
how accurate would the correction pass be in FP4 (`tcgen05.mma.kind::mxf4`, K = 64: 1.5x instead of 2x the single-pass
tensor work) compared with the shipped e4m3 pass (`tc2`)?

Every trunk convolution (pre_bottleneck, non_bottleneck_conv, post_merge_*, final/*) is evaluated as

    conv(h(x), h(W))  +  conv(Q(x), Q(W - h(W)))  +  conv(Q(x - h(x)), Q(W))          h = round to fp16

in float64 (no accumulation error: the tensor core's FP32 accumulation floor, 2.9e-4 for `tc2` on the GPU, comes on top),
with Q one of: nothing (single pass), e4m3, e2m1 with UNIT block scales (one power-of-two scale per tensor), e2m1 with
per-pixel / per-filter block scales over 32 channels (ue8m0, the mxf4 hardware format).  The bottleneck branches run
unmodified (they are single-pass in `mixed` and re-enter the trunk with weight 1/256).  Run: python tests/probes/fp4_correction_emulation.py
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from oracle import poisson_oracle as O
from tests.helpers import pcnn_configs, all_weights
from poisson_cnn_b200.synthetic import make_problem

E2M1 = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0], dtype=torch.float64)


def h16(v):
    return v.to(torch.float16).to(torch.float64)


def q_e2m1(v):
    """round to nearest e2m1 magnitude (ties to the even mantissa), saturating at 6."""
    a = v.abs().clamp(max=6.0)
    mid = (E2M1[1:] + E2M1[:-1]) / 2
    idx = torch.bucketize(a, mid, right=False)            # a == midpoint -> lower bucket ...
    tie = (idx < 7) & (a == mid[idx.clamp(max=6)])
    idx = torch.where(tie & (idx % 2 == 1), idx + 1, idx)  # ... unless the lower code has an odd mantissa
    return torch.sign(v) * E2M1[idx]


def q_e4m3(v):
    """e4m3 (bias 7, max 448, subnormals of 2^-9)."""
    a = v.abs().clamp(max=448.0)
    e = torch.floor(torch.log2(a.clamp(min=2.0 ** -20))).clamp(min=-6.0)
    step = torch.pow(2.0, e - 3)
    return torch.sign(v) * torch.round(a / step) * step


def pow2_at_least(x):
    return torch.pow(2.0, torch.ceil(torch.log2(x.clamp(min=1e-300))))


def quant(v, kind, block_dim=None):
    """kind: 'e4m3' | 'fp4_unit' | 'fp4_block'.  block_dim: the channel axis blocks of 32 run along."""
    if kind == "e4m3":
        s = pow2_at_least(v.abs().max() / 256.0)          # per-tensor power of two: the maximum lands in [128, 256]
        return q_e4m3(v / s) * s
    if kind == "fp4_unit":
        s = pow2_at_least(v.abs().max() / 6.0)
        return q_e2m1(v / s) * s
    if kind == "fp4_block":
        C = v.shape[block_dim]
        out = torch.empty_like(v)
        for c0 in range(0, C, 32):
            sl = [slice(None)] * v.dim()
            sl[block_dim] = slice(c0, min(c0 + 32, C))
            blk = v[tuple(sl)]
            s = pow2_at_least(blk.abs().amax(dim=block_dim, keepdim=True) / 6.0)
            out[tuple(sl)] = q_e2m1(blk / s) * s
        return out
    raise ValueError(kind)


MODE = {"kind": None, "active": True}
_plain_conv = O.conv_nd
_plain_block = O.bottleneck_block


def emulated_conv(x, kernel, bias, act, pad_mode="CONSTANT", pad_value=0.0):
    kernel = torch.as_tensor(kernel, dtype=x.dtype)
    if MODE["kind"] is None or not MODE["active"] or kernel.dim() != 4:
        return _plain_conv(x, kernel, bias, act, pad_mode, pad_value)
    kind = MODE["kind"]
    xp = O.advanced_pad(x, [kernel.shape[0], kernel.shape[1]], pad_mode, pad_value)
    w = kernel.permute(3, 2, 0, 1).contiguous()                            # [Cout, Cin, kh, kw]
    xh, wh = h16(xp), h16(w)
    y = F.conv2d(xh, wh)
    if kind != "single":
        xl, wl = xp - xh, w - wh
        y = y + F.conv2d(quant(xh, kind, 1), quant(wl, kind, 1)) + F.conv2d(quant(xl, kind, 1), quant(wh, kind, 1))
    if bias is not None:
        y = y + torch.as_tensor(bias, dtype=x.dtype).view(1, -1, 1, 1)
    return O.activation(y, act)


def plain_block(*a, **k):
    MODE["active"] = False
    try:
        return _plain_block(*a, **k)
    finally:
        MODE["active"] = True


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    hp, db = pcnn_configs()
    w = all_weights(hp, db)
    O.conv_nd = emulated_conv
    O.bottleneck_block = plain_block
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    for seed in (1001, 7):
        p = make_problem(2, n, n, seed=seed)
        rhs, dx = p["rhs"].double(), p["dx"].double()
        rhs = rhs / rhs.abs().amax(dim=(1, 2, 3), keepdim=True)
        with torch.no_grad():
            MODE["kind"] = None
            ref = O.hpnn_forward(hp, w, rhs, dx, "hpnn/")
            for kind in ("single", "e4m3", "fp4_block", "fp4_unit"):
                MODE["kind"] = kind
                out = O.hpnn_forward(hp, w, rhs, dx, "hpnn/")
                e = ((out - ref).flatten(1).norm(dim=1) / ref.flatten(1).norm(dim=1)).tolist()
                print("seed %d  %dx%d  trunk correction %-9s  rel-L2 per sample: %s" % (seed, n, n, kind, " ".join("%.2e" % v for v in e)), flush=True)


if __name__ == "__main__":
    main()
