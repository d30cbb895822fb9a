"""GPU tests of the tensor-core path (tcgen05 conv on BLK8 fp16 tensors) against the oracle.
Operands are rounded to fp16 (11-bit significand, the same as TF32) and accumulated in fp32, so a
single conv matches the float64 oracle to ~3e-4 relative; the oracle evaluated on fp16-rounded
inputs/weights matches to ~1e-5 (accumulation order only)."""
import numpy as np
import pytest
import torch

from tests.helpers import rel_l2
from oracle import poisson_oracle as O

pytestmark = pytest.mark.gpu
ACTS = {0: "linear", 1: "leaky_relu", 2: "tanh"}


def dev(t):
    return torch.as_tensor(t).float().cuda()


def h16(t):
    return t.half().double()


@pytest.fixture(scope="module")
def ops():
    from poisson_cnn_b200 import ops as _ops
    return _ops


def test_blk8_roundtrip_and_halo(ops):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 13, 9, 21, generator=g)
    t = ops.to_blk8(dev(x))
    back = ops.from_blk8(t)
    assert back.shape == x.shape
    np.testing.assert_array_equal(back.cpu().numpy(), x.half().float().numpy())
    # raw buffer: halo zero, SYMMETRIC fill mirrors the interior
    Hp, P = 9 + 14, 21 + 14
    raw = lambda: t.buf[: 2 * 2 * Hp * P * 8].view(2, 2, Hp, P, 8).float().cpu()
    r0 = raw()
    assert float(r0[:, :, :7].abs().max()) == 0 and float(r0[:, :, :, :7].abs().max()) == 0
    ops.blk8_halo_fill(t, 5, 1)
    r1 = raw()
    ref = O.advanced_pad(x.half().float(), [11, 11], "SYMMETRIC")              # pad 5 each side
    got = r1[:, :, 2:2 + 19, 2:2 + 31].permute(0, 1, 4, 2, 3).reshape(2, 16, 19, 31)[:, :13]
    np.testing.assert_array_equal(got.numpy(), ref.numpy())


@pytest.mark.parametrize("B,Cin,Cout,H,W,k,act", [
    (2, 32, 32, 8, 256, 15, 1),       # the dominant layer shape: full 256-wide tiles
    (1, 32, 32, 7, 300, 15, 1),       # ragged: H not a multiple of 4, two column tiles
    (2, 32, 28, 12, 64, 13, 1),
    (1, 64, 32, 9, 80, 7, 1),         # post_merge_conv: 4 K-chunks
    (2, 29, 23, 20, 21, 7, 2),        # DBCNN first 2-D conv, odd channel counts, tanh
    (1, 16, 16, 16, 40, 5, 0),
    (3, 12, 8, 5, 33, 3, 1),
    (1, 32, 32, 4, 16, 1, 0),         # 1x1
    (2, 29, 23, 33, 300, 7, 1),       # 17 <= Cout <= 24: 5 output rows x 24 channel slots (120 of 128 accumulator rows), odd row count
    (1, 24, 24, 12, 256, 9, 1),
    (2, 23, 19, 9, 40, 7, 2),
    (2, 15, 15, 37, 300, 5, 1),       # Cout <= 16: 8 output rows x 16 channel slots per tile, ragged both ways
    (2, 19, 15, 8, 256, 5, 2),
    (2, 7, 5, 37, 70, 3, 2),          # Cout <= 8: 16 output rows x 8 channel slots
    (1, 5, 5, 50, 256, 3, 1),
    (1, 11, 7, 16, 31, 5, 0),
])
def test_conv2d_tc_parity(ops, B, Cin, Cout, H, W, k, act):
    g = torch.Generator().manual_seed(B * 1000 + Cin * 10 + k)
    x = torch.randn(B, Cin, H, W, generator=g)
    kern = torch.randn(k, k, Cin, Cout, generator=g) / (k * Cin ** 0.5)
    bias = torch.randn(Cout, generator=g) * 0.1
    wp = ops.pack_conv_weights_tc(dev(kern))
    out = ops.conv2d_tc(ops.to_blk8(dev(x)), wp, dev(bias), act)
    got = ops.from_blk8(out)
    assert got.shape == (B, Cout, H, W)
    ref16 = O.conv_nd(h16(x), h16(kern), bias.double(), ACTS[act], "CONSTANT", 0.0)    # same operand rounding
    assert rel_l2(got, ref16) < 6e-4      # fp16 output rounding (2^-11) + accumulation order
    ref = O.conv_nd(x.double(), kern.double(), bias.double(), ACTS[act], "CONSTANT", 0.0)
    assert rel_l2(got, ref) < 2e-3


def test_conv2d_tc_symmetric_bn_residual_scale_concat(ops):
    g = torch.Generator().manual_seed(9)
    B, C, H, W, k = 2, 32, 18, 50, 11
    x = torch.randn(B, C, H, W, generator=g)
    res = torch.randn(B, C, H, W, generator=g)
    kern = torch.randn(k, k, C, C, generator=g) / (k * C ** 0.5)
    bias = torch.randn(C, generator=g) * 0.1
    s = torch.rand(C, generator=g) + 0.5
    t = torch.randn(C, generator=g) * 0.1
    scale = torch.randn(B, C, generator=g)
    ref = O.conv_nd(h16(x), h16(kern), bias.double(), "leaky_relu", "SYMMETRIC")
    ref = (ref * s.double().view(1, C, 1, 1) + t.double().view(1, C, 1, 1)) * scale.double().view(B, C, 1, 1)
    ref = h16(ref.float()) + h16(res)
    wp = ops.pack_conv_weights_tc(dev(kern))
    # write into channels [32, 64) of a 64-channel buffer (in-place concat), read residual from another
    out = ops.Blk8(B, 64, H, W, torch.device("cuda"))
    xin = ops.to_blk8(dev(x))
    rin = ops.to_blk8(dev(res))
    tmp = ops.conv2d_tc(xin, wp, dev(bias), 1, pad_mode=1, bn=(dev(s), dev(t)), residual=rin, out_scale=dev(scale))
    assert rel_l2(ops.from_blk8(tmp), ref) < 1e-3
    assert xin.halo == (1, 5)
    # a CONSTANT consumer after a SYMMETRIC one must see zeros again
    ref0 = O.conv_nd(h16(x), h16(kern), bias.double(), "linear", "CONSTANT")
    again = ops.conv2d_tc(xin, wp, dev(bias), 0, pad_mode=0)
    assert rel_l2(ops.from_blk8(again), ref0) < 6e-4


@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("B,Cin,Cout,H,W,k", [(2, 32, 32, 18, 50, 11), (1, 16, 12, 40, 270, 5), (2, 8, 5, 33, 9, 3), (1, 24, 20, 23, 60, 7)])
def test_fused_symmetric_halo_equals_halo_fill(ops, mode, B, Cin, Cout, H, W, k):
    """A producer asked for out_halo=SYMMETRIC leaves exactly the buffer that pcnn_blk8_halo_fill would make."""
    g = torch.Generator().manual_seed(H * W + mode)
    x = torch.randn(B, Cin, H, W, generator=g)
    kern = torch.randn(k, k, Cin, Cout, generator=g) / (k * Cin ** 0.5)
    wp = ops.pack_conv_weights_tc(dev(kern), nsplit=mode)
    xin = ops.to_blk8(dev(x), split=mode, halo=1)
    assert xin.halo == (1, 7)
    ref_in = ops.to_blk8(dev(x), split=mode)
    ops.blk8_halo_fill(ref_in, 7, 1)
    assert torch.equal(xin.buf, ref_in.buf) and (mode == 1 or torch.equal(xin.lo, ref_in.lo))
    a = ops.conv2d_tc(xin, wp, None, 1, pad_mode=1, out_halo=1)
    assert a.halo == (1, 7)
    b = ops.conv2d_tc(ref_in, wp, None, 1, pad_mode=1)
    ops.blk8_halo_fill(b, 7, 1)
    assert torch.equal(a.buf, b.buf) and (mode == 1 or torch.equal(a.lo, b.lo))


def test_conv2d_tc_chain_matches_fp32_path(ops):
    """Three chained TC convs stay in BLK8 (no re-layout) and track the strict-FP32 kernels to fp16 accuracy."""
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 32, 24, 70, generator=g)
    kerns = [torch.randn(7, 7, 32, 32, generator=g) / (7 * 32 ** 0.5) * 1.4 for _ in range(3)]
    biases = [torch.randn(32, generator=g) * 0.1 for _ in range(3)]
    a = dev(x)
    t = ops.to_blk8(a)
    for kern, bias in zip(kerns, biases):
        a = ops.conv2d(a, dev(kern), dev(bias), 1, 0, 0.0)
        t = ops.conv2d_tc(t, ops.pack_conv_weights_tc(dev(kern)), dev(bias), 1)
    assert rel_l2(ops.from_blk8(t), a) < 2e-3


@pytest.mark.parametrize("H,W,C,mode", [(256, 256, 32, 3), (200, 300, 32, 1), (37, 50, 8, 3), (64, 80, 24, 2)])
def test_upsample_merge_blk8(ops, H, W, C, mode):
    """Fused deconv (k == stride) + resize branch sum written straight into a BLK8 concat buffer."""
    from poisson_cnn_b200.config import resize_enum
    g = torch.Generator().manual_seed(H + C)
    B = 2
    dc, rs, ref = [], [], 0.0
    for s in (16, 8, 4, 3, 2):
        ih, iw = -(-H // s), -(-W // s)
        x = torch.randn(B, C, ih, iw, generator=g)
        kern = torch.randn(s, s, C, C, generator=g) / C ** 0.5
        bias = torch.randn(C, generator=g) * 0.1
        dc.append((dev(x), ops.pack_deconv_kernel(dev(kern)), dev(bias), s, 1))
        ref = ref + O.deconv_same(x.double(), kern.double(), bias.double(), "leaky_relu", (H, W), s)
    for (ih, iw), m in (((2, 2), "bilinear"), ((4, 5), "bicubic"), ((8, 8), "nearest")):
        x = torch.randn(B, C, ih, iw, generator=g)
        rs.append((dev(x), resize_enum(m)))
        ref = ref + O.resize(x.double(), (H, W), m)
    alpha = 1.0 / 8
    ref = ref * alpha
    out = ops.Blk8(B, 2 * C if C % 16 == 0 else 16 + C, H, W, torch.device("cuda"), split=mode)
    c_off = C if C % 16 == 0 else 16
    ops.upsample_merge_blk8(dc, rs, alpha, out, c_off, H, W)
    got = ops.from_blk8(out, C=C, c_offset=c_off)
    tol = {1: 6e-4, 2: 2e-6, 3: 4e-5}[mode]          # storage precision of the destination
    assert rel_l2(got, ref) < tol
    assert float(ops.from_blk8(out, C=c_off, c_offset=0).abs().max()) == 0.0     # neighbouring channels untouched


@pytest.mark.parametrize("B,Cin,Cout,H,W,k,act", [(3, 29, 23, 37, 300, 7, 2), (2, 29, 23, 256, 256, 7, 1), (2, 16, 12, 19, 40, 5, 0),
                                                  (1, 32, 32, 8, 64, 3, 1), (2, 5, 4, 33, 20, 3, 2)])
def test_conv2d_tc_rowweights_separable_input(ops, B, Cin, Cout, H, W, k, act):
    """A convolution whose input is h[b,m,y] * S[m,x] with the row taps folded into per-row weights
    (pcnn_conv2d_tc_rowweights): equals the full convolution of the expanded tensor (zero padding)."""
    g = torch.Generator().manual_seed(B * 100 + Cin + H)
    sig = torch.randn(B, Cin, W, generator=g)
    basis = torch.randn(Cin, H, generator=g)
    x = sig[:, :, None, :] * basis[None, :, :, None]
    kern = torch.randn(k, k, Cin, Cout, generator=g) / (k * Cin ** 0.5)
    bias = torch.randn(Cout, generator=g) * 0.1
    wp = ops.pack_rowweights_tc(dev(kern), dev(basis))
    xr = ops.to_blk8(dev(sig.view(B, Cin, 1, W).contiguous()))
    got = ops.from_blk8(ops.conv2d_tc_rowweights(xr, wp, dev(bias), act))
    assert got.shape == (B, Cout, H, W)
    ref = O.conv_nd(x.double(), kern.double(), bias.double(), ACTS[act], "CONSTANT", 0.0)
    assert rel_l2(got, ref) < 1.5e-3          # fp16 operands (row weights and signals), fp16 output
    full = ops.from_blk8(ops.conv2d_tc(ops.to_blk8(dev(x)), ops.pack_conv_weights_tc(dev(kern)), dev(bias), act))
    assert rel_l2(got, full) < 2e-3           # the k x k convolution of the expanded tensor on the same kernel family


@pytest.mark.parametrize("H,W,mode,B", [(256, 256, 3, 2), (200, 300, 1, 3), (37, 50, 3, 2), (64, 80, 2, 9), (112, 120, 3, 17)])
def test_upsample_merge_tc_blk8(ops, H, W, mode, B):
    """Tensor-core version: the transpose convolutions run as mma.sync products from BLK8 fp16 branch outputs.  Same
    operator as test_upsample_merge_blk8; the reference sees the fp16-rounded inputs and kernels the tensor cores see."""
    from poisson_cnn_b200.config import resize_enum
    g = torch.Generator().manual_seed(H + W + B)
    C = 32
    dc, rs, ref, parts = [], [], 0.0, []
    for s, act in ((16, 1), (8, 1), (4, 2), (3, 0), (2, 1)):
        ih, iw = -(-H // s), -(-W // s)
        x = torch.randn(B, C, ih, iw, generator=g)
        kern = torch.randn(s, s, C, C, generator=g) / C ** 0.5
        bias = torch.randn(C, generator=g) * 0.1
        dc.append((ops.to_blk8(dev(x)), ops.pack_deconv_kernel_tc(dev(kern)), dev(bias) if s != 3 else None, s, act))
        parts.append(O.deconv_same(h16(x), h16(kern), bias.double() if s != 3 else torch.zeros(C, dtype=torch.float64), ACTS[act], (H, W), s))
        ref = ref + parts[-1]
    for (ih, iw), m in (((2, 2), "bilinear"), ((4, 5), "bicubic"), ((8, 8), "nearest")):
        x = torch.randn(B, C, ih, iw, generator=g)
        rs.append((dev(x), resize_enum(m)))
        ref = ref + O.resize(x.double(), (H, W), m)
    alpha = 1.0 / 8
    ref = ref * alpha
    out = ops.Blk8(B, 2 * C, H, W, torch.device("cuda"), split=mode)
    ops.upsample_merge_tc_blk8(dc, rs, alpha, out, C, H, W)
    got = ops.from_blk8(out, C=C, c_offset=C)
    assert bool(torch.isfinite(got).all())
    tol = {1: 6e-4, 2: 3e-6, 3: 4e-5}[mode]          # storage precision of the destination
    assert rel_l2(got, ref) < tol
    assert float(ops.from_blk8(out, C=C, c_offset=0).abs().max()) == 0.0     # neighbouring channels untouched
    if mode == 1:                                     # two branches, no resize branches, whole tensor
        out2 = ops.Blk8(B, C, H, W, torch.device("cuda"))
        ops.upsample_merge_tc_blk8(dc[:2], [], 1.0, out2, 0, H, W)
        assert rel_l2(ops.from_blk8(out2), parts[0] + parts[1]) < tol


# ------------------------------------------------------------------ whole models in tensor-core mode
def _models(hp_cfg, db_cfg, w):
    from poisson_cnn_b200 import convert_tf_object_names, models
    hp = models.Homogeneous_Poisson_NN_Legacy(**convert_tf_object_names(hp_cfg))
    db = models.Dirichlet_BC_NN_Legacy_2(**convert_tf_object_names(db_cfg))
    return models.Poisson_CNN_Legacy(hp, db).load_weights(w).set_precision("tc")


# Tensor-core mode budget (BASELINE.json north_star): rel-L2 <= 2e-3 vs the reference arithmetic.
TC_TOL = 2e-3


def test_models_tc_mode_vs_oracle():
    import os
    from tests.helpers import GOLDEN, pcnn_configs, all_weights
    from poisson_cnn_b200.synthetic import make_problem
    hp, db = pcnn_configs()
    w = all_weights(hp, db)
    model = _models(hp, db, w)
    keys = ("rhs", "left", "top", "right", "bottom", "dx")
    g = np.load(os.path.join(GOLDEN, "pcnn_112x120.npz"))
    out = model([dev(g[k]) for k in keys])
    e_pcnn = rel_l2(out, g["out"])
    fp32 = model.set_precision("fp32")([dev(g[k]) for k in keys])
    model.set_precision("tc")
    print("PCNN 112x120 tc vs oracle: %.3e   (tc vs fp32 path: %.3e)" % (e_pcnn, rel_l2(out, fp32)))
    # DBCNN alone, square grid
    p = make_problem(2, 96, 96, seed=51, magnitudes=False)
    ref = O.dbcnn_forward(db, w, p["left"].double(), p["dx"].double(), 96, "dbcnn/")
    e_db = rel_l2(model.dbcnn([dev(p["left"]), dev(p["dx"]), 96]), ref)
    # HPNN alone
    p = make_problem(2, 128, 112, seed=52, magnitudes=False)
    ref = O.hpnn_forward(hp, w, p["rhs"].double(), p["dx"].double(), "hpnn/")
    e_hp = rel_l2(model.hpnn([dev(p["rhs"]), dev(p["dx"])]), ref)
    print("tc-mode rel-L2 vs float64 oracle: pcnn %.3e  dbcnn %.3e  hpnn %.3e" % (e_pcnn, e_db, e_hp))
    assert e_pcnn < TC_TOL and e_db < TC_TOL
    # The single-pass mode is NOT a compliant mode for the 45-convolution-deep HPNN (4.4e-3 at 256x256 against the 2e-3
    # budget): the class says so (COMPLIANT_PRECISIONS) and selecting it warns.  What is asserted here
    # is only that the kernels work (sane error); compliance is carried by tc2 / tc3 / mixed (tests below, test_gpu_round2).
    assert "tc" not in model.hpnn.COMPLIANT_PRECISIONS and "tc" in model.dbcnn.COMPLIANT_PRECISIONS
    assert e_hp < 1e-2
    with pytest.warns(UserWarning, match="(?i)outside the 2e-3"):
        model.hpnn.set_precision("tc")


# ------------------------------------------------------------------ split precision (tc3)
@pytest.mark.parametrize("B,Cin,Cout,H,W,k,act", [
    (2, 32, 32, 8, 256, 15, 1),
    (1, 64, 32, 9, 80, 7, 1),
    (2, 29, 23, 20, 21, 7, 2),
    (3, 12, 8, 5, 33, 3, 0),
    (2, 15, 15, 37, 300, 5, 1),
    (2, 7, 5, 37, 70, 3, 2),
    (1, 11, 7, 16, 31, 5, 1),
    (2, 28, 24, 33, 300, 9, 1),
    (2, 24, 20, 9, 40, 7, 2),
])
def test_conv2d_tc3_parity(ops, B, Cin, Cout, H, W, k, act):
    """hi/lo split operands, three MMAs: ~22 operand bits.  Measured floor ~8e-6 on the K=7200 layer with or
    without weight pre-scaling: it is the tensor core's FP32 accumulation (not round-to-nearest over the
    450-MMA chain), not the operand split -- still 50x tighter than a single FP16 pass."""
    g = torch.Generator().manual_seed(B * 1000 + Cin * 10 + k + 1)
    x = torch.randn(B, Cin, H, W, generator=g)
    kern = torch.randn(k, k, Cin, Cout, generator=g) / (k * Cin ** 0.5)
    bias = torch.randn(Cout, generator=g) * 0.1
    res = torch.randn(B, Cout, H, W, generator=g)
    wp = ops.pack_conv_weights_tc(dev(kern), nsplit=2)
    out = ops.conv2d_tc(ops.to_blk8(dev(x), split=True), wp, dev(bias), act, residual=ops.to_blk8(dev(res), split=True))
    got = ops.from_blk8(out)
    ref = O.conv_nd(x.double(), kern.double(), bias.double(), ACTS[act], "CONSTANT", 0.0) + res.double()
    assert rel_l2(got, ref) < 2e-5


def test_models_tc3_mode_vs_oracle():
    import os
    from tests.helpers import GOLDEN, pcnn_configs, all_weights
    from poisson_cnn_b200.synthetic import make_problem
    hp, db = pcnn_configs()
    w = all_weights(hp, db)
    model = _models(hp, db, w).set_precision("tc3")
    keys = ("rhs", "left", "top", "right", "bottom", "dx")
    g = np.load(os.path.join(GOLDEN, "pcnn_112x120.npz"))
    e_pcnn = rel_l2(model([dev(g[k]) for k in keys]), g["out"])
    p = make_problem(2, 128, 112, seed=52, magnitudes=False)
    ref = O.hpnn_forward(hp, w, p["rhs"].double(), p["dx"].double(), "hpnn/")
    e_hp = rel_l2(model.hpnn([dev(p["rhs"]), dev(p["dx"])]), ref)
    print("tc3-mode rel-L2 vs float64 oracle: pcnn %.3e  hpnn %.3e" % (e_pcnn, e_hp))
    assert e_pcnn < 5e-4 and e_hp < 5e-4      # >= 4x inside the 2e-3 tensor-core budget (floor: tensor-core fp32 accumulation)


# ------------------------------------------------------------------ fp8-corrected mode (tc2)
@pytest.mark.parametrize("B,Cin,Cout,H,W,k,act", [
    (2, 32, 32, 8, 256, 15, 1),
    (1, 64, 32, 9, 80, 7, 1),
    (2, 29, 23, 20, 21, 7, 2),
    (3, 12, 8, 5, 33, 3, 0),
    (2, 15, 15, 37, 300, 5, 1),
    (2, 7, 5, 37, 70, 3, 2),
    (1, 11, 7, 16, 31, 5, 1),
    (2, 28, 24, 33, 300, 9, 1),
    (2, 24, 20, 9, 40, 7, 2),
])
def test_conv2d_tc2_parity(ops, B, Cin, Cout, H, W, k, act):
    """fp16 main MMA + one e4m3 K=32 MMA carrying both correction terms: ~10x tighter than a single fp16 pass."""
    g = torch.Generator().manual_seed(B * 1000 + Cin * 10 + k + 2)
    x = torch.randn(B, Cin, H, W, generator=g)
    kern = torch.randn(k, k, Cin, Cout, generator=g) / (k * Cin ** 0.5)
    bias = torch.randn(Cout, generator=g) * 0.1
    res = torch.randn(B, Cout, H, W, generator=g)
    xin = ops.to_blk8(dev(x), split=3)
    assert rel_l2(ops.from_blk8(xin), x) < 4e-5                      # hi + e4m3 remainder: ~15 bits
    wp = ops.pack_conv_weights_tc(dev(kern), nsplit=3)
    out = ops.conv2d_tc(xin, wp, dev(bias), act, residual=ops.to_blk8(dev(res), split=3))
    got = ops.from_blk8(out)
    ref = O.conv_nd(x.double(), kern.double(), bias.double(), ACTS[act], "CONSTANT", 0.0) + res.double()
    err = rel_l2(got, ref)
    print("tc2 conv k%d %d->%d rel-L2 %.2e" % (k, Cin, Cout, err))
    assert err < 3e-5


def test_models_tc2_mode_vs_oracle():
    import os
    from tests.helpers import GOLDEN, pcnn_configs, all_weights
    from poisson_cnn_b200.synthetic import make_problem
    hp, db = pcnn_configs()
    w = all_weights(hp, db)
    model = _models(hp, db, w).set_precision("tc2")
    keys = ("rhs", "left", "top", "right", "bottom", "dx")
    g = np.load(os.path.join(GOLDEN, "pcnn_112x120.npz"))
    e_pcnn = rel_l2(model([dev(g[k]) for k in keys]), g["out"])
    p = make_problem(2, 128, 112, seed=52, magnitudes=False)
    ref = O.hpnn_forward(hp, w, p["rhs"].double(), p["dx"].double(), "hpnn/")
    e_hp = rel_l2(model.hpnn([dev(p["rhs"]), dev(p["dx"])]), ref)
    print("tc2-mode rel-L2 vs float64 oracle: pcnn %.3e  hpnn %.3e" % (e_pcnn, e_hp))
    assert e_pcnn < 5e-4 and e_hp < 5e-4      # >= 4x inside the 2e-3 tensor-core budget


def test_models_mixed_mode_vs_oracle():
    """precision='mixed': tc2 in the HPNN, single-pass tc in the DBCNN -- the accuracy of tc2 at ~0.8x the tensor work."""
    import os
    from tests.helpers import GOLDEN, pcnn_configs, all_weights
    from poisson_cnn_b200.synthetic import make_problem
    hp, db = pcnn_configs()
    w = all_weights(hp, db)
    model = _models(hp, db, w).set_precision("mixed")
    assert (model.precision, model.hpnn.precision, model.dbcnn.precision) == ("mixed", "tc2", "tc")
    keys = ("rhs", "left", "top", "right", "bottom", "dx")
    g = np.load(os.path.join(GOLDEN, "pcnn_112x120.npz"))
    e_pcnn = rel_l2(model([dev(g[k]) for k in keys]), g["out"])
    p = make_problem(2, 160, 144, seed=53)
    ref = O.pcnn_forward(hp, db, w, *[p[k].double() for k in keys])
    e2 = rel_l2(model([dev(p[k]) for k in keys]), ref)
    print("mixed-mode rel-L2 vs float64 oracle: pcnn golden %.3e  160x144 %.3e" % (e_pcnn, e2))
    # inside the 2e-3 tensor-core budget with margin; the exact value is a noise draw (a 1-ulp change of an input or of a
    # host table moves it by its own size: 4.8e-4 with numpy's cos table, 6.0e-4 with the library's correctly rounded one)
    assert e_pcnn < 1e-3 and e2 < 1e-3
