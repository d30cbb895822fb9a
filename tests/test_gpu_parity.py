"""GPU parity tests (-m gpu): every CUDA entry point and the three models against the CPU oracle on
identical seeded inputs and weights, all called through the C ABI (poisson_cnn_b200.ops -> ctypes).

Tolerances (BASELINE.json north_star): strict-FP32 mode rel-L2 <= 1e-5 vs the float64 oracle.
"""
import os

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN, pcnn_configs, all_weights, rel_l2
from oracle import poisson_oracle as O

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
ACTS = {0: "linear", 1: "leaky_relu", 2: "tanh"}
PADS = {0: "CONSTANT", 1: "SYMMETRIC", 2: "REFLECT"}


def dev(t):
    return torch.as_tensor(t).float().cuda()


@pytest.fixture(scope="module")
def ops():
    from poisson_cnn_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("B,Cin,Cout,H,W,k,pad,act", [
    (2, 3, 4, 37, 70, 15, 1, 1),      # pre-bottleneck style, SYMMETRIC, ragged tile edges
    (2, 32, 32, 24, 40, 15, 0, 1),    # final stage 0
    (1, 32, 28, 19, 33, 13, 0, 1),
    (2, 29, 23, 20, 21, 7, 0, 2),     # DBCNN first 2-D conv, tanh
    (1, 64, 32, 16, 80, 7, 0, 1),     # post_merge_conv
    (2, 12, 8, 9, 130, 3, 0, 1),
    (3, 4, 1, 8, 8, 3, 0, 0),         # final linear 4 -> 1
    (2, 8, 8, 4, 4, 7, 1, 1),         # SYMMETRIC pad nearly as large as the tensor (pooled branches)
    (2, 5, 7, 12, 18, 4, 2, 0),       # even kernel + REFLECT
    (1, 6, 6, 2, 3, 5, 0, 1),         # tensor smaller than the kernel, constant pad value 2.0
])
def test_conv2d_parity(ops, B, Cin, Cout, H, W, k, pad, act):
    g = torch.Generator().manual_seed(B * 1000 + Cin * 10 + k)
    x = torch.randn(B, Cin, H, W, generator=g)
    kern = torch.randn(k, k, Cin, Cout, generator=g) / (k * Cin ** 0.5)
    bias = torch.randn(Cout, generator=g) * 0.1
    pv = 2.0 if (H < k and pad == 0) else 0.0
    ref = O.conv_nd(x.double(), kern.double(), bias.double(), ACTS[act], PADS[pad], pv)
    got = ops.conv2d(dev(x), dev(kern), dev(bias), act, pad, pv)
    assert rel_l2(got, ref) < FP32_TOL


def test_conv2d_fused_epilogue_and_strides(ops):
    """BN affine + residual + per-(b,c) scale, reading/writing channel slices of wider buffers."""
    g = torch.Generator().manual_seed(7)
    B, C, H, W, k = 2, 16, 21, 67, 5
    x = torch.randn(B, C, H, W, generator=g)
    res = torch.randn(B, C, H, W, generator=g)
    kern = torch.randn(k, k, C, C, generator=g) / (k * 4)
    bias = torch.randn(C, generator=g) * 0.1
    bn = {n: torch.rand(C, generator=g) + 0.5 for n in ("gamma", "var")}
    bn.update({n: torch.randn(C, generator=g) * 0.1 for n in ("beta", "mean")})
    scale = torch.randn(B, C, generator=g)
    ref = O.conv_nd(x.double(), kern.double(), bias.double(), "leaky_relu", "SYMMETRIC")
    ref = O.batchnorm(ref, {n: v.double() for n, v in bn.items()}) + res.double()
    ref = ref * scale.double().view(B, C, 1, 1)
    s = bn["gamma"] / torch.sqrt(bn["var"] + 1e-3)
    t = bn["beta"] - bn["mean"] * s
    xin = torch.zeros(B, 2 * C, H, W).cuda(); xin[:, C:] = dev(x)
    out = torch.zeros(B, 3 * C, H, W).cuda()
    ops.conv2d(xin[:, C:], dev(kern), dev(bias), 1, 1, 0.0, bn=(dev(s), dev(t)), residual=dev(res), out_scale=dev(scale), out=out[:, C:2 * C])
    assert rel_l2(out[:, C:2 * C], ref) < FP32_TOL
    assert float(out[:, :C].abs().max()) == 0 and float(out[:, 2 * C:].abs().max()) == 0


def test_conv1d_parity(ops):
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 3, 77, generator=g)
    for k, cin, cout in ((19, 3, 2), (5, 3, 27)):
        kern = torch.randn(k, cin, cout, generator=g) / k
        bias = torch.randn(cout, generator=g) * 0.1
        ref = O.conv_nd(x[:, :cin].double(), kern.double(), bias.double(), "leaky_relu", "SYMMETRIC")
        got = ops.conv1d(dev(x[:, :cin].contiguous()), dev(kern), dev(bias), 1, 1)
        assert rel_l2(got, ref) < FP32_TOL


@pytest.mark.parametrize("H,W,pad", [(8, 8, "CONSTANT"), (2, 2, "SYMMETRIC"), (4, 7, "SYMMETRIC"), (3, 5, "REFLECT")])
def test_smallmap_stack_parity(ops, H, W, pad):
    """Fused Conv2D stack on a tiny map (conv, then two resnets with BN) vs the oracle's layer-by-layer evaluation."""
    from poisson_cnn_b200.config import padding_enum
    g = torch.Generator().manual_seed(H * 10 + W)
    B, C, k = 3, 32, 5
    x = torch.randn(B, C, H, W, generator=g)
    layers, ref, saved = [], x.double(), None
    for has_bn, fl in [(False, 0)] + [(True, 1), (True, 2), (False, 0)] * 2:
        kern = torch.randn(k, k, C, C, generator=g) / (k * C ** 0.5)
        bias = torch.randn(C, generator=g) * 0.1
        bn = (torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1) if has_bn else None
        layers.append({"kernel": dev(kern), "bias": dev(bias), "bn": None if bn is None else (dev(bn[0]), dev(bn[1])), "flags": fl})
        if fl == 1:
            saved = ref
        ref = O.conv_nd(ref, kern.double(), bias.double(), "leaky_relu", pad, 0.0)
        if bn is not None:
            ref = ref * bn[0].double().view(1, -1, 1, 1) + bn[1].double().view(1, -1, 1, 1)
        if fl == 2:
            ref = ref + saved
    assert ops.smallmap_stack_supported(H, W, layers)
    got = ops.smallmap_stack(dev(x), layers, 1, padding_enum(pad), 0.0)
    assert got.shape == ref.shape
    assert rel_l2(got, ref) < FP32_TOL


@pytest.mark.parametrize("n,pad,act", [(256, "SYMMETRIC", "leaky_relu"), (50, "CONSTANT", "tanh"), (131, "REFLECT", "leaky_relu")])
def test_boundary_stack_parity(ops, n, pad, act):
    """Fused Conv1D stack (conv + BN, then a resnet, per stage) vs the oracle's layer-by-layer evaluation."""
    from poisson_cnn_b200.config import activation_enum, padding_enum
    g = torch.Generator().manual_seed(n)
    B, chans, ks = 3, [3, 4, 12, 27], [19, 9, 5]
    x = torch.randn(B, chans[0], n, generator=g)
    layers, ref = [], x.double()
    for cin, cout, k in zip(chans[:-1], chans[1:], ks):
        specs = [(cin, cout, True, 0), (cout, cout, True, 1), (cout, cout, True, 2), (cout, cout, False, 0)]
        saved = None
        for ci, co, has_bn, fl in specs:
            kern = torch.randn(k, ci, co, generator=g) / (k * ci) ** 0.5
            bias = torch.randn(co, generator=g) * 0.1
            bn = (torch.rand(co, generator=g) + 0.5, torch.randn(co, generator=g) * 0.1) if has_bn else None
            layers.append({"kernel": dev(kern), "bias": dev(bias), "bn": None if bn is None else (dev(bn[0]), dev(bn[1])), "flags": fl})
            if fl == 1:
                saved = ref
            ref = O.conv_nd(ref, kern.double(), bias.double(), act, pad, 0.25)
            if bn is not None:
                ref = ref * bn[0].double().view(1, -1, 1) + bn[1].double().view(1, -1, 1)
            if fl == 2:
                ref = ref + saved
    assert ops.boundary_stack_supported(n, layers)
    got = ops.boundary_stack(dev(x), layers, activation_enum(act), padding_enum(pad), 0.25)
    assert got.shape == ref.shape
    assert rel_l2(got, ref) < FP32_TOL


@pytest.mark.parametrize("H,W", [(200, 300), (64, 64), (109, 130)])
def test_avgpool_same_parity(ops, H, W):
    x = torch.randn(2, 5, H, W, generator=torch.Generator().manual_seed(H))
    for s in (2, 3, 4, 8, 16, 32, 64, 128):
        ref = O.avg_pool_same(x.double(), s)
        got = ops.avgpool_same(dev(x), s)
        assert got.shape == ref.shape
        assert rel_l2(got, ref) < FP32_TOL


@pytest.mark.parametrize("H,W", [(200, 300), (64, 64), (37, 50)])
def test_deconv_same_parity(ops, H, W):
    g = torch.Generator().manual_seed(W)
    for s in (2, 3, 4, 8, 16):
        ih, iw = -(-H // s), -(-W // s)
        x = torch.randn(2, 6, ih, iw, generator=g)
        kern = torch.randn(s, s, 5, 6, generator=g) / 3
        bias = torch.randn(5, generator=g) * 0.1
        ref = O.deconv_same(x.double(), kern.double(), bias.double(), "linear", (H, W), s)
        got = ops.deconv_same(dev(x), dev(kern), dev(bias), (H, W), s)
        assert rel_l2(got, ref) < FP32_TOL
        acc = torch.ones(2, 5, H, W).cuda()
        ops.deconv_same(dev(x), dev(kern), dev(bias), (H, W), s, alpha=0.25, out=acc, accumulate=True)
        assert rel_l2(acc, 1.0 + 0.25 * ref) < FP32_TOL
    with pytest.raises(ValueError):
        ops.deconv_same(dev(torch.randn(1, 6, 3, 3)), dev(torch.randn(2, 2, 5, 6)), None, (9, 9), 2)


@pytest.mark.parametrize("H,W,Cin,Cout", [(256, 256, 32, 32), (200, 300, 32, 32), (37, 50, 8, 5), (64, 64, 12, 32)])
def test_deconv_k_equals_stride_fast_path(ops, H, W, Cin, Cout):
    """k == stride with Cin % 4 == 0 takes the row-segment kernel (all column phases at once)."""
    g = torch.Generator().manual_seed(H + Cin)
    for s in (2, 3, 4, 8, 16):
        ih, iw = -(-H // s), -(-W // s)
        x = torch.randn(2, Cin, ih, iw, generator=g)
        kern = torch.randn(s, s, Cout, Cin, generator=g) / Cin ** 0.5
        bias = torch.randn(Cout, generator=g) * 0.1
        ref = O.deconv_same(x.double(), kern.double(), bias.double(), "leaky_relu", (H, W), s)
        got = ops.deconv_same(dev(x), dev(kern), dev(bias), (H, W), s, act=1)
        assert rel_l2(got, ref) < FP32_TOL, s
        acc = torch.ones(2, Cout, H, W).cuda()
        ops.deconv_same(dev(x), dev(kern), dev(bias), (H, W), s, act=1, alpha=0.25, out=acc, accumulate=True)
        assert rel_l2(acc, 1.0 + 0.25 * ref) < FP32_TOL, s


def test_resize_parity(ops):
    from poisson_cnn_b200.config import resize_enum
    g = torch.Generator().manual_seed(3)
    for (ih, iw), (oh, ow) in (((2, 3), (200, 300)), ((4, 5), (200, 300)), ((7, 10), (220, 317)), ((2, 2), (256, 256)), ((8, 8), (256, 256))):
        x = torch.randn(2, 4, ih, iw, generator=g)
        for m in ("nearest", "bilinear", "bicubic"):
            ref = O.resize(x.double(), (oh, ow), m)
            got = ops.resize(dev(x), (oh, ow), resize_enum(m))
            assert rel_l2(got, ref) < FP32_TOL, (m, ih, iw)


def test_spp_dense_maxabs_parity(ops):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 4, 10, 13, generator=g)
    lv = [[2, 2], 3, 5]
    assert rel_l2(ops.spatial_pyramid_pool(dev(x), lv, 1, 2), O.spatial_pyramid_pool(x.double(), lv, "max")) < 1e-7
    x1 = torch.randn(3, 27, 64, generator=g)
    lv1 = [2, 3, 4, 5, 8, 11, 15, 30, 45]
    assert rel_l2(ops.spatial_pyramid_pool(dev(x1), lv1, 0, 1), O.spatial_pyramid_pool(x1.double(), lv1, "avg")) < FP32_TOL
    e = ops.spatial_pyramid_pool(dev(torch.ones(1, 1, 3, 3)), [5], 1, 2)          # empty bins -> -inf like tf.reduce_max
    assert torch.isinf(e).any()
    v = torch.randn(5, 126, generator=g); k = torch.randn(126, 512, generator=g) / 11; b = torch.randn(512, generator=g)
    ref = O.dense(v.double(), {"kernel": k.double(), "bias": b.double()}, "tanh")
    assert rel_l2(ops.dense(dev(v), dev(k), dev(b), 2), ref) < FP32_TOL
    y = torch.randn(4, 1, 50, 60, generator=g) * torch.tensor([1.0, 0.01, 30.0, 2.0]).view(4, 1, 1, 1)
    m = ops.maxabs(dev(y))
    np.testing.assert_array_equal(m.cpu().numpy(), y.abs().amax(dim=(1, 2, 3)).numpy())
    assert rel_l2(ops.scale_inv(dev(y), m), O.set_max_magnitude(y.double())[0]) < 1e-6


def _device_model(kind, hp_cfg, db_cfg, weights):
    from poisson_cnn_b200 import convert_tf_object_names, models
    hp = models.Homogeneous_Poisson_NN_Legacy(**convert_tf_object_names(hp_cfg))
    db = models.Dirichlet_BC_NN_Legacy_2(**convert_tf_object_names(db_cfg))
    if kind == "hpnn":
        return hp.load_weights(weights, "hpnn/")
    if kind == "dbcnn":
        return db.load_weights(weights, "dbcnn/")
    return models.Poisson_CNN_Legacy(hp, db).load_weights(weights)


def test_hpnn_matches_golden_and_oracle():
    """BASELINE config 1 shape: HPNN forward, 64x64, zero Dirichlet ring (small-domain scaling block)."""
    hp, db = pcnn_configs(small_scaling=True)
    w = all_weights(hp, db)
    g = np.load(os.path.join(GOLDEN, "hpnn_64x64.npz"))
    model = _device_model("hpnn", hp, db, w)
    out = model([dev(g["rhs"]), dev(g["dx"])])
    assert out.shape == (2, 1, 64, 64)
    assert rel_l2(out, g["out"]) < FP32_TOL
    assert float(out[:, :, 0].abs().max()) == 0 and float(out[:, :, :, -1].abs().max()) == 0


def test_hpnn_neumann_and_nonsquare():
    from poisson_cnn_b200.synthetic import make_problem
    hp, db = pcnn_configs()
    hp = dict(hp, bc_type="neumann")
    w = all_weights(hp, db)
    p = make_problem(2, 112, 131, seed=21, magnitudes=False)
    ref = O.hpnn_forward(hp, w, p["rhs"].double(), p["dx"].double(), "hpnn/")
    out = _device_model("hpnn", hp, db, w)([dev(p["rhs"]), dev(p["dx"])])
    assert rel_l2(out, ref) < FP32_TOL


def test_dbcnn_matches_golden():
    hp, db = pcnn_configs(small_scaling=True)
    w = all_weights(hp, db)
    g = np.load(os.path.join(GOLDEN, "dbcnn_56x48.npz"))
    out = _device_model("dbcnn", hp, db, w)([dev(g["bc"]), dev(g["dx"]), 56])
    assert out.shape == (2, 1, 56, 48)
    assert rel_l2(out, g["out"]) < FP32_TOL
    np.testing.assert_array_equal(out[:, :, 0, :].cpu().numpy(), g["bc"])      # row 0 is the BC, bit exact


def test_pcnn_matches_golden():
    hp, db = pcnn_configs()
    w = all_weights(hp, db)
    g = np.load(os.path.join(GOLDEN, "pcnn_112x120.npz"))
    model = _device_model("pcnn", hp, db, w)
    out = model([dev(g[k]) for k in ("rhs", "left", "top", "right", "bottom", "dx")])
    assert rel_l2(out, g["out"]) < FP32_TOL


def test_pcnn_square_grid_batched_boundaries():
    """nx == ny takes the 4B-batched DBCNN path; compare with the oracle's four separate calls."""
    from poisson_cnn_b200.synthetic import make_problem
    hp, db = pcnn_configs()
    w = all_weights(hp, db)
    p = make_problem(2, 110, 110, seed=22)
    ref = O.pcnn_forward(hp, db, w, *(p[k].double() for k in ("rhs", "left", "top", "right", "bottom", "dx")))
    out = _device_model("pcnn", hp, db, w)([dev(p[k]) for k in ("rhs", "left", "top", "right", "bottom", "dx")])
    assert rel_l2(out, ref) < FP32_TOL
    # linearity in the boundary amplitude / homogeneity: scaling all inputs by c scales the output by c
    out2 = _device_model("pcnn", hp, db, w)([dev(p[k] * (3.0 if k != "dx" else 1.0)) for k in ("rhs", "left", "top", "right", "bottom", "dx")])
    assert rel_l2(out2, 3.0 * out) < 1e-5


def test_laplacian_residual_and_dst(ops):
    from poisson_cnn_b200.synthetic import make_problem
    from poisson_cnn_b200.losses import linear_operator_loss
    from poisson_cnn_b200.solvers import dst_poisson_solve
    p = make_problem(3, 70, 93, seed=31)
    gs = torch.cat([p["dx"], p["dx"] * 1.5], 1)
    sol = torch.randn(3, 1, 70, 93, generator=torch.Generator().manual_seed(1)) * 1e-3
    for st in (3, 5):
        ref = O.laplacian_residual(p["rhs"].double(), sol.double(), gs.double(), st)
        got = linear_operator_loss(st, 2, ndims=2)(dev(p["rhs"]), dev(sol), dev(gs))
        assert abs(float(got) - float(ref)) < 1e-4 * float(ref)
        refn = O.laplacian_residual(p["rhs"].double(), sol.double(), gs.double(), st, normalize=True, per_sample=True)
        gotn = linear_operator_loss(st, 2, ndims=2, normalize=True).per_sample_squared_sums(dev(p["rhs"]), dev(sol), dev(gs))
        assert rel_l2(gotn, refn) < 1e-4
    ref = O.dst_poisson_solve(p["rhs"], p["left"], p["top"], p["right"], p["bottom"], p["dx"])
    got = dst_poisson_solve(dev(p["rhs"]), {k: dev(p[k]) for k in ("left", "top", "right", "bottom")}, dev(p["dx"]))
    assert rel_l2(got, ref) < 1e-6
    # a direct solve has (near-)zero 3-point residual: the size-independent property used at full sizes
    r = linear_operator_loss(3, 2, ndims=2)(dev(p["rhs"]), got, dev(torch.cat([p["dx"], p["dx"]], 1)))
    assert float(r) < 1e-3 * float((p["rhs"] ** 2).mean())


@pytest.mark.parametrize("W", [37, 36, 4])          # scalar kernel / float4 kernel / float4 with one vector per row
def test_jacobi_parity(ops, W):
    from poisson_cnn_b200.synthetic import make_problem
    p = make_problem(2, 40, W, seed=41)
    gs = torch.cat([p["dx"], 1.5 * p["dx"]], 1)
    guess = torch.randn(2, 1, 40, W, generator=torch.Generator().manual_seed(2)) * 1e-2
    ref = O.jacobi_iterations(guess.double(), p["rhs"].double(), gs.double(), 3)
    got = ops.jacobi(dev(guess), dev(p["rhs"]), dev(gs), 3)
    assert rel_l2(got, ref) < FP32_TOL


def test_errors_are_loud(ops):
    with pytest.raises(ValueError):
        ops.conv2d(torch.zeros(1, 3, 8, 8), dev(torch.zeros(3, 3, 3, 4)))             # CPU tensor: no fallback
    with pytest.raises(ValueError):
        ops.conv2d(dev(torch.zeros(1, 3, 2, 2)), dev(torch.zeros(7, 7, 3, 4)), pad_mode=1)   # SYMMETRIC pad > tensor
    with pytest.raises(ValueError):
        ops.conv2d(dev(torch.zeros(1, 5, 8, 8)), dev(torch.zeros(3, 3, 3, 4)))         # channel mismatch
