"""CPU tests: pin the oracle against everything the reference offers for this path
 and against the committed golden vectors."""
import json
import os

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN, pcnn_configs, all_weights, rel_l2
from oracle import poisson_oracle as O


def test_fd_coefficients_match_reference_fixture():
    """Fixture produced by executing the reference's get_fd_coefficients (make_reference_fixtures.py)."""
    ref = json.load(open(os.path.join(GOLDEN, "reference_fd_coefficients.json")))
    for case in ref["cases"]:
        got = O.fd_coefficients(case["positions"], case["order"])
        np.testing.assert_allclose(got, case["coefficients"], rtol=1e-10, atol=1e-12)


def test_fd_stencils_known_answers():
    np.testing.assert_allclose(O.fd_coefficients([-1, 0, 1], 2), [1, -2, 1], atol=1e-12)
    np.testing.assert_allclose(O.fd_coefficients([-2, -1, 0, 1, 2], 2), [-1 / 12, 4 / 3, -5 / 2, 4 / 3, -1 / 12], atol=1e-12)
    c = O.build_fd_coefficients(3, 2, 2)
    assert c.shape == (2, 3, 3)
    np.testing.assert_allclose(c[0], [[0, 1, 0], [0, -2, 0], [0, 1, 0]], atol=1e-12)
    np.testing.assert_allclose(c[1], [[0, 0, 0], [1, -2, 1], [0, 0, 0]], atol=1e-12)


def test_split_indices_docstring_example():
    # poisson_CNN/dataset/utils/split_indices.py:13
    assert list(O.split_indices(229, 4)) == [0, 58, 115, 172, 229]
    for n, s in [(10, 3), (45, 45), (7, 2), (123, 11)]:
        sizes = [len(a) for a in np.array_split(np.arange(n), s)]
        assert list(np.diff(O.split_indices(n, s))) == sizes


def test_flip_and_rotate_matches_rot90():
    # poisson_CNN/dataset/utils/flip_and_rotate_tensor.py:49-61 compares with tf.image.rot90
    r = (torch.arange(5).view(5, 1) + 10 * torch.arange(4).view(1, 4)).double().view(1, 1, 5, 4)
    ref = np.asarray(r[0, 0])
    for k in range(4):
        np.testing.assert_array_equal(np.asarray(O.flip_and_rotate(r, k)[0, 0]), np.rot90(ref, k))
    # bottom boundary: rot90 k=1 then flip axis 2 is a plain transpose; top: out[i,j] = r[ny-1-j, i]
    np.testing.assert_array_equal(np.asarray(O.flip_and_rotate(r, 1, [2])[0, 0]), ref.T)
    top = np.asarray(O.flip_and_rotate(r, 3)[0, 0])
    for i in range(4):
        for j in range(5):
            assert top[i, j] == ref[5 - 1 - j, i]


def test_pad_modes_match_numpy():
    x = torch.arange(12.0).view(1, 1, 3, 4)
    for k in (3, 5, 4, 2):
        l, r = k // 2, k // 2 - (1 - k % 2)
        for mode, npmode in (("SYMMETRIC", "symmetric"), ("REFLECT", "reflect")):
            if mode == "REFLECT" and l >= 3:
                continue
            got = O.advanced_pad(x, [k, k], mode)[0, 0].numpy()
            np.testing.assert_array_equal(got, np.pad(x[0, 0].numpy(), ((l, r), (l, r)), mode=npmode))
        got = O.advanced_pad(x, [k, k], "CONSTANT", 2.0)[0, 0].numpy()
        np.testing.assert_array_equal(got, np.pad(x[0, 0].numpy(), ((l, r), (l, r)), constant_values=2.0))


def test_avg_pool_same_excludes_padding():
    x = torch.ones(1, 1, 5, 7, dtype=torch.float64)
    for s in (2, 3, 4, 8):
        out = O.avg_pool_same(x, s)
        assert out.shape[2:] == (-(-5 // s), -(-7 // s))
        np.testing.assert_allclose(out.numpy(), 1.0, atol=1e-14)      # valid-count divisor
    x = torch.arange(5.0, dtype=torch.float64).view(1, 1, 1, 5)
    # W=5, s=2: out=3, pad_total=1, pad_before=0 -> windows [0,1],[2,3],[4]
    np.testing.assert_allclose(O.avg_pool_same(x, 2)[0, 0, 0].numpy(), [0.5, 2.5, 4.0])
    # W=5, s=4: out=2, pad_total=3, pad_before=1 -> windows [0,1,2],[3,4]
    np.testing.assert_allclose(O.avg_pool_same(x, 4)[0, 0, 0].numpy(), [1.0, 3.5])


def test_deconv_same_is_adjoint_of_same_conv():
    """conv2d_transpose(SAME) is the adjoint of the SAME strided conv: <conv(x), y> == <x, deconv(y)>."""
    torch.manual_seed(0)
    for N, s in ((10, 2), (11, 3), (13, 4), (9, 8)):
        k = s
        x = torch.randn(1, 2, N, N, dtype=torch.float64)
        wt = torch.randn(k, k, 2, 3, dtype=torch.float64)           # conv2d_transpose layout [kh,kw,Cout(=x ch),Cin]
        on = -(-N // s)
        y = torch.randn(1, 3, on, on, dtype=torch.float64)
        pad = max((on - 1) * s + k - N, 0)
        xp = torch.nn.functional.pad(x, (pad // 2, pad - pad // 2, pad // 2, pad - pad // 2))
        conv = torch.nn.functional.conv2d(xp, wt.permute(3, 2, 0, 1), stride=s)      # [1,3,on,on]
        lhs = (conv * y).sum()
        rhs = (x * O.deconv_same(y, wt, None, "linear", (N, N), s)).sum()
        assert abs(float(lhs - rhs)) < 1e-9 * max(1.0, abs(float(lhs)))


def test_resize_identities():
    torch.manual_seed(1)
    x = torch.randn(1, 2, 4, 6, dtype=torch.float64)
    for m in ("nearest", "bilinear", "bicubic"):
        np.testing.assert_allclose(O.resize(x, (4, 6), m).numpy(), x.numpy(), atol=1e-6)   # same size = identity
        c = torch.full((1, 1, 3, 5), 2.5, dtype=torch.float64)
        np.testing.assert_allclose(O.resize(c, (37, 41), m).numpy(), 2.5, atol=1e-6)        # weights sum to 1
    # nearest 2 -> 4 with half-pixel centres: [a,a,b,b]
    v = torch.tensor([1.0, 2.0], dtype=torch.float64).view(1, 1, 1, 2)
    np.testing.assert_allclose(O.resize(v, (1, 4), "nearest")[0, 0, 0].numpy(), [1, 1, 2, 2])
    # bilinear 2 -> 4: src = (o+0.5)/2-0.5 = -0.25,0.25,0.75,1.25 -> [1,1.25,1.75,2]
    np.testing.assert_allclose(O.resize(v, (1, 4), "bilinear")[0, 0, 0].numpy(), [1, 1.25, 1.75, 2], atol=1e-7)


def test_spp_shapes_and_values():
    x = torch.arange(2 * 3 * 6 * 10, dtype=torch.float64).view(2, 3, 6, 10)
    out = O.spatial_pyramid_pool(x, [[2, 2], 3, 5], "max")
    assert out.shape == (2, 4 + 9 + 25)
    assert float(out[0, 3]) == float(x[0].max())             # last bin of the 2x2 level holds the global max
    x1 = torch.arange(2 * 3 * 50, dtype=torch.float64).view(2, 3, 50)
    o1 = O.spatial_pyramid_pool(x1, [2, 3, 4, 5, 8, 11, 15, 30, 45], "avg")
    assert o1.shape == (2, 123)
    np.testing.assert_allclose(float(o1[0, 0]), float(x1[0, :, :25].mean()))
    # empty bins (SURVEY 7 "degenerate shapes"): 3x3 map with a 5-bin level -> -inf like tf.reduce_max
    e = O.spatial_pyramid_pool(torch.ones(1, 1, 3, 3, dtype=torch.float64), [5], "max")
    assert torch.isinf(e).any()


def test_dst_solve_satisfies_reference_system():
    """DST-I solve vs a sparse direct solve of the reference's system (cholesky.py:45-119,
    multigrid.py:122-148) and the 3-point residual identity."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from poisson_cnn_b200.synthetic import make_problem
    p = {k: v.double() for k, v in make_problem(2, 21, 17, seed=5).items()}
    sol = O.dst_poisson_solve(p["rhs"], p["left"], p["top"], p["right"], p["bottom"], p["dx"])
    nx, ny = 21, 17
    mx, my = nx - 2, ny - 2
    T = lambda m: sp.diags([-1, 2, -1], [-1, 0, 1], shape=(m, m))
    A = sp.kron(T(mx), sp.identity(my)) + sp.kron(sp.identity(mx), T(my))          # pyamg.gallery.poisson((mx,my))
    for b in range(2):
        Fm = -(float(p["dx"][b]) ** 2) * p["rhs"][b, 0].numpy().copy()
        Fm[1:-1, 1] += p["bottom"][b, 0, 1:-1].numpy(); Fm[1:-1, -2] += p["top"][b, 0, 1:-1].numpy()
        Fm[1, 1:-1] += p["left"][b, 0, 1:-1].numpy(); Fm[-2, 1:-1] += p["right"][b, 0, 1:-1].numpy()
        u = spla.spsolve(A.tocsc(), Fm[1:-1, 1:-1].reshape(-1)).reshape(mx, my)
        np.testing.assert_allclose(sol[b, 0, 1:-1, 1:-1].numpy(), u, rtol=1e-9, atol=1e-11)
    gs = torch.cat([p["dx"], p["dx"]], 1)
    res = O.laplacian_residual(p["rhs"], sol, gs, 3)
    assert float(res) < 1e-18 * float((p["rhs"] ** 2).mean()) + 1e-16
    # ring equals the BCs, left/right written last (corners)
    np.testing.assert_array_equal(sol[:, 0, 0, :].numpy(), p["left"][:, 0].numpy())
    np.testing.assert_array_equal(sol[:, 0, -1, :].numpy(), p["right"][:, 0].numpy())


def test_model_invariants():
    hp, db = pcnn_configs(small_scaling=True)
    w = all_weights(hp, db)
    from poisson_cnn_b200.synthetic import make_problem
    p = {k: v.double() for k, v in make_problem(1, 64, 64, seed=3, magnitudes=False).items()}
    out = O.hpnn_forward(hp, w, p["rhs"], p["dx"], "hpnn/")
    assert float(out[:, :, 0].abs().max()) == 0 and float(out[:, :, :, -1].abs().max()) == 0   # Dirichlet ring exactly 0
    hpn = dict(hp); hpn["bc_type"] = "neumann"
    outn = O.hpnn_forward(hpn, w, p["rhs"], p["dx"], "hpnn/")
    np.testing.assert_array_equal(outn[:, :, 0, 1:-1].numpy(), outn[:, :, 1, 1:-1].numpy())      # SYMMETRIC ring
    d = O.dbcnn_forward(db, w, p["left"], p["dx"], 50, "dbcnn/")
    np.testing.assert_array_equal(d[:, :, 0, :].numpy(), p["left"].numpy())                      # row 0 == BC
    assert abs(float(d[:, :, 1:].abs().max()) - 1.0) < 1e-12 or float(d[:, :, 0].abs().max()) >= 1.0
    with pytest.raises(ValueError):
        O.hpnn_forward(dict(hp, bc_type="robin"), w, p["rhs"], p["dx"], "hpnn/")


@pytest.mark.parametrize("name", ["hpnn_64x64", "dbcnn_56x48", "pcnn_112x120"])
def test_oracle_reproduces_golden(name):
    """The committed golden vectors are what the oracle computes today (float32 mode within 2e-5 of the
    float64 golden; float64 mode to 1e-6 after the fp32 storage rounding)."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    t = lambda k, dt: torch.from_numpy(g[k]).to(dt)
    for dt, tol in ((torch.float64, 1e-6), (torch.float32, 2e-5)):
        if name.startswith("hpnn"):
            hp, db = pcnn_configs(small_scaling=True)
            out = O.hpnn_forward(hp, all_weights(hp, db), t("rhs", dt), t("dx", dt), "hpnn/")
        elif name.startswith("dbcnn"):
            hp, db = pcnn_configs(small_scaling=True)
            out = O.dbcnn_forward(db, all_weights(hp, db), t("bc", dt), t("dx", dt), 56, "dbcnn/")
        else:
            if dt == torch.float32:
                continue   # keep the CPU suite short; the fp64 run pins it
            hp, db = pcnn_configs()
            out = O.pcnn_forward(hp, db, all_weights(hp, db), *(t(k, dt) for k in ("rhs", "left", "top", "right", "bottom", "dx")))
        assert rel_l2(out, g["out"]) < tol


def test_legacy_bicubic_image_resize_cross_checked_against_torch():
    """dataset/utils/image_resize.py:20 (tf.compat.v1 resize_images BICUBIC, align_corners=True).  TF itself cannot
    run here (parity unpinned); the restatement is cross-checked against torch's independent implementation of the
    same filter (cubic convolution a = -0.75, align_corners, clamped taps), which differs only by TF's 1024-step
    coefficient table (|dw| <= ~1e-3), and on the exact properties: control points are reproduced, constants stay
    constant."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(5)
    x = 2 * torch.rand(3, 1, 6, 9, generator=g, dtype=torch.float64) - 1
    got = O.image_resize(x, (41, 73))
    ref = F.interpolate(x, size=(41, 73), mode="bicubic", align_corners=True)
    assert float((got - ref).abs().max()) < 3e-3
    # align_corners: output point i*(out-1)/(in-1) sits on control point i
    same = O.image_resize(x, (11, 17))       # (6-1)*2+1, (9-1)*2+1
    np.testing.assert_allclose(same[:, :, ::2, ::2].numpy(), x.numpy(), atol=1e-12)
    const = O.image_resize(torch.full((1, 1, 5, 5), 0.7, dtype=torch.float64), (20, 33))
    np.testing.assert_allclose(const.numpy(), 0.7, atol=1e-6)      # table rows sum to 1 within float32 rounding


def test_legacy_bicubic_host_tables_match_oracle():
    from poisson_cnn_b200 import ops
    from poisson_cnn_b200.config import RESIZE_BICUBIC_LEGACY_AC
    for n_in, n_out in [(5, 64), (8, 256), (3, 7), (1, 1), (12, 300), (20, 2048)]:
        idx, w = ops.resize_axis_table(n_in, n_out, RESIZE_BICUBIC_LEGACY_AC)
        oi, ow = O.legacy_bicubic_axis_weights(n_in, n_out)
        np.testing.assert_array_equal(idx, oi)
        np.testing.assert_array_equal(w.astype(np.float64), ow)


def test_separable_input_convolution_identity():
    """The algebra behind pcnn_conv2d_tc_rowweights (the DBCNN's first 2-D convolution, Dirichlet_BC_NN_Legacy.py:137-160):
    for in[m,x,y] = h[m,y] * S[m,x] and zero padding, conv(in)[co,x,y] = sum_{b,m} A_x[b,m,co] * h[m, y+b-p] with
    A_x[b,m,co] = sum_a W[a,b,m,co] * S[m, x+a-p]."""
    g = torch.Generator().manual_seed(3)
    Cin, Cout, H, W, k = 6, 5, 9, 11, 5
    p = k // 2
    h = torch.randn(1, Cin, W, generator=g, dtype=torch.float64)
    S = torch.randn(Cin, H, generator=g, dtype=torch.float64)
    kern = torch.randn(k, k, Cin, Cout, generator=g, dtype=torch.float64)
    full = O.conv_nd(h[:, :, None, :] * S[None, :, :, None], kern, None, "linear", "CONSTANT", 0.0)
    Sp = torch.nn.functional.pad(S, (p, p))
    A = torch.einsum("mxa,abmc->xbmc", Sp.unfold(1, k, 1), kern)                       # [H, k, Cin, Cout]
    hp = torch.nn.functional.pad(h[0], (p, p)).unfold(1, k, 1)                          # [Cin, W, k]: h[m, y+b-p]
    folded = torch.einsum("xbmc,myb->cxy", A, hp)
    np.testing.assert_allclose(folded.numpy(), full[0].numpy(), atol=1e-12)


def test_tf2_resize_cross_checked_against_pillow():
    """layers/Upsample.py:56-59 -> tf.image.resize(method, antialias=False) with half-pixel centres.  TF cannot run here
    (parity unpinned); for UP-sampling, Pillow's resize implements the same filters independently: bilinear = triangle
    filter, bicubic = Keys a = -0.5, taps outside the image dropped and the rest renormalised.  The oracle must agree with
    it up to TF's 1024-step coefficient table (bicubic) / float32 rounding (bilinear)."""
    Image = pytest.importorskip("PIL.Image")
    g = torch.Generator().manual_seed(8)
    for (ih, iw), (oh, ow) in (((2, 2), (64, 80)), ((4, 5), (64, 80)), ((8, 8), (256, 256)), ((3, 7), (29, 100))):
        x = torch.randn(ih, iw, generator=g)
        img = Image.fromarray(x.numpy().astype(np.float32), mode="F")
        for method, pil, tol in (("bilinear", Image.BILINEAR, 2e-6), ("bicubic", Image.BICUBIC, 2e-3)):
            ref = np.asarray(img.resize((ow, oh), resample=pil), dtype=np.float64)
            got = O.resize(x.double()[None, None], (oh, ow), method)[0, 0].numpy()
            assert np.abs(got - ref).max() < tol * max(1.0, np.abs(ref).max()), (method, ih, iw, np.abs(got - ref).max())


def test_pressure_projection_oracle_is_consistent():
    """f4 (Navier_Stokes_2D/solvers.py:153-334): the replicate-padded stencil of the CG restatement IS the reference's
    kron-built Neumann matrix, and CG on the mean-projected system converges to the direct solve of the augmented
    (zero-integral Lagrange) system."""
    g = torch.Generator().manual_seed(5)
    m, n, dh = 9, 12, 0.03
    A = O.pressure_poisson_matrix(m, n, dh)
    v = torch.randn(1, 1, m, n, generator=g, dtype=torch.float64)
    got = (A[:-1, :-1] @ v.reshape(-1).numpy()).reshape(m, n)
    vp = torch.nn.functional.pad(v, (1, 1, 1, 1), mode="replicate")[0, 0]
    ref = (4 * v[0, 0] - vp[:-2, 1:-1] - vp[2:, 1:-1] - vp[1:-1, :-2] - vp[1:-1, 2:]) / dh ** 2
    np.testing.assert_allclose(got, ref.numpy(), rtol=1e-12, atol=1e-9)
    rhs = torch.randn(2, 1, m, n, generator=g, dtype=torch.float64)
    dx = torch.tensor([[0.03], [0.011]], dtype=torch.float64)
    direct = O.pressure_poisson_reference(rhs, dx)
    cg, hist = O.neumann_cg(rhs, dx, torch.zeros_like(rhs), 150)
    assert float((cg - direct).norm() / direct.norm()) < 1e-6
    assert float(hist[-1].max()) < 1e-5 and abs(float(direct.mean())) < 1e-10
