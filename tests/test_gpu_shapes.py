"""GPU tests at the BASELINE.json shapes the oracle is too slow for (configs 3-5): variable-aspect and
large grids.  Parity is carried by size-independent properties:
  * the tensor-core path (tc2) against the strict-FP32 CUDA path on the same inputs (the FP32 path is
    pinned to the oracle on the golden vectors in test_gpu_parity.py),
  * homogeneity of the merged model (scaling every input by c scales the output by c),
  * the boundary row of each DBCNN output equals the boundary condition,
  * a DST direct solve has (near-)zero 3-point Laplacian residual and reproduces its boundary data.
"""
import pytest
import torch

from tests.helpers import pcnn_configs, all_weights, rel_l2

pytestmark = pytest.mark.gpu
KEYS = ("rhs", "left", "top", "right", "bottom", "dx")


@pytest.fixture(scope="module")
def model():
    from poisson_cnn_b200 import convert_tf_object_names, models
    hp, db = pcnn_configs()
    m = models.Poisson_CNN_Legacy(models.Homogeneous_Poisson_NN_Legacy(**convert_tf_object_names(hp)),
                                  models.Dirichlet_BC_NN_Legacy_2(**convert_tf_object_names(db)))
    return m.load_weights(all_weights(hp, db))


def _problem(B, nx, ny, seed):
    from poisson_cnn_b200.synthetic import make_problem
    p = make_problem(B, nx, ny, seed=seed)
    return [p[k].cuda() for k in KEYS]


@pytest.mark.parametrize("nx,ny,B", [(384, 128, 2), (512, 256, 2), (200, 300, 3), (256, 256, 2), (512, 512, 1)])
def test_variable_aspect_grids_tc2_tracks_fp32(model, nx, ny, B):
    """BASELINE config 3 shapes (non-square: 2B+2B boundary batches, ragged tiles, two column tiles)."""
    inp = _problem(B, nx, ny, seed=1003)
    ref = model.set_precision("fp32")(inp)
    for mode in ("tc2", "mixed"):
        out = model.set_precision(mode)(inp)
        assert out.shape == (B, 1, nx, ny) and bool(torch.isfinite(out).all())
        assert rel_l2(out, ref) < 2e-3, mode      # tensor-core budget; typically ~3e-4
    # homogeneity: the merged model normalises its inputs per sample and undoes it; with a power-of-two factor
    # every intermediate is bit-identical, so the outputs must scale exactly (a factor like 3 perturbs the
    # normalised inputs by an ulp, which this chaotic seeded network amplifies to ~1e-4 in tensor-core mode)
    out4 = model([t * 4.0 if i < 5 else t for i, t in enumerate(inp)])
    assert rel_l2(out4, 4.0 * out) < 1e-6
    model.set_precision("fp32")


def test_large_grid_1024_with_residual(model):
    """BASELINE config 4 (1024x1024): forward in tc2 vs FP32 path, then the 5-point residual kernel."""
    from poisson_cnn_b200.losses import linear_operator_loss
    inp = _problem(1, 1024, 1024, seed=1004)
    ref = model.set_precision("fp32")(inp)
    out = model.set_precision("tc2")(inp)
    model.set_precision("fp32")
    assert rel_l2(out, ref) < 2e-3
    gs = torch.cat([inp[5], inp[5]], 1)
    loss = linear_operator_loss(3, 2, ndims=2)
    r_tc, r_32 = float(loss(inp[0], out, gs)), float(loss(inp[0], ref, gs))
    assert abs(r_tc - r_32) < 2e-2 * r_32       # the residual of the two paths agrees (it is huge: random weights)


def test_micro_batching_is_transparent(model):
    inp = _problem(5, 120, 112, seed=1005)
    model.set_precision("fp32")
    full = model(inp)
    model.microbatch_samples = 2
    try:
        chunked = model(inp)
    finally:
        model.microbatch_samples = None
    assert torch.equal(full, chunked)


@pytest.mark.parametrize("n", [1024, 2048])
def test_dst_solve_large(n):
    """BASELINE config 5: the DST ground-truth solve at 1024^2 / 2048^2 satisfies its own discrete system."""
    from poisson_cnn_b200.losses import linear_operator_loss
    from poisson_cnn_b200.solvers import dst_poisson_solve
    rhs, left, top, right, bottom, dx = _problem(1, n, n, seed=1005)
    sol = dst_poisson_solve(rhs, {"left": left, "top": top, "right": right, "bottom": bottom}, dx)
    assert torch.equal(sol[:, 0, 0, :], left[:, 0]) and torch.equal(sol[:, 0, -1, :], right[:, 0])
    assert torch.equal(sol[:, 0, 1:-1, 0], bottom[:, 0, 1:-1]) and torch.equal(sol[:, 0, 1:-1, -1], top[:, 0, 1:-1])
    r = float(linear_operator_loss(3, 2, ndims=2)(rhs, sol, torch.cat([dx, dx], 1)))
    # fp32 storage of u limits the residual: |u| * 2^-24 / dx^2 per point
    bound = (float(sol.abs().max()) * 2.0 ** -23 / float(dx.min()) ** 2) ** 2 * 64
    assert r < max(bound, 1e-6 * float((rhs ** 2).mean()))


def test_hpnn_config1_shape_all_modes():
    """BASELINE config 1: Homogeneous_Poisson_NN forward, batch 4, 64x64, zero Dirichlet ring."""
    import numpy as np, os
    from tests.helpers import GOLDEN
    from poisson_cnn_b200 import convert_tf_object_names, models
    hp, db = pcnn_configs(small_scaling=True)
    w = all_weights(hp, db)
    m = models.Homogeneous_Poisson_NN_Legacy(**convert_tf_object_names(hp)).load_weights(w, "hpnn/")
    g = np.load(os.path.join(GOLDEN, "hpnn_64x64.npz"))
    rhs = torch.from_numpy(np.concatenate([g["rhs"], g["rhs"][::-1].copy()])).cuda()
    dx = torch.from_numpy(np.concatenate([g["dx"], g["dx"][::-1].copy()])).cuda()
    gold = torch.from_numpy(np.concatenate([g["out"], g["out"][::-1].copy()]))
    for mode, tol in (("fp32", 1e-5), ("tc2", 2e-3), ("tc3", 2e-3)):
        out = m.set_precision(mode)([rhs, dx])
        assert out.shape == (4, 1, 64, 64)
        assert rel_l2(out, gold) < tol, mode
        assert float(out[:, :, 0].abs().max()) == 0.0 and float(out[:, :, :, -1].abs().max()) == 0.0


def test_load_weights_from_tf_checkpoint_prefix(model, tmp_path):
    """model.load_weights(<TF checkpoint prefix>) -- the reference's weight format (train/utils.py:12-15) -- through
    the pure-Python tensor-bundle reader gives the same network as the in-memory weights."""
    from poisson_cnn_b200 import convert_tf_object_names, models, tf_checkpoint as T
    from tests.helpers import pcnn_configs
    hp, db = pcnn_configs()
    prefix = str(tmp_path / "chkpt" / "cp-0001")
    T.save_checkpoint_weights(prefix, model.get_weights_dict(), model.keras_key_map())
    other = models.Poisson_CNN_Legacy(models.Homogeneous_Poisson_NN_Legacy(**convert_tf_object_names(hp)),
                                      models.Dirichlet_BC_NN_Legacy_2(**convert_tf_object_names(db))).load_weights(prefix)
    inp = _problem(2, 120, 112, seed=1006)
    model.set_precision("fp32"); other.set_precision("fp32")
    a, b = model(inp), other(inp)
    assert bool(torch.isfinite(a).all()) and torch.equal(a, b)


def test_random_grid_shapes_tensor_core_modes_track_fp32(model):
    """Ragged tiles for every accumulator-tile variant (4x32, 5x24, 8x16, 16x8 rows x channel slots), odd sizes, two
    column tiles: the tensor-core modes stay inside the 2e-3 budget against the strict-FP32 CUDA path."""
    import random
    rng = random.Random(7)
    checked = 0
    while checked < 6:
        nx, ny, B = rng.randint(100, 330), rng.randint(100, 330), rng.randint(1, 3)
        inp = _problem(B, nx, ny, seed=checked)
        try:
            ref = model.set_precision("fp32")(inp)
        except ValueError:           # sizes the reference's bottleneck size arithmetic rejects (int((N/ds)*us) != N)
            continue
        if not bool(torch.isfinite(ref).all()):
            continue
        for mode in ("mixed", "tc2"):
            out = model.set_precision(mode)(inp)
            assert bool(torch.isfinite(out).all()) and rel_l2(out, ref) < 2e-3, (nx, ny, B, mode)
        checked += 1


def test_host_buffer_call_matches_device_call(model):
    """model(host tensors) streams slices of the batch through the device on separate copy streams: same bits as the
    device-resident call, pinned or pageable inputs, caller-provided output buffer."""
    from poisson_cnn_b200.synthetic import make_problem
    p = make_problem(5, 112, 120, seed=1010)
    host = [p[k] for k in KEYS]
    model.set_precision("mixed")
    model.microbatch_samples = 2                 # 3 slices: 2 + 2 + 1
    try:
        ref = model([t.cuda() for t in host])
        got = model(host)
        torch.cuda.synchronize()
        assert not got.is_cuda and got.is_pinned() and bool(torch.isfinite(got).all()) and torch.equal(got, ref.cpu())
        out = torch.empty(5, 1, 112, 120).pin_memory()
        got2 = model([t.pin_memory() for t in host], out=out)
        torch.cuda.synchronize()
        assert got2 is out and torch.equal(out, ref.cpu())
        with pytest.raises(ValueError):
            model(host, out=torch.empty(4, 1, 112, 120))
        with pytest.raises(ValueError):
            model([host[0].cuda()] + host[1:])
    finally:
        model.microbatch_samples = None


def test_cuda_graph_replay_matches_eager(model):
    """model.capture(): the replayed CUDA graph (private buffer pool, steady-state halo contents) returns the same
    bits as the eager call, for several different inputs and interleaved with eager calls of other shapes."""
    from poisson_cnn_b200.synthetic import make_problem
    model.set_precision("mixed")
    ex = [t.cuda() for t in (make_problem(2, 112, 120, seed=1)[k] for k in KEYS)]
    g = model.capture(ex)
    for seed in (2, 3, 4):
        inp = [make_problem(2, 112, 120, seed=seed)[k].cuda() for k in KEYS]
        ref = model(inp)
        other = model(_problem(1, 128, 128, seed=9))          # eager traffic through the process-wide pool in between
        got = g(inp).clone()
        assert bool(torch.isfinite(ref).all()) and torch.equal(got, ref)
        assert bool(torch.isfinite(other).all())
    model.set_precision("fp32")
    g32 = model.capture(ex)
    inp = [make_problem(2, 112, 120, seed=5)[k].cuda() for k in KEYS]
    assert torch.equal(g32(inp), model(inp))
    with pytest.raises(ValueError):
        g32([t[:1] for t in inp])
