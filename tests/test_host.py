"""CPU tests of the host side: the C ABI loads and exports every symbol include/pcnn.h declares, config
plumbing mirrors the reference helpers, weight specs, resize/SPP tables against the oracle, and
the world_size-2 (gloo) sharding path."""
import os
import re
import socket

import numpy as np
import pytest
import torch

from tests.helpers import ROOT, pcnn_configs
from oracle import poisson_oracle as O


def test_cabi_exports_every_declared_symbol():
    import ctypes
    hdr = open(os.path.join(ROOT, "include", "pcnn.h")).read()
    declared = set(re.findall(r"\b(pcnn_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(os.path.join(ROOT, "poisson_cnn_b200", "libpcnn.so"))
    for name in declared:
        assert hasattr(lib, name), name
    from poisson_cnn_b200 import _lib
    assert set(_lib.SIGNATURES) == declared
    assert _lib.lib.pcnn_version() == 200


def test_cabi_rejects_bad_arguments_without_gpu():
    from poisson_cnn_b200 import _lib
    # argument validation happens before any CUDA call, so it works on a CPU-only box
    st = _lib.lib.pcnn_conv2d_f32(None, None, None, None, None, None, None, None, 1, 1, 1, 1, 1, 1, 1, 0, 0.0, 0, 0, 0, 0, None)
    assert st == -1
    assert b"null pointer" in _lib.lib.pcnn_last_error()
    with pytest.raises(ValueError):
        _lib.check(st, "conv2d")
    # workspace queries are pure host arithmetic: the CG solver keeps r, two p buffers and q plus 12 double accumulators per sample
    assert _lib.lib.pcnn_neumann_cg_workspace_bytes(3, 40, 36) == 4 * 3 * 40 * 36 * 4 + 12 * 3 * 8
    assert _lib.lib.pcnn_neumann_cg_workspace_bytes(0, 40, 36) == 0 and _lib.lib.pcnn_neumann_cg_workspace_bytes(1, 1, 36) == 0
    assert _lib.lib.pcnn_neumann_cg_solve(None, None, None, None, 1, 8, 8, 1, 1e-6, None, None, None) == -1
    assert _lib.lib.pcnn_dst_fft_workspace_bytes(2, 10, 12) == (2 + 1) * 8 * 10 * 8      # (B + 1) * (nx - 2) * (ny - 2) doubles
    assert _lib.lib.pcnn_dst_fft_passes() == 3


def test_product_has_no_cpu_fallback():
    from poisson_cnn_b200 import ops
    with pytest.raises(ValueError, match="CUDA tensor"):
        ops.conv2d(torch.zeros(1, 1, 4, 4), torch.zeros(3, 3, 1, 1))
    # the product package must not import the oracle
    import subprocess, sys
    code = "import sys; import poisson_cnn_b200.models, poisson_cnn_b200.losses, poisson_cnn_b200.solvers; print(any(m.startswith('oracle') for m in sys.modules))"
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert out.stdout.strip() == "False", out.stderr


def test_config_helpers():
    from poisson_cnn_b200 import config as C
    cfg = {"key1": 3, "key2": [0, 1, 2, 3, 4], "key3": [6, 7, 8, 9, 10]}
    assert C.get_init_arguments_from_config(cfg, 2, ["key2", "key3"], ["key2p", "key3p"]) == {"key1": 3, "key2p": 2, "key3p": 8}
    conv = C.convert_tf_object_names({"a": "tf.nn.leaky_relu", "b": ["tf.nn.tanh", "tanh", 3], "c": {"d": "linear"}})
    assert C.activation_enum(conv["a"]) == C.ACT_LEAKY_RELU and C.activation_enum(conv["b"][0]) == C.ACT_TANH
    assert conv["b"][1] == "tanh" and conv["c"]["d"] == "linear"
    with pytest.raises(ValueError):
        C.convert_tf_object_names("tf.nn.relu")
    with pytest.raises(ValueError):
        C.activation_enum("tf.nn.softmax")
    assert C.process_normalizations(None) == {"rhs_max_magnitude": False}
    assert C.process_normalizations({"rhs_max_magnitude": True})["rhs_max_magnitude"] == 1.0
    assert C.process_output_scaling_modes({"soln_max_magnitude": True})["max_domain_size_squared"] is False
    # float64 truncation of int((N/ds)*us): exact for the shipped factors, off by one for ds=50 at e.g. N=29
    for n in range(1, 600):
        for ds in (2, 3, 4, 8, 16, 32, 64, 128):
            assert C.bottleneck_output_size(n, ds, ds) == n
    assert C.bottleneck_output_size(29, 50, 50) == 28


def test_model_construction_errors_and_param_counts():
    from poisson_cnn_b200 import models, convert_tf_object_names
    hp, db = pcnn_configs()
    hp, db = convert_tf_object_names(hp), convert_tf_object_names(db)
    h = models.Homogeneous_Poisson_NN_Legacy(**hp)
    d = models.Dirichlet_BC_NN_Legacy_2(**db)
    p = models.Poisson_CNN_Legacy(h, d)
    assert (h.count_params(), d.count_params(), p.count_params()) == (5559108, 483878, 6042986)
    assert p.data_format == "channels_first" and p.hpnn is h and p.dbcnn is d
    assert [b.downsampling_factor for b in h.bottleneck_deconv_blocks] == [16, 8, 4, 3, 2]
    assert [b.downsampling_factor for b in h.bottleneck_multilinear_blocks] == [128, 64, 32]
    for missing in ("pre_bottleneck_convolutions_config", "bottleneck_deconv_config", "final_convolutions_config"):
        with pytest.raises(ValueError):
            models.Homogeneous_Poisson_NN_Legacy(**{k: v for k, v in hp.items() if k != missing})
    with pytest.raises(ValueError):
        models.Homogeneous_Poisson_NN_Legacy(**dict(hp, bc_type="robin"))
    for missing in ("boundary_conv_config", "spp_config", "domain_info_mlp_config", "final_convolutions_config"):
        with pytest.raises(ValueError):
            models.Dirichlet_BC_NN_Legacy_2(**{k: v for k, v in db.items() if k != missing})
    with pytest.raises(RuntimeError):
        h.w("pre_bottleneck/0/kernel")       # no weights loaded yet


def test_weights_npz_roundtrip(tmp_path):
    from poisson_cnn_b200 import weights as W
    hp, db = pcnn_configs()
    sm = W.dbcnn_weight_specs(db, "dbcnn/")
    w = W.synthetic_weights(sm, seed=3)
    assert set(w) == set(sm[0]) and all(w[k].shape == tuple(sm[0][k]) for k in w)
    w2 = W.synthetic_weights(sm, seed=3)
    assert all(np.array_equal(w[k], w2[k]) for k in w)              # deterministic
    path = str(tmp_path / "w.npz")
    W.save_npz(path, w)
    back = W.load_npz(path)
    assert set(back) == set(w) and all(np.array_equal(back[k], w[k]) for k in w)


def test_host_tables_match_oracle():
    from poisson_cnn_b200 import ops
    from poisson_cnn_b200.config import resize_enum
    for n_in, n_out in ((2, 256), (4, 200), (8, 300), (5, 5), (7, 3)):
        for m in ("nearest", "bilinear", "bicubic"):
            i0, w0 = O.resize_axis_weights(n_in, n_out, m)
            i1, w1 = ops.resize_axis_table(n_in, n_out, resize_enum(m))
            np.testing.assert_array_equal(i0, i1)
            np.testing.assert_allclose(w0, w1, atol=1e-7)
    b = ops.spp_boxes([[2, 2], 3, 5], 10, 13, 2)
    assert b.shape == (38, 4) and tuple(b[0]) == (0, 5, 0, 7) and tuple(b[3]) == (5, 10, 7, 13)
    b1 = ops.spp_boxes([2, 3, 4, 5, 8, 11, 15, 30, 45], 1, 229, 1)
    assert b1.shape == (123, 4) and tuple(b1[0]) == (0, 1, 0, 115)
    np.testing.assert_allclose(ops.tf_linspace01(7), O.tf_linspace01(7, torch.float32).numpy(), atol=0)


def test_shard_bounds_cover_batch():
    from poisson_cnn_b200.sharding import shard_bounds, bucket_by_shape
    for n in (1, 7, 16, 128, 256):
        for ws in (1, 2, 4, 8):
            spans = [shard_bounds(n, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert bucket_by_shape([(384, 128), (512, 256), (384, 128)]) == {(384, 128): [0, 2], (512, 256): [1]}


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from poisson_cnn_b200.sharding import init_from_env, shard_problem, ErrorStats
    from poisson_cnn_b200.synthetic import make_problem
    init_from_env("gloo")
    full = make_problem(5, 12, 9, seed=9)
    mine = shard_problem(full, world, rank)
    noise = torch.full_like(mine["rhs"], 1e-3 * (rank + 1))
    st = ErrorStats().update(mine["rhs"] + noise, mine["rhs"]).combine().result()
    q.put((rank, mine["rhs"].shape[0], st))
    dist.destroy_process_group()


def test_world_size_2_gloo_sharding_and_stats():
    import torch.multiprocessing as mp
    from poisson_cnn_b200.synthetic import make_problem
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert [r[1] for r in res] == [3, 2]                       # 5 samples -> 3 + 2
    full = make_problem(5, 12, 9, seed=9)["rhs"].double()
    sq_err = 3 * 108 * 1e-6 + 2 * 108 * 4e-6
    expect = (sq_err / float(full.pow(2).sum())) ** 0.5
    for r in res:                                              # both ranks hold the combined statistics
        assert r[2]["samples"] == 5
        assert abs(r[2]["rel_l2"] - expect) < 1e-3 * expect
        assert abs(r[2]["max_abs_err"] - 2e-3) < 1e-6


def _gloo_spatial_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import types
    import torch.distributed as dist
    from poisson_cnn_b200.sharding import init_from_env
    from poisson_cnn_b200.spatial import SpatialHPNN, band_bounds, _view, HALO
    init_from_env("gloo")
    B, C, H, W = 2, 32, 96, 20
    bnd = band_bounds(H, world, [2, 3, 4])                     # lcm 12: 36 + 24 + 36 rows on 3 ranks
    h = bnd[rank + 1] - bnd[rank]
    planes = (C + 15) // 16 * 2
    g = torch.Generator().manual_seed(5)
    full = torch.randn((B, planes, H + 2 * HALO, W + 2 * HALO, 8), generator=g).to(torch.float16)     # same on every rank
    full_lo = (full.float() * 3).to(torch.float16)

    def band_of(src):
        t = torch.zeros((B, planes, h + 2 * HALO, W + 2 * HALO, 8), dtype=torch.float16)
        t[:, :, HALO:HALO + h] = src[:, :, HALO + bnd[rank]:HALO + bnd[rank + 1]]
        return t
    blk = types.SimpleNamespace(B=B, C=C, H=h, W=W, halo=(0, 0), buf=band_of(full).reshape(-1), lo=band_of(full_lo).reshape(-1))
    sp = SpatialHPNN(None)                                     # the default group: one band per rank
    sp.bounds = bnd
    sp._exchange({rank: blk})
    sp._exchange({rank: blk})                                  # the staging buffers are reused: a second exchange must not corrupt
    ok = True
    for which, src in (("buf", full), ("lo", full_lo)):
        v = _view(blk, which)
        lo_row = HALO if rank == 0 else 0                      # physical edges keep their (zero) halo rows
        hi_row = h + HALO if rank == world - 1 else h + 2 * HALO
        want = src[:, :, bnd[rank] + lo_row:bnd[rank] + hi_row]
        ok = ok and torch.equal(v[:, :, lo_row:hi_row], want)
        if rank == 0:
            ok = ok and float(v[:, :, :HALO].abs().max()) == 0.0
    # ragged gather of a pooled level (div = 5 does not divide the 24-row band: ceil)
    div = 5
    rows = [-(-(bnd[i + 1] - bnd[i]) // div) for i in range(world)]
    part = torch.full((B, 3, rows[rank], 4), float(rank + 1))
    got = sp._gather_rows({rank: part}, div=div)
    want = torch.cat([torch.full((B, 3, rows[i], 4), float(i + 1)) for i in range(world)], 2)
    q.put((rank, bool(ok), bool(torch.equal(got, want)), blk.halo, bnd))
    dist.destroy_process_group()


def test_world_size_3_gloo_halo_exchange_and_row_gather():
    """The N > 1 host logic of the spatial decomposition (poisson_cnn_b200/spatial.py) on CPU tensors: after the exchange
    the halo rows of a band hold the neighbour's edge rows of BOTH operand buffers, physical edges are untouched, and the
    gather of uneven pooled bands reassembles the map."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_spatial_worker, args=(r, 3, port, q)) for r in range(3)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=180) for _ in range(3))
    [p.join(60) for p in procs]
    assert [r[0] for r in res] == [0, 1, 2]
    for r in res:
        assert r[1], "halo rows of rank %d differ from the neighbour's edge rows" % r[0]
        assert r[2], "row gather on rank %d" % r[0]
        assert r[3][1] == 7 and r[4] == [0, 36, 60, 96]


# ------------------------------------------------------------------ TF checkpoint (tensor bundle) reader
def test_crc32c_and_snappy_known_answers():
    from poisson_cnn_b200 import tf_checkpoint as T
    assert T.crc32c(b"123456789") == 0xE3069283                       # the CRC-32C check value
    assert T._mask_crc(T.crc32c(b"")) == 0xa282ead8
    # snappy: literal "abcd" + copy(offset 4, length 8) -> "abcdabcdabcd"; preamble = uncompressed length
    stream = bytes([12, (4 - 1) << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4])
    assert T._snappy_decompress(stream) == b"abcdabcdabcd"


def test_table_roundtrip_prefix_compression_and_blocks(tmp_path):
    from poisson_cnn_b200 import tf_checkpoint as T
    items = [(("layer/%03d/kernel" % i).encode(), bytes([i % 251]) * (i % 40)) for i in range(300)] + [(b"", b"header")]
    path = str(tmp_path / "t.index")
    T.write_table(path, items, block_size=512)                          # many data blocks, restarts every 16 keys
    assert T.read_table(path, verify=True) == sorted(items)
    raw = bytearray(open(path, "rb").read())
    raw[10] ^= 0xFF                                                     # corrupt a data block: the checksum must notice
    open(path, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        T.read_table(path, verify=True)
    with pytest.raises(ValueError):
        T.read_table(__file__)                                          # not a table at all


def test_keras_key_map_is_a_bijection_onto_the_weight_specs():
    from poisson_cnn_b200 import load_experiment, weights as W, tf_checkpoint as T
    cfg = load_experiment("pcnn_end_to_end")
    hp, db = cfg["hpnn_model"], cfg["dbcnn_model"]
    specs = {**W.hpnn_weight_specs(hp, "hpnn/")[0], **W.dbcnn_weight_specs(db, "dbcnn/")[0]}
    m = T.pcnn_key_map(hp, db)
    assert sorted(m.values()) == sorted(specs)                          # every variable has exactly one reference key
    assert len(set(m)) == len(m)
    # spot checks against the reference's attribute structure
    assert m["hpnn/pre_bottleneck_convolutions/0/kernel"] == "hpnn/pre_bottleneck/0/kernel"
    ds = hp["bottleneck_deconv_config"]["downsampling_factors"]
    j = sorted(range(len(ds)), key=lambda i: ds[i], reverse=True).index(0)   # the reference sorts blocks by ds, descending
    assert m["hpnn/bottleneck_deconv_blocks/%d/upsample_layer/kernel" % j] == "hpnn/bottleneck_deconv/0/deconv/kernel"
    assert m["hpnn/bottleneck_deconv_blocks/%d/conv_layers/1/conv_layers/2/bias" % j] == "hpnn/bottleneck_deconv/0/resnet1/conv2/bias"
    assert m["hpnn/final_convolutions/1/conv_layers/0/kernel"] == "hpnn/final/0/resnet/conv0/kernel"
    assert m["dbcnn/domain_info_dense_layers/0/kernel"] == "dbcnn/mlp/0/kernel"


def test_tf_checkpoint_roundtrip_through_load_checkpoint_weights(tmp_path):
    from poisson_cnn_b200 import load_experiment, weights as W, tf_checkpoint as T
    cfg = load_experiment("pcnn_end_to_end")
    hp, db = cfg["hpnn_model"], cfg["dbcnn_model"]
    hs, ds = W.hpnn_weight_specs(hp, "hpnn/"), W.dbcnn_weight_specs(db, "dbcnn/")
    specs = {k: v for k, v in {**hs[0], **ds[0]}.items() if int(np.prod(v)) <= 4096}     # small tensors: the pure-Python CRC is slow
    meta = {k: {**hs[1], **ds[1]}[k] for k in specs}
    w = W.synthetic_weights((specs, meta), seed=3)
    key_map = {k: v for k, v in T.pcnn_key_map(hp, db).items() if v in specs}
    prefix = str(tmp_path / "ckpt" / "weights-01")
    T.save_checkpoint_weights(prefix, w, key_map)
    assert os.path.isfile(prefix + ".index") and os.path.isfile(prefix + ".data-00000-of-00001")
    back = T.load_checkpoint_weights(prefix, key_map, verify=True)
    assert sorted(back) == sorted(w)
    for k in w:
        np.testing.assert_array_equal(back[k], w[k])
    raw = T.read_tensor_bundle(prefix)
    assert all(k.endswith(T.SUFFIX) for k in raw)
    with pytest.raises(ValueError):                                     # a variable the model needs but the file lacks
        T.load_checkpoint_weights(prefix, {**key_map, "hpnn/not_there": "hpnn/not_there"})


def test_upsample_merge_tc_shared_memory_plan():
    """The tensor-core upsample-merge kernel stages every branch operand in shared memory; the host asks the library whether
    a configuration fits (grids beyond ~400 pixels a side fall back to the FP32-FMA kernel)."""
    from poisson_cnn_b200 import ops
    strides = [2, 3, 4, 8, 16]
    assert ops.upsample_merge_tc_fits(strides, [(8, 8), (4, 4), (2, 2)])            # 256 x 256
    assert ops.upsample_merge_tc_fits(strides, [(16, 8), (8, 4), (4, 2)])           # 512 x 256
    assert not ops.upsample_merge_tc_fits(strides, [(16, 16), (8, 8), (4, 4)])      # 512 x 512
    assert not ops.upsample_merge_tc_fits([], [(2, 2)])                              # needs a transpose-conv branch
    assert not ops.upsample_merge_tc_fits([64], [])                                  # stride beyond 32


def test_rowweights_image_layout():
    """include/pcnn.h, pcnn_conv2d_tc_rowweights: slot order, K halves, channel padding and pre-scale of the per-row weights
    (an explicit loop over the documented layout against the vectorised packing)."""
    from poisson_cnn_b200 import ops
    from poisson_cnn_b200._lib import lib
    g = torch.Generator().manual_seed(4)
    k, Cin, Cout, H = 3, 19, 7, 11
    kern = torch.randn(k, k, Cin, Cout, generator=g)
    basis = torch.randn(Cin, H, generator=g)
    cp = lib.pcnn_conv_tc_channel_slots(Cout, k)
    rt = 5 if cp == 24 else 128 // cp
    T = lib.pcnn_conv_tc_rowweight_slots(Cout, k, H)
    assert cp == 8 and rt == 16 and T == -(-H // rt) * rt + 2
    img, scale = ops.rowweights_image(kern, basis, cp, rt, T)
    assert img.shape == (2, k, 2, T, cp, 8) and img.dtype == torch.float16
    assert scale == 2.0 ** np.floor(np.log2(scale)) and scale > 0
    p = k // 2
    for t in (0, 3, rt - 1, rt, T - 1):
        x = (t // rt) * rt + (rt - 1 - t % rt)
        for b in range(k):
            for ci in (0, 7, 8, 18, 19, 31):
                for co in (0, 6, 7):
                    want = 0.0
                    if x < H and ci < Cin and co < Cout:
                        want = scale * sum(float(kern[a, b, ci, co]) * float(basis[ci, x + a - p]) for a in range(k) if 0 <= x + a - p < H)
                    got = float(img[ci // 16, b, (ci % 16) // 8, t, co, ci % 8])
                    assert abs(got - want) <= 1e-3 * max(1.0, abs(want)), (t, x, b, ci, co, got, want)


# ------------------------------------------------------------------ model-level C ABI (csrc/engine.cu), host-only parts
def _create(cfg):
    import ctypes, json
    from poisson_cnn_b200 import _lib
    h = ctypes.c_void_p()
    st = _lib.lib.pcnn_create(json.dumps(cfg).encode(), 0, ctypes.byref(h))
    return st, h, _lib.lib.pcnn_last_error().decode()


def test_engine_create_parses_reference_configs_without_gpu():
    from poisson_cnn_b200 import _lib, load_experiment
    cfg = load_experiment("pcnn_end_to_end")
    st, h, msg = _create({"hpnn_model": cfg["hpnn_model"], "dbcnn_model": cfg["dbcnn_model"]})
    assert st == 0, msg
    # a forward / workspace query before finalize_weights is an argument error, not a crash
    import ctypes
    n = ctypes.c_size_t(0)
    assert _lib.lib.pcnn_workspace_bytes(h, 4, 64, 64, ctypes.byref(n)) == -1
    assert b"finalize" in _lib.lib.pcnn_last_error()
    assert _lib.lib.pcnn_destroy(h) == 0
    for name in ("hpnn_neumann", "hpnn_smalldomain"):
        st, h, msg = _create({"model": load_experiment(name)["model"]})
        assert st == 0, msg
        _lib.lib.pcnn_destroy(h)


def test_engine_create_reports_the_reference_config_errors():
    import copy
    from poisson_cnn_b200 import load_experiment
    cfg = load_experiment("pcnn_end_to_end")
    hp, db = cfg["hpnn_model"], cfg["dbcnn_model"]
    cases = [
        ({k: v for k, v in hp.items() if k != "pre_bottleneck_convolutions_config"}, "hpnn_model", "Provide a config for pre bottleneck convolutions"),
        ({k: v for k, v in hp.items() if k != "bottleneck_multilinear_config"}, "hpnn_model", "Provide a config for bottleneck blocks"),
        ({k: v for k, v in hp.items() if k != "final_convolutions_config"}, "hpnn_model", "Provide a config for final convolutions"),
        (dict(hp, bc_type="robin"), "hpnn_model", "bc_type can only be neumann or dirichlet."),
        ({k: v for k, v in db.items() if k != "spp_config"}, "dbcnn_model", "Provide a config for the Spatial Pyramid Pooling."),
        ({k: v for k, v in db.items() if k != "domain_info_mlp_config"}, "dbcnn_model", "Provide a config for the domain info MLP."),
    ]
    for sub, key, text in cases:
        st, h, msg = _create({key: sub})
        assert st == -1 and text in msg, (key, msg)
    bad = copy.deepcopy(hp)
    bad["pre_bottleneck_convolutions_config"]["activation"] = "tf.nn.softmax"
    st, h, msg = _create({"hpnn_model": bad})
    assert st == -1 and "unsupported activation" in msg
    from poisson_cnn_b200 import _lib
    import ctypes
    assert _lib.lib.pcnn_create(b"{not json", 0, ctypes.byref(ctypes.c_void_p())) == -1
    assert _lib.lib.pcnn_create(b"{}", 0, ctypes.byref(ctypes.c_void_p())) == -1


def test_engine_host_tables_match_the_python_tables():
    """The C++ engine builds its tables on the host (csrc/engine.cu); they must agree with the numpy restatements of the
    Python host (ops.resize_axis_table: bit-exact; cos / sinh bases: within one float32 ulp of numpy's own libm)."""
    import math
    from poisson_cnn_b200 import ops
    for n in (1, 2, 7, 256, 300):
        ref = np.cos(np.float32(math.pi) * ops.tf_linspace01(n)).astype(np.float32)
        got = ops.host_table("pos", n, n)
        assert np.abs(got - ref).max() <= 6e-8 * 1.01
    M, xres = 27, 112
    xbar = ops.tf_linspace01(xres)
    arg = np.arange(1, M + 1, dtype=np.float32)[:, None] * (np.float32(math.pi) * (xbar - np.float32(1)))[None, :]
    s = np.sinh(arg.astype(np.float32)).astype(np.float32)
    s = s * (np.float32(1.0) / np.abs(s).max(1, keepdims=True))
    got = ops.host_table("sinh", M * xres, M, xres).reshape(M, xres)
    np.testing.assert_allclose(got, s, rtol=3e-7, atol=1e-37)
    for (a, b) in ((2, 256), (4, 300), (8, 109), (7, 64), (1, 5)):
        for m in (0, 1, 2):
            idx, w = ops.resize_axis_table(a, b, m)
            taps = idx.shape[1]
            np.testing.assert_array_equal(ops.host_table("resize_idx", b * taps, a, b, m, np.int32).reshape(b, taps), idx)
            np.testing.assert_array_equal(ops.host_table("resize_w", b * taps, a, b, m).reshape(b, taps), w)


def test_engine_host_rowweights_match_the_torch_packer():
    """pcnn_host_rowweights (what the engine uploads) against ops.rowweights_image (torch einsum): same layout, same
    power-of-two scale; values equal up to one fp16 ulp on a handful of elements (float summation order)."""
    import ctypes, math
    from poisson_cnn_b200 import ops, _lib
    g = torch.Generator().manual_seed(0)
    k, Cin, Cout, H = 7, 29, 23, 112
    kern = torch.randn(k, k, Cin, Cout, generator=g) / 30
    M = Cin - 2
    S = torch.from_numpy(ops.host_table("sinh", M * H, M, H).reshape(M, H))
    pos = torch.from_numpy(ops.host_table("pos", H, H))
    basis = torch.cat([S, pos[None], torch.ones(1, H)], 0)
    cp = _lib.lib.pcnn_conv_tc_channel_slots(Cout, k)
    T = _lib.lib.pcnn_conv_tc_rowweight_slots(Cout, k, H)
    rt = 5 if cp == 24 else 128 // cp
    ref, scale = ops.rowweights_image(kern, basis, cp, rt, T)
    img = np.zeros(ref.numel(), dtype=np.float16)
    sc = ctypes.c_float(0)
    kk = np.ascontiguousarray(kern.numpy())
    n = _lib.lib.pcnn_host_rowweights(kk.ctypes.data_as(ctypes.c_void_p), k, Cin, Cout, H, img.ctypes.data_as(ctypes.c_void_p), img.nbytes, ctypes.byref(sc))
    assert n == ref.numel() and abs(sc.value - 1.0 / scale) == 0.0
    r = ref.numpy().ravel().astype(np.float32)
    d = np.abs(img.astype(np.float32) - r)
    assert (d > 0).sum() < 1e-3 * d.size
    assert np.all(d <= np.maximum(np.abs(r), 2.0 ** -14) * 2.0 ** -10 * 1.01)


def test_spatial_band_bounds():
    """Band boundaries of the single-grid decomposition: multiples of the lcm of the transpose-conv strides, near-even."""
    from poisson_cnn_b200.spatial import band_bounds
    assert band_bounds(2048, 2, [16, 8, 4, 3, 2]) == [0, 1008, 2048]
    b = band_bounds(2048, 8, [16, 8, 4, 3, 2])
    assert b[0] == 0 and b[-1] == 2048 and all(r % 48 == 0 for r in b[:-1]) and min(b[i + 1] - b[i] for i in range(8)) >= 240
    assert band_bounds(128, 4, [16, 8, 4, 3, 2]) == [0, 32, 64, 96, 128]          # too small for aligned bands: equal split
    assert band_bounds(256, 4, [16, 8, 4, 3, 2]) == [0, 48, 144, 192, 256]
    with pytest.raises(ValueError):
        band_bounds(48, 4, [2])
