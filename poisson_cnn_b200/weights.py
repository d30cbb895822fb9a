"""Name-addressed weights for the hot-path models.

The reference keeps its variables inside Keras layers and saves them with
ModelCheckpoint(save_weights_only=True) (poisson_CNN/train/pcnn_end_to_end.py:43).  No trained
weights ship with the reference, so this module provides
  * weight_specs(): every variable of a model (name -> shape, Keras layouts:
      Conv2D [kh,kw,Cin,Cout], Conv1D [k,Cin,Cout], Dense [in,out], BN gamma/beta/mean/var [C],
      deconvupscale kernel [kh,kw,Cout,Cin] (poisson_CNN/layers/deconvupscale.py:58)),
  * synthetic_weights(): a seeded, variance-preserving random initialisation,
  * save_npz()/load_npz(): the .npz interchange format `model.load_weights()` reads.
"""
import math
import zlib

import numpy as np

from .config import get_init_arguments_from_config, activation_enum, ACT_LEAKY_RELU, ACT_TANH


def _conv(specs, meta, name, ksize, cin, cout, act, use_bias=True, ndim=2, resnet_tail=False):
    specs[name + "/kernel"] = tuple([ksize] * ndim + [cin, cout])
    meta[name + "/kernel"] = ("conv", act, resnet_tail)
    if use_bias:
        specs[name + "/bias"] = (cout,)
        meta[name + "/bias"] = ("bias", act, False)


def _bn(specs, meta, name, c):
    for k in ("gamma", "beta", "mean", "var"):
        specs[name + "/" + k] = (c,)
        meta[name + "/" + k] = ("bn_" + k, 0, False)


def _resnet(specs, meta, name, ksize, c, act, use_bn, use_bias=True, ndim=2):
    for i in range(3):
        _conv(specs, meta, "%s/conv%d" % (name, i), ksize, c, c, act, use_bias, ndim, resnet_tail=(i == 2))
    if use_bn:
        _bn(specs, meta, name + "/bn0", c)
        _bn(specs, meta, name + "/bn1", c)


def _dense(specs, meta, name, cin, cout, act):
    specs[name + "/kernel"] = (cin, cout)
    meta[name + "/kernel"] = ("dense", act, False)
    specs[name + "/bias"] = (cout,)
    meta[name + "/bias"] = ("bias", act, False)


def spp_bins(levels, ndims):
    n = 0
    for lv in levels:
        if isinstance(lv, int):
            n += lv ** ndims
        elif len(lv) == 1:
            n += lv[0] ** ndims
        else:
            n += int(np.prod(lv))
    return n


def hpnn_weight_specs(cfg, prefix=""):
    """Variables of Homogeneous_Poisson_NN_Legacy (poisson_CNN/models/Homogeneous_Poisson_NN_Legacy.py:11-115)."""
    specs, meta = {}, {}
    use_bn = cfg.get("use_batchnorm", False)
    pre = cfg["pre_bottleneck_convolutions_config"]
    cin = 3 if cfg.get("use_positional_embeddings", True) else 1
    act = activation_enum(pre.get("activation"))
    for k, (f, ks) in enumerate(zip(pre["filters"], pre["kernel_sizes"])):
        _conv(specs, meta, "pre_bottleneck/%d" % k, ks, cin, f, act, pre.get("use_bias", True))
        if use_bn:
            _bn(specs, meta, "pre_bottleneck/%d/bn" % k, f)
        cin = f
    c0 = cin
    for kind in ("deconv", "multilinear"):
        bc = cfg["bottleneck_%s_config" % kind]
        F = bc["filters"]
        cact = activation_enum(bc.get("conv_activation"))
        for i in range(len(bc["downsampling_factors"])):
            name = "bottleneck_%s/%d" % (kind, i)
            ks = bc["conv_kernel_sizes"][i]
            _conv(specs, meta, name + "/conv0", ks, c0, F, cact, bc.get("conv_use_bias", True))
            for r in range(1, bc["n_convs"][i]):
                _resnet(specs, meta, "%s/resnet%d" % (name, r), ks, F, cact, use_bn, bc.get("conv_use_bias", True))
            if kind == "deconv":
                dk = bc["deconv_kernel_sizes"][i]
                specs[name + "/deconv/kernel"] = (dk, dk, F, F)
                meta[name + "/deconv/kernel"] = ("deconv", activation_enum(bc.get("deconv_activation")), False)
                if bc.get("deconv_use_bias", True):
                    specs[name + "/deconv/bias"] = (F,)
                    meta[name + "/deconv/bias"] = ("bias", 0, False)
    F = cfg["bottleneck_deconv_config"]["filters"]
    _conv(specs, meta, "non_bottleneck_conv", 5, c0, F, ACT_LEAKY_RELU)
    _conv(specs, meta, "post_merge_conv", 7, 2 * F, F, ACT_LEAKY_RELU)
    _resnet(specs, meta, "post_merge_resnet", 7, F, ACT_LEAKY_RELU, False)
    for i, (a, b) in enumerate(((3, 100), (100, 100), (100, F))):
        _dense(specs, meta, "dx_dense/%d" % i, a, b, ACT_LEAKY_RELU if i < 2 else 0)
    fin = cfg["final_convolutions_config"]
    n_reg = fin.get("final_regular_conv_stages", 2)
    fact = activation_enum(fin.get("activation"))
    cin = F
    S = len(fin["filters"])
    for k in range(S):
        f, ks = fin["filters"][k], fin["kernel_sizes"][k]
        if k < S - n_reg:
            _conv(specs, meta, "final/%d/conv" % k, ks, cin, f, fact, fin.get("use_bias", True))
            _resnet(specs, meta, "final/%d/resnet" % k, ks, f, fact, False, fin.get("use_bias", True))
        else:
            _conv(specs, meta, "final/%d/conv" % k, ks, cin, f, 0, fin.get("use_bias", True))
        cin = f
    if cfg.get("use_scaling", False):
        sc = cfg["scaling_config"]
        sact = activation_enum(sc.get("activation"))
        cin = cfg["final_convolutions_config"]["filters"][-1] + 1
        for s in range(sc.get("stages", 2)):
            _conv(specs, meta, "scaling/conv%d" % s, sc["kernel_size"], cin, sc["filters"], sact, sc.get("use_bias", True))
            cin = sc["filters"]
        nb = spp_bins(sc.get("spp_levels", [[2, 2], 3, 5]), 2)
        for i, (a, b) in enumerate(((nb, 100), (100, 25), (25, 1))):
            _dense(specs, meta, "scaling/dense%d" % i, a, b, ACT_LEAKY_RELU if i < 2 else 0)
    return ({prefix + k: v for k, v in specs.items()}, {prefix + k: v for k, v in meta.items()})


def dbcnn_weight_specs(cfg, prefix=""):
    """Variables of Dirichlet_BC_NN_Legacy_2 (poisson_CNN/models/Dirichlet_BC_NN_Legacy.py:15-101)."""
    specs, meta = {}, {}
    use_bn = cfg.get("use_batchnorm", False)
    bc = cfg["boundary_conv_config"]
    act = activation_enum(bc.get("activation"))
    cin = 3
    for k, (f, ks) in enumerate(zip(bc["filters"], bc["kernel_sizes"])):
        _conv(specs, meta, "boundary/%d/conv" % k, ks, cin, f, act, bc.get("use_bias", True), ndim=1)
        if use_bn:
            _bn(specs, meta, "boundary/%d/bn" % k, f)
        _resnet(specs, meta, "boundary/%d/resnet" % k, ks, f, act, use_bn, bc.get("use_bias", True), ndim=1)
        cin = f
    mlp = cfg["domain_info_mlp_config"]
    din = 3 + spp_bins(cfg["spp_config"]["levels"], 1)
    for i, u in enumerate(mlp["units"]):
        _dense(specs, meta, "mlp/%d" % i, din, u, activation_enum(mlp["activations"][i]))
        din = u
    fin = cfg["final_convolutions_config"]
    n_reg = fin.get("final_regular_conv_stages", 2)
    fact = activation_enum(fin.get("activation"))
    cin = mlp["units"][-1] + 2
    S = len(fin["filters"])
    for k in range(S):
        f, ks = fin["filters"][k], fin["kernel_sizes"][k]
        if k < S - n_reg:
            _conv(specs, meta, "final/%d/conv" % k, ks, cin, f, fact, fin.get("use_bias", True))
            _resnet(specs, meta, "final/%d/resnet" % k, ks, f, fact, False, fin.get("use_bias", True))
        else:
            _conv(specs, meta, "final/%d/conv" % k, ks, cin, f, ACT_TANH, fin.get("use_bias", True))
        cin = f
    return ({prefix + k: v for k, v in specs.items()}, {prefix + k: v for k, v in meta.items()})


def synthetic_weights(specs_meta, seed=0, dtype=np.float32):
    """Seeded random weights that keep activations O(1) through the ~45-conv-deep chain:
    kernels N(0, g/fan_in) with g = 2/(1+0.2^2) for leaky_relu layers and 1 otherwise, the last
    conv of every resnet scaled by 1/sqrt(2) (it consumes x + f(x)); bias U(-0.1,0.1);
    BN gamma U(0.8,1.2), beta/mean U(-0.1,0.1), var U(0.5,1.5).  Each tensor has its own
    stream keyed by (seed, crc32(name)) so values do not depend on iteration order."""
    specs, meta = specs_meta
    out = {}
    for name, shape in specs.items():
        rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
        kind, act, tail = meta[name]
        if kind in ("conv", "dense", "deconv"):
            if kind == "dense":
                fan_in = shape[0]
            elif kind == "deconv":
                fan_in = shape[-1]          # k == stride: exactly one tap per output pixel
            else:
                fan_in = int(np.prod(shape[:-1]))
            gain = 2.0 / (1.0 + 0.2 ** 2) if act == ACT_LEAKY_RELU else 1.0
            if tail:
                gain *= 0.5
            w = rng.standard_normal(shape) * math.sqrt(gain / fan_in)
        elif kind == "bias":
            w = rng.uniform(-0.1, 0.1, shape)
        elif kind == "bn_gamma":
            w = rng.uniform(0.8, 1.2, shape)
        elif kind in ("bn_beta", "bn_mean"):
            w = rng.uniform(-0.1, 0.1, shape)
        elif kind == "bn_var":
            w = rng.uniform(0.5, 1.5, shape)
        else:
            raise ValueError(kind)
        out[name] = np.ascontiguousarray(w, dtype=dtype)
    return out


def count_parameters(specs):
    return int(sum(int(np.prod(s)) for s in specs.values()))


def save_npz(path, weights):
    np.savez(path, **{k.replace("/", "."): v for k, v in weights.items()})


def load_npz(path):
    with np.load(path) as f:
        return {k.replace(".", "/"): f[k] for k in f.files}
