"""Config plumbing of the hot path (host side, pure Python).

Mirrors the reference helpers
  poisson_CNN/utils/convert_tf_object_names.py:3-21
  poisson_CNN/models/Homogeneous_Poisson_NN_Metalearning.py:10-57
without TensorFlow: the reference `eval`s strings such as "tf.nn.leaky_relu" into TF
callables; here they are mapped to activation enums understood by the CUDA kernels.
"""
import copy
import json
import os

# activation enums shared with csrc/pcnn_common.cuh
ACT_LINEAR, ACT_LEAKY_RELU, ACT_TANH = 0, 1, 2
PAD_CONSTANT, PAD_SYMMETRIC, PAD_REFLECT = 0, 1, 2
RESIZE_NEAREST, RESIZE_BILINEAR, RESIZE_BICUBIC = 0, 1, 2
RESIZE_BICUBIC_LEGACY_AC = 3   # tf.compat.v1 resize_images(BICUBIC, align_corners=True): the dataset generators' resize

_ACTS = {
    "tf.nn.leaky_relu": ACT_LEAKY_RELU, "leaky_relu": ACT_LEAKY_RELU,
    "tf.nn.tanh": ACT_TANH, "tf.tanh": ACT_TANH, "tf.math.tanh": ACT_TANH, "tanh": ACT_TANH,
    "tf.keras.activations.tanh": ACT_TANH,
    "linear": ACT_LINEAR, "tf.keras.activations.linear": ACT_LINEAR, "None": ACT_LINEAR,
}
_PADS = {"CONSTANT": PAD_CONSTANT, "SYMMETRIC": PAD_SYMMETRIC, "REFLECT": PAD_REFLECT}
_RESIZE = {"nearest": RESIZE_NEAREST, "bilinear": RESIZE_BILINEAR, "bicubic": RESIZE_BICUBIC}


class Activation(str):
    """A string that stands in for the TF callable the reference would have eval'd."""
    @property
    def enum(self):
        return activation_enum(self)


def activation_enum(a):
    if a is None:
        return ACT_LINEAR
    if isinstance(a, int):
        return a
    name = a if isinstance(a, str) else getattr(a, "__name__", str(a))
    if name in _ACTS:
        return _ACTS[name]
    for key in ("leaky_relu", "tanh", "linear"):
        if key in name:
            return _ACTS[key]
    raise ValueError("unsupported activation: %r (supported: leaky_relu, tanh, linear)" % (a,))


def padding_enum(mode):
    m = str(mode).upper()
    if m not in _PADS:
        raise ValueError("unsupported padding mode %r" % (mode,))
    return _PADS[m]


def resize_enum(method):
    m = str(method).lower()
    if m not in _RESIZE:
        raise ValueError("unsupported resize method %r (nearest, bilinear, bicubic)" % (method,))
    return _RESIZE[m]


def convert_tf_object_names(x):
    """Same traversal as the reference (lists/dicts, strings containing 'tf.'), but a 'tf.*'
    string becomes an `Activation` marker instead of being eval'd."""
    if isinstance(x, list):
        return [Activation(i) if (isinstance(i, str) and "tf." in i) else
                convert_tf_object_names(i) if isinstance(i, (list, dict)) else i for i in x]
    if isinstance(x, dict):
        return {k: (Activation(v) if (isinstance(v, str) and "tf." in v) else
                    convert_tf_object_names(v) if isinstance(v, (list, dict)) else v) for k, v in x.items()}
    raise ValueError("The input must be a list or dict")


def get_init_arguments_from_config(cfg, k, fields_in_cfg, fields_in_args):
    """cfg={'a':3,'b':[0,1,2]}, k=2, fields b->bp  =>  {'a':3,'bp':2}."""
    out = {key: cfg[key] for key in cfg if key not in fields_in_cfg}
    out.update({a: cfg[c][k] for a, c in zip(fields_in_args, fields_in_cfg)})
    return out


def process_normalizations(normalizations):
    types, defaults = ["rhs_max_magnitude"], [False]
    if normalizations is None:
        return dict(zip(types, defaults))
    if isinstance(normalizations, dict):
        for key, d in zip(types, defaults):
            normalizations.setdefault(key, d)
        if isinstance(normalizations["rhs_max_magnitude"], bool) and normalizations["rhs_max_magnitude"]:
            normalizations["rhs_max_magnitude"] = 1.0
    return normalizations


def process_output_scaling_modes(output_scalings):
    modes = ["rhs_max_magnitude", "max_domain_size_squared", "match_peak_laplacian_magnitude_to_peak_rhs", "soln_max_magnitude"]
    output_scalings = copy.deepcopy(output_scalings)
    if output_scalings is None:
        return {m: False for m in modes}
    if isinstance(output_scalings, dict):
        for m in modes:
            output_scalings.setdefault(m, False)
    return output_scalings


def experiments_dir():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "experiments")


def load_experiment(name):
    """Load one of the shipped architecture JSONs (e.g. 'pcnn_end_to_end')."""
    path = name if os.path.isfile(name) else os.path.join(experiments_dir(), name + ("" if name.endswith(".json") else ".json"))
    with open(path) as f:
        cfg = json.load(f)
    cfg.pop("_comment", None)
    return cfg


def bottleneck_output_size(n, ds, us):
    """poisson_CNN/blocks/bottleneck_block.py:82 -- tf.cast((n/ds)*us, int32), float64 arithmetic."""
    return int((n / ds) * us)
