"""Shared weight handling of the hot-path models."""
import os

import numpy as np
import torch

from .. import weights as W

BN_EPS = 1e-3   # tf.keras.layers.BatchNormalization default epsilon


class WeightedModel:
    """Holds name-addressed device tensors; BatchNorm statistics are folded to (scale, shift) once.

    The reference creates its variables on the first call and fills them with
    model.load_weights(checkpoint_prefix) (poisson_CNN/train/utils.py:12-15).  Here
    load_weights() accepts the prefix of such a TF checkpoint (pure-Python reader, tf_checkpoint.py), an .npz
    written by weights.save_npz(), or a {name: array} dict.
    """

    def __init__(self):
        self._w = {}
        self._bn = {}
        self._tc = {}
        self.device = None
        self.precision = "fp32"
        self.requested_precision = "fp32"
        # The forward pass runs inside libpcnn.so (model-level C ABI, csrc/engine.cu).  PCNN_PY_PROGRAM=1 (or
        # use_engine = False) drives the same kernels op by op from Python instead: the development / probing path.
        self.use_engine = os.environ.get("PCNN_PY_PROGRAM", "0") != "1"
        self._engine = None

    # -- to be provided by subclasses
    def weight_specs(self, prefix=""):
        raise NotImplementedError

    def load_weights(self, source, prefix="", device=None):
        if isinstance(source, str):
            source = self._read_weight_file(source, prefix)
        if device is None:
            device = self.device or torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        specs, _ = self.weight_specs(prefix)
        missing = [k for k in specs if k not in source]
        if missing:
            raise ValueError("load_weights: %d variables missing, e.g. %s" % (len(missing), missing[:3]))
        self._w, self._bn = {}, {}
        for name, shape in specs.items():
            a = np.asarray(source[name], dtype=np.float32)
            if tuple(a.shape) != tuple(shape):
                raise ValueError("load_weights: %s has shape %s, expected %s" % (name, a.shape, shape))
            self._w[name[len(prefix):]] = torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        # fold BN: y = gamma*(x-mean)/sqrt(var+eps)+beta = x*scale + shift
        for name in list(self._w):
            if name.endswith("/gamma"):
                base = name[:-len("/gamma")]
                g, b, m, v = (self._w[base + "/" + k] for k in ("gamma", "beta", "mean", "var"))
                scale = g / torch.sqrt(v + BN_EPS)
                self._bn[base] = (scale.contiguous(), (b - m * scale).contiguous())
        self._on_weights_loaded()
        return self

    def _on_weights_loaded(self):
        self._tc = {}
        if self._engine is not None:
            self._engine.close()
        self._engine = None

    # -- model-level C ABI
    def _engine_config(self):
        """The JSON handed to pcnn_create (sections of the reference's experiment files)."""
        raise NotImplementedError

    def _engine_weights(self):
        """{name: numpy array} with the 'hpnn/' / 'dbcnn/' prefixes the engine expects."""
        raise NotImplementedError

    def engine(self):
        """The pcnn_handle of this model (created on first use, re-finalised when the precision mode changes)."""
        from ..engine import Engine
        if self.device is None:
            raise RuntimeError("model has no weights: call load_weights() or init_synthetic_weights() first")
        if self._engine is None:
            self._engine = Engine(self._engine_config(), self.device).set_weights(self._engine_weights())
        if self._engine.precision != self.requested_precision:
            self._engine.finalize(self.requested_precision)
        return self._engine

    def keras_key_map(self, prefix=""):
        """{reference Keras attribute path: variable name} (tf_checkpoint.py); provided by the model classes."""
        raise NotImplementedError

    def _read_weight_file(self, path, prefix=""):
        """`path` is either an .npz written by weights.save_npz() or the PREFIX of a TensorFlow checkpoint as
        written by the reference (model.save_weights / ModelCheckpoint(save_weights_only=True)):
        <prefix>.index + <prefix>.data-00000-of-00001, read by the pure-Python tensor-bundle reader."""
        import os
        if os.path.isfile(path + ".index"):
            from .. import tf_checkpoint as T
            return T.load_checkpoint_weights(path, self.keras_key_map(prefix))
        return W.load_npz(path)

    PRECISIONS = ("fp32", "tc", "tc2", "tc3", "mixed")
    # modes that hold the north-star budgets for THIS network (1e-5 strict, 2e-3 tensor-core); overridden by the HPNN
    COMPLIANT_PRECISIONS = ("fp32", "tc", "tc2", "tc3", "mixed")
    _TC_MODE = {"tc": 1, "tc3": 2, "tc2": 3}
    _MIXED_AS = "tc2"      # what 'mixed' means for this network on its own (the DBCNN overrides it with 'tc')

    def set_precision(self, precision):
        """'fp32': strict FP32 CUDA-core kernels (rel-L2 <= 1e-5 vs the reference arithmetic);
        'tc'  : tcgen05 tensor cores, one pass of FP16 operands (11-bit significand like TF32), FP32 accumulation;
        'tc3' : tcgen05 with split FP16 operands (x = hi + lo, W = hi + lo; three MMAs per product term),
                ~22 significand bits -- the error-compensated mode that holds the 2e-3 budget on any input;
        'tc2' : FP16 main pass + ONE e4m3 K=32 MMA for both correction terms (2x the tensor work of 'tc',
                ~15 significand bits);
        'mixed': 'tc2' in the HPNN (45 convolutions deep: it carries essentially all of the rounding error of the
                merged model) and single-pass 'tc' in the DBCNN (shallower, tanh-bounded and max-normalised:
                measured contribution < 1e-4) -- same accuracy as 'tc2' at ~0.8x its tensor work."""
        if precision not in self.PRECISIONS:
            raise ValueError("precision must be one of %s" % (self.PRECISIONS,))
        if precision not in self.COMPLIANT_PRECISIONS:
            import warnings
            warnings.warn("%s: precision %r is a speed/diagnostic mode OUTSIDE the 2e-3 tensor-core error budget for this "
                          "network (single fp16 pass through 45 convolutions: up to 4.4e-3 at 256x256); use 'mixed', "
                          "'tc2' or 'tc3' for compliant results" % (type(self).__name__, precision), UserWarning, stacklevel=2)
        subs = [getattr(self, sub) for sub in ("hpnn", "dbcnn") if hasattr(self, sub)]
        # a single network resolves 'mixed' to its own mode; the merged model keeps the name and hands it down
        self.precision = precision if (subs or precision != "mixed") else self._MIXED_AS
        self.requested_precision = precision
        for sub in subs:
            sub.set_precision(precision)
        return self

    def tc_conv(self, name, mode=None):
        """(packed fp16 operand image, bias) of a conv layer for the tensor-core kernel, packed once per
        precision mode (default: the model's)."""
        from .. import ops
        nsplit = mode or self._TC_MODE[self.precision]
        key = (name, nsplit)
        if key not in self._tc:
            self._tc[key] = ops.pack_conv_weights_tc(self._w[name + "/kernel"], nsplit)
        return self._tc[key], self._w.get(name + "/bias")

    @property
    def tc_split(self):
        """precision mode of the BLK8 tensors (ops.Blk8 split argument): 1, 2 (tc3) or 3 (tc2)"""
        return self._TC_MODE.get(self.precision, 1)

    def init_synthetic_weights(self, seed=0, device=None):
        """Seeded random weights (no trained weights ship with the reference)."""
        return self.load_weights(W.synthetic_weights(self.weight_specs(), seed=seed), device=device)

    def w(self, name):
        try:
            return self._w[name]
        except KeyError:
            if not self._w:
                raise RuntimeError("model has no weights: call load_weights() or init_synthetic_weights() first")
            raise

    def conv(self, name):
        return self._w[name + "/kernel"], self._w.get(name + "/bias")

    def bn(self, name):
        return self._bn.get(name)

    def get_weights_dict(self, prefix=""):
        return {prefix + k: v.detach().cpu().numpy() for k, v in self._w.items()}

    def count_params(self):
        return W.count_parameters(self.weight_specs()[0])
