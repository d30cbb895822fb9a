"""poisson_cnn_b200 -- B200-native batched inference for aligirayhanozbay/poisson_CNN's hot path.

Importing the package never touches the GPU; the CUDA library (csrc -> libpcnn.so) is loaded on
first use by `poisson_cnn_b200._lib` and raises if it is missing (there is no CPU fallback).
"""
from .config import convert_tf_object_names, load_experiment  # noqa: F401
from . import config, weights  # noqa: F401

__version__ = "0.1.0"
