"""Shared test helpers: model construction, oracle bridging, error metrics."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from poisson_cnn_b200 import load_experiment, convert_tf_object_names, weights as W  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def rel_l2(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm())


def pcnn_configs(small_scaling=False):
    cfg = load_experiment("pcnn_end_to_end")
    hp, db = cfg["hpnn_model"], cfg["dbcnn_model"]
    if small_scaling:
        hp = load_experiment("hpnn_smalldomain")["model"]
    return hp, db


def all_weights(hp_cfg, db_cfg, seed=0):
    hs = W.hpnn_weight_specs(hp_cfg, "hpnn/")
    ds = W.dbcnn_weight_specs(db_cfg, "dbcnn/")
    return W.synthetic_weights(({**hs[0], **ds[0]}, {**hs[1], **ds[1]}), seed=seed)
