"""Generates the forward-pass golden vectors with the CPU oracle (float64), seeded synthetic weights
(poisson_cnn_b200.weights.synthetic_weights, seed 0) and seeded inputs.  The reference ships neither
weights nor golden vectors, so these pin the ORACLE, not TensorFlow.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.helpers import pcnn_configs, all_weights  # noqa: E402
from poisson_cnn_b200.synthetic import make_problem  # noqa: E402
from oracle import poisson_oracle as O  # noqa: E402

torch.set_num_threads(8)
d = lambda t: t.double()

# (a) HPNN, small-domain scaling block, 64x64 (BASELINE config 1 shape), dirichlet
hp_s, db = pcnn_configs(small_scaling=True)
w = all_weights(hp_s, db)
p = make_problem(2, 64, 64, seed=1001, magnitudes=False)
out = O.hpnn_forward(hp_s, w, d(p["rhs"]), d(p["dx"]), "hpnn/")
np.savez_compressed(os.path.join(HERE, "hpnn_64x64.npz"), rhs=p["rhs"].numpy(), dx=p["dx"].numpy(), out=out.float().numpy())

# (b) DBCNN, n=48 boundary points expanded to x_res=56
p = make_problem(2, 56, 48, seed=1002, magnitudes=False)
out = O.dbcnn_forward(db, w, d(p["left"]), d(p["dx"]), 56, "dbcnn/")
np.savez_compressed(os.path.join(HERE, "dbcnn_56x48.npz"), bc=p["left"].numpy(), dx=p["dx"].numpy(), out=out.float().numpy())

# (c) full PCNN, shipped config, 112x120 (>= 109 per side so the shipped Scaling SPP has no empty bin)
hp, db = pcnn_configs()
w = all_weights(hp, db)
p = make_problem(1, 112, 120, seed=1003)
out = O.pcnn_forward(hp, db, w, *(d(p[k]) for k in ("rhs", "left", "top", "right", "bottom", "dx")))
np.savez_compressed(os.path.join(HERE, "pcnn_112x120.npz"), **{k: v.numpy() for k, v in p.items()}, out=out.float().numpy())
print("golden vectors written")
