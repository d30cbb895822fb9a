"""The model-level C ABI driven by a host written in plain C (examples/c_host/pcnn_host.c: gcc + libcudart, no Python, no
torch in the process): config JSON, weights and inputs from files, pcnn_create -> pcnn_set_weight -> pcnn_finalize_weights ->
pcnn_workspace_bytes -> pcnn_forward, result compared with the oracle's golden vector and with the Python host's result."""
import os
import struct
import subprocess

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN, ROOT, pcnn_configs, all_weights, rel_l2

pytestmark = pytest.mark.gpu
KEYS = ("rhs", "left", "top", "right", "bottom", "dx")
HOST = os.path.join(ROOT, "build", "pcnn_host")


@pytest.mark.skipif(not os.path.isfile(HOST), reason="build/pcnn_host not built (make)")
@pytest.mark.parametrize("precision,tol", [(0, 1e-5), (4, 2e-3)])
def test_plain_c_host_runs_the_forward_pass(tmp_path, precision, tol):
    hp, db = pcnn_configs()
    w = all_weights(hp, db)
    wpath, ipath, opath = (str(tmp_path / n) for n in ("weights.bin", "inputs.bin", "out.bin"))
    with open(wpath, "wb") as f:
        for name, a in w.items():
            a = np.ascontiguousarray(a, dtype=np.float32)
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)) + nb + struct.pack("<I", a.ndim) + struct.pack("<%dq" % a.ndim, *a.shape) + a.tobytes())
    g = np.load(os.path.join(GOLDEN, "pcnn_112x120.npz"))
    B, _, H, W = g["rhs"].shape
    with open(ipath, "wb") as f:
        for k in KEYS:
            f.write(np.ascontiguousarray(g[k], dtype=np.float32).tobytes())
    cfg = os.path.join(ROOT, "poisson_cnn_b200", "experiments", "pcnn_end_to_end.json")      # the reference's experiment file, as is
    res = subprocess.run([HOST, cfg, wpath, ipath, str(B), str(H), str(W), str(precision), opath], capture_output=True, text=True, timeout=300)
    print(res.stdout.strip(), res.stderr.strip())
    assert res.returncode == 0, res.stderr
    out = np.fromfile(opath, dtype=np.float32).reshape(B, 1, H, W)
    assert rel_l2(out, g["out"]) < tol
    # the Python host (same library, same layer program) gives the same bits
    from poisson_cnn_b200 import convert_tf_object_names, models
    m = models.Poisson_CNN_Legacy(models.Homogeneous_Poisson_NN_Legacy(**convert_tf_object_names(hp)),
                                  models.Dirichlet_BC_NN_Legacy_2(**convert_tf_object_names(db))).load_weights(w)
    ref = m.set_precision({0: "fp32", 4: "mixed"}[precision])([torch.from_numpy(g[k]).cuda() for k in KEYS])
    assert np.array_equal(ref.cpu().numpy(), out)
    # a config error surfaces as the reference's ValueError text on stderr
    bad = str(tmp_path / "bad.json")
    open(bad, "w").write('{"hpnn_model": {"bc_type": "robin"}}')
    res = subprocess.run([HOST, bad, wpath, ipath, str(B), str(H), str(W), "0", opath], capture_output=True, text=True, timeout=60)
    assert res.returncode != 0 and "bc_type can only be neumann or dirichlet." in res.stderr
