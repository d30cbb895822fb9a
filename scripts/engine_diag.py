"""Engine (C++ layer program) vs op-by-op Python program vs oracle golden, per network and precision mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import GOLDEN, pcnn_configs, all_weights, rel_l2
from poisson_cnn_b200 import convert_tf_object_names, models
from poisson_cnn_b200.synthetic import make_problem
KEYS = ("rhs", "left", "top", "right", "bottom", "dx")
hp, db = pcnn_configs(); w = all_weights(hp, db)
def build(py):
    m = models.Poisson_CNN_Legacy(models.Homogeneous_Poisson_NN_Legacy(**convert_tf_object_names(hp)),
                                  models.Dirichlet_BC_NN_Legacy_2(**convert_tf_object_names(db))).load_weights(w)
    if py: m.use_engine = m.hpnn.use_engine = m.dbcnn.use_engine = False
    return m
me, mp = build(False), build(True)
g = np.load(os.path.join(GOLDEN, "pcnn_112x120.npz"))
inp = [torch.from_numpy(g[k]).cuda() for k in KEYS]
p = make_problem(2, 128, 128, seed=7, magnitudes=False)
import warnings; warnings.simplefilter("ignore")
for mode in ("fp32", "tc", "tc2", "tc3", "mixed"):
    me.set_precision(mode); mp.set_precision(mode)
    a, b = me(inp), mp(inp)
    a2, b2 = me(inp), mp(inp)
    ha, hb = me.hpnn([p["rhs"].cuda(), p["dx"].cuda()]), mp.hpnn([p["rhs"].cuda(), p["dx"].cuda()])
    da, dbb = me.dbcnn([p["left"].cuda(), p["dx"].cuda(), 128]), mp.dbcnn([p["left"].cuda(), p["dx"].cuda(), 128])
    print("%-5s pcnn: eng-vs-py %.2e | eng-vs-gold %.2e py-vs-gold %.2e | rerun eng %.1e py %.1e || hpnn eng-vs-py %.2e | dbcnn eng-vs-py %.2e" % (
        mode, rel_l2(a, b), rel_l2(a, g["out"]), rel_l2(b, g["out"]), rel_l2(a2, a), rel_l2(b2, b), rel_l2(ha, hb), rel_l2(da, dbb)), flush=True)
print("workspace B=256 256^2 mixed: %.2f GB" % (me.set_precision("mixed").engine().workspace_bytes("pcnn", 256, 256, 256) / 1e9))
print("workspace B=1 256^2 mixed: %.1f MB" % (me.engine().workspace_bytes("pcnn", 1, 256, 256) / 1e6))
