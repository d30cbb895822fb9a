"""How much of the engine-vs-Python-program difference is the modes' own sensitivity to 1-ulp input changes?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, warnings
warnings.simplefilter("ignore")
from tests.helpers import pcnn_configs, all_weights, rel_l2
from poisson_cnn_b200 import convert_tf_object_names, models
from poisson_cnn_b200.synthetic import make_problem
hp, db = pcnn_configs(); w = all_weights(hp, db)
def build(py):
    m = models.Poisson_CNN_Legacy(models.Homogeneous_Poisson_NN_Legacy(**convert_tf_object_names(hp)),
                                  models.Dirichlet_BC_NN_Legacy_2(**convert_tf_object_names(db))).load_weights(w)
    if py: m.use_engine = m.hpnn.use_engine = m.dbcnn.use_engine = False
    return m
me, mp = build(False), build(True)
p = make_problem(2, 128, 128, seed=7, magnitudes=False)
rhs, dx, bc = p["rhs"].cuda(), p["dx"].cuda(), p["left"].cuda()
eps = torch.nextafter(rhs, rhs * 2)          # +1 ulp on every element
bce = torch.nextafter(bc, bc * 2)
for mode in ("fp32", "tc", "tc2", "tc3"):
    me.set_precision(mode); mp.set_precision(mode)
    h0, h1, he = mp.hpnn([rhs, dx]), mp.hpnn([eps, dx]), me.hpnn([rhs, dx])
    d0, d1, de = mp.dbcnn([bc, dx, 128]), mp.dbcnn([bce, dx, 128]), me.dbcnn([bc, dx, 128])
    print("%-5s hpnn: py(x) vs py(x+1ulp) %.2e | eng vs py %.2e   dbcnn: py(x) vs py(x+1ulp) %.2e | eng vs py %.2e" % (
        mode, rel_l2(h1, h0), rel_l2(he, h0), rel_l2(d1, d0), rel_l2(de, d0)), flush=True)
