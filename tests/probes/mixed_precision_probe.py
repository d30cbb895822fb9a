"""Error of mixed precision assignments (HPNN mode, DBCNN mode) vs the float64 oracle on bench-like inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from oracle import poisson_oracle as O
from poisson_cnn_b200.synthetic import make_problem
dev = torch.device("cuda", 0)
model, (hp_cfg, db_cfg, w) = bench.build_model(dev, "tc2")
for seed, n in ((1001, 256), (7, 256), (11, 192)):
    p = make_problem(2, n, n, seed=seed)
    ref = O.pcnn_forward(hp_cfg, db_cfg, w, *[p[k].double() for k in bench.KEYS])
    inp = [p[k].cuda() for k in bench.KEYS]
    for hm, dm in (("tc2", "tc2"), ("tc2", "tc"), ("tc", "tc2"), ("tc", "tc")):
        model.hpnn.set_precision(hm); model.dbcnn.set_precision(dm)
        out = model(inp).double().cpu()
        print("seed %d n %d hpnn %s dbcnn %s: rel-L2 %.3e" % (seed, n, hm, dm, float((out - ref).norm() / ref.norm())), flush=True)
