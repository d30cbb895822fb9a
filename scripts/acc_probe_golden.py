import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from poisson_cnn_b200.synthetic import make_problem
dev = torch.device("cuda", 0)
model, _ = bench.build_model(dev, "mixed")
g = np.load("/root/repo/tests/golden/pcnn_112x120.npz")
keys = ("rhs", "left", "top", "right", "bottom", "dx")
out = model([torch.from_numpy(g[k]).cuda() for k in keys])
ref = torch.from_numpy(g["out"]).double()
print("UM_TC=%s SEP=%s NOPAIR=%s: golden 112x120 mixed rel-L2 %.3e" % (os.environ.get("PCNN_UM_TC", "1"), os.environ.get("PCNN_DBCNN_SEPARABLE", "1"), os.environ.get("PCNN_TC_NO_PAIR", "0"),
      float((out.double().cpu() - ref).norm() / ref.norm())))
