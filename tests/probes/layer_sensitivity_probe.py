"""Error of the merged model (HPNN tc2, DBCNN tc) when groups of HPNN layers skip their correction pass."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from oracle import poisson_oracle as O
from poisson_cnn_b200.synthetic import make_problem
dev = torch.device("cuda", 0)
model, (hp_cfg, db_cfg, w) = bench.build_model(dev, "mixed")
groups = {
    "none": (),
    "bottleneck branches": ("bottleneck_",),
    "pre_bottleneck": ("pre_bottleneck",),
    "non_bottleneck+post_merge_conv": ("non_bottleneck_conv", "post_merge_conv"),
    "post_merge_resnet": ("post_merge_resnet",),
    "final/0 (k15)": ("final/0/",),
    "final/1 (k13)": ("final/1/",),
    "final/2-3 (k9,k7)": ("final/2/", "final/3/"),
    "final/4+ (k<=5)": ("final/4/", "final/5/", "final/6/", "final/7/", "final/8/"),
    "all": ("",),
}
probs = []
for seed, n in ((1001, 256), (7, 256)):
    p = make_problem(2, n, n, seed=seed)
    ref = O.pcnn_forward(hp_cfg, db_cfg, w, *[p[k].double() for k in bench.KEYS])
    probs.append(([p[k].cuda() for k in bench.KEYS], ref))
for name, pf in groups.items():
    model.hpnn.tc_uncorrected = pf
    errs = []
    for inp, ref in probs:
        out = model(inp).double().cpu()
        errs.append(float((out - ref).norm() / ref.norm()))
    print("%-34s rel-L2 %s" % (name, "  ".join("%.3e" % e for e in errs)), flush=True)
