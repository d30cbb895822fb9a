"""Single-grid spatial decomposition: one large Neumann-HPNN problem (the pressure-projection use case) on N GPUs.
  python scripts/spatial_bench.py [n] [precision]                                  # N = 1: the engine, the op-by-op program and 1 band
  python -m torch.distributed.run --nproc-per-node N scripts/spatial_bench.py ...  # N bands, one per GPU
Prints one JSON line on rank 0: ms per forward (CUDA events, max over ranks), bit-equality with the single-GPU result."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from poisson_cnn_b200 import convert_tf_object_names, load_experiment, models, weights as W
from poisson_cnn_b200.sharding import init_from_env
from poisson_cnn_b200.spatial import SpatialHPNN
from poisson_cnn_b200.synthetic import make_problem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
prec = sys.argv[2] if len(sys.argv) > 2 else "mixed"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
rank, world, local = init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
cfg = load_experiment("hpnn_neumann")["model"]
w = W.synthetic_weights(W.hpnn_weight_specs(cfg, "hpnn/"), seed=0)
m = models.Homogeneous_Poisson_NN_Legacy(**convert_tf_object_names(cfg)).load_weights(w, "hpnn/", device=dev).set_precision(prec)
p = make_problem(1, n, n, seed=5, magnitudes=False)
rhs, dx = p["rhs"].to(dev), p["dx"].to(dev)


def timed(fn, reps=5):
    for _ in range(2):
        out = fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t), out


res = {"grid": n, "precision": prec, "n_gpus": world}
t_eng, ref = timed(lambda: m([rhs, dx]))
res["engine_1gpu_ms"] = t_eng
sp = SpatialHPNN(m) if world > 1 else SpatialHPNN(m, world=1)
t_sp, out = timed(lambda: sp([rhs, dx]))
res["spatial_ms"] = t_sp
res["speedup_vs_engine_1gpu"] = t_eng / t_sp
res["bit_identical"] = bool(torch.equal(out, ref))
sp.profile = {}
sp([rhs, dx])
res["phases_ms_rank0"] = {k: round(v, 2) for k, v in sp.profile.items()}
sp.profile = None
if world == 1:
    m.use_engine = False
    t_py, _ = timed(lambda: m([rhs, dx]))
    res["python_program_1gpu_ms"] = t_py
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
