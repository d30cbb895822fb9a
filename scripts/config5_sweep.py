"""BASELINE config 5: GPU DST reference solve vs CNN surrogate, 64^2 .. 2048^2 (one B200).
Throughput of both, and for the DST its own discrete residual (it is the ground truth); the CNN runs on seeded
synthetic weights (no trained weights ship with the reference), so its distance to the DST solution measures
nothing about the method -- it is reported only to exercise the accuracy harness end to end."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from poisson_cnn_b200 import dataset
from poisson_cnn_b200.solvers import dst_poisson_solve
from poisson_cnn_b200.losses import linear_operator_loss
dev = torch.device("cuda", 0)
model, _ = bench.build_model(dev, "mixed")
loss = linear_operator_loss(3, 2, ndims=2)
print("grid | batch | generator ms | DST ms | DST sol/s | DST GB/s (8 B/pt) | DST residual rel | CNN ms | CNN sol/s | CNN vs DST rel-L2")
for n, B in ((64, 256), (128, 256), (256, 256), (512, 64), (1024, 16), (2048, 4)):
    # problems generated on the GPU (poisson_cnn_b200.dataset: reference dataset/generators/numerical.py): B distinct fields
    gen = torch.Generator(device=dev).manual_seed(1005)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rhs = dataset.generate_random_RHS(B, [n, n], smoothness=6, max_magnitude=1.0, device=dev, generator=gen)
    bnd = dataset.generate_random_boundaries([n, n], batch_size=B, smoothness=5, max_magnitude={k: 1.0 for k in ("left", "top", "right", "bottom")},
                                             return_with_expanded_dims=True, device=dev, generator=gen)
    dx = 5e-3 + (5e-2 - 5e-3) * torch.rand((B, 1), device=dev, generator=gen)
    e1.record(); torch.cuda.synchronize()
    t_gen = e0.elapsed_time(e1)
    left, top, right, bottom = bnd["left"], bnd["top"], bnd["right"], bnd["bottom"]
    inp = [rhs, left, top, right, bottom, dx]
    u = dst_poisson_solve(rhs, bnd, dx); torch.cuda.synchronize()
    e0.record(); u = dst_poisson_solve(rhs, bnd, dx); e1.record(); torch.cuda.synchronize()
    t_dst = e0.elapsed_time(e1)
    gs = torch.cat([dx, dx], 1)
    res = float(loss(rhs, u, gs)) ** 0.5 / float((rhs[..., 1:-1, 1:-1] ** 2).mean()) ** 0.5
    try:
        out = model(inp); torch.cuda.synchronize()
        e0.record(); out = model(inp); e1.record(); torch.cuda.synchronize()
        t_cnn = e0.elapsed_time(e1)
        err = float((out.double() - u.double()).norm() / u.double().norm())
        cnn = "%8.1f | %8.1f | %.2e" % (t_cnn, B * 1e3 / t_cnn, err)
    except Exception as e:      # e.g. 64^2: the shipped Scaling config has empty SPP bins on tiny maps (NaN in the reference too)
        cnn = "n/a (%s)" % type(e).__name__
    print("%4d^2 | %3d | %6.2f | %8.2f | %9.1f | %7.1f | %.1e | %s" % (n, B, t_gen, t_dst, B * 1e3 / t_dst, B * n * n * 8 / t_dst / 1e6, res, cnn), flush=True)
