#!/usr/bin/env python
"""Headline benchmark: 256x256 Poisson solutions/s of the full Poisson_CNN_Legacy forward
(BASELINE.json configs[1]: homogeneous + Dirichlet-BC networks merged, batch 256 per GPU).

  python bench.py --gpus N --steps K --warmup W                 # this repo (CUDA kernels via the C ABI)
  python bench.py --impl reference --gpus N --steps K --warmup W  # CPU restatement of the reference path
  torchrun --nproc-per-node N bench.py --gpus N ...             # one rank per GPU, weak scaling

Prints ONE JSON line (rank 0).  A step = one forward pass over one batch of synthetic problems.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

FLOP_PER_SOLUTION_256 = 441.43e9      # conv/transpose-conv MACs x2 at true channel counts
METRIC = "256x256 Poisson solutions/sec (full Poisson_CNN forward)"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "source": "fallback"}


def pcnn_flops(nx, ny):
    """Analytic conv FLOPs per solution (scaled from the 256x256 count; weak shape dependence ignored)."""
    return FLOP_PER_SOLUTION_256 * (nx * ny) / 65536.0


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def build_model(device, precision):
    from poisson_cnn_b200 import convert_tf_object_names, load_experiment, models, weights as W
    cfg = load_experiment("pcnn_end_to_end")
    hp_cfg, db_cfg = cfg["hpnn_model"], cfg["dbcnn_model"]
    hs, ds = W.hpnn_weight_specs(hp_cfg, "hpnn/"), W.dbcnn_weight_specs(db_cfg, "dbcnn/")
    w = W.synthetic_weights(({**hs[0], **ds[0]}, {**hs[1], **ds[1]}), seed=0)
    hp = models.Homogeneous_Poisson_NN_Legacy(**convert_tf_object_names(hp_cfg))
    db = models.Dirichlet_BC_NN_Legacy_2(**convert_tf_object_names(db_cfg))
    model = models.Poisson_CNN_Legacy(hp, db).load_weights(w, device=device)
    if hasattr(model, "set_precision"):
        model.set_precision(precision)
    elif precision != "fp32":
        raise SystemExit("precision %s not built" % precision)
    return model, (hp_cfg, db_cfg, w)


KEYS = ("rhs", "left", "top", "right", "bottom", "dx")


def oracle_throughput(hp_cfg, db_cfg, w, nx, ny, budget_s, steps=None, warmup=0):
    """Times the CPU restatement (oracle, torch-CPU fp32, all host threads) on single-sample forwards of
    the same workload; returns (solutions/s, samples timed, threads)."""
    from oracle import poisson_oracle as O
    from poisson_cnn_b200.synthetic import make_problem
    threads = torch.get_num_threads()
    p = make_problem(1, nx, ny, seed=1001)
    args = [p[k] for k in KEYS]
    with torch.no_grad():
        for _ in range(warmup):
            O.pcnn_forward(hp_cfg, db_cfg, w, *args)
        n, t0 = 0, time.perf_counter()
        while True:
            O.pcnn_forward(hp_cfg, db_cfg, w, *args)
            n += 1
            el = time.perf_counter() - t0
            if (steps is not None and n >= steps) or (steps is None and el >= budget_s):
                break
    return n / el, n, threads


def run_reference(args):
    """--impl reference: the reference's own implementation is TensorFlow (not installable here, see
    DESIGN.md); the stand-in is the oracle restatement on the host cores, bounded to one sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it can
    try:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        torch.set_num_threads(os.cpu_count() or 1)
    from poisson_cnn_b200 import load_experiment, weights as W
    cfg = load_experiment("pcnn_end_to_end")
    hp_cfg, db_cfg = cfg["hpnn_model"], cfg["dbcnn_model"]
    hs, ds = W.hpnn_weight_specs(hp_cfg, "hpnn/"), W.dbcnn_weight_specs(db_cfg, "dbcnn/")
    w = W.synthetic_weights(({**hs[0], **ds[0]}, {**hs[1], **ds[1]}), seed=0)
    steps = max(1, min(args.steps, 8))
    val, n, threads = oracle_throughput(hp_cfg, db_cfg, w, args.grid, args.grid, None, steps=steps, warmup=1 if args.warmup else 0)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "solutions/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": 1 if args.warmup else 0, "ms_per_step": 1000.0 / val, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Poisson_CNN_Legacy forward, %dx%d grids, pcnn_end_to_end architecture, 1 sample per step (bounded CPU sample of the batch-256 workload)" % (args.grid, args.grid)},
        "cpu_baseline": {"value": val, "unit": "solutions/s", "cores": threads, "kind": "port",
                         "sample": "%d single-sample forwards, torch-CPU fp32 oracle (reference is TensorFlow: not installable offline)" % n},
        "e2e": {"value": val, "unit": "solutions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    # stdout carries exactly one JSON line: NCCL's version banner (NCCL_DEBUG=VERSION) would land there too
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch.distributed as dist
    from poisson_cnn_b200 import ops, _lib
    from poisson_cnn_b200.sharding import init_from_env
    from poisson_cnn_b200.synthetic import make_problem

    rank, world, local = init_from_env()
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    B, nx, ny = args.batch, args.grid, args.grid
    model, (hp_cfg, db_cfg, w) = build_model(device, args.precision)

    # synthetic problems: a small pool of distinct smooth fields tiled to the batch (host generation of
    # 256 bicubic fields per rank is slow and irrelevant to the measurement); dx differs per sample
    base = make_problem(min(B, 16), nx, ny, seed=1001 + rank)
    reps = -(-B // base["rhs"].shape[0])
    host = {k: v.repeat(reps, *([1] * (v.dim() - 1)))[:B].contiguous().pin_memory() for k, v in base.items()}
    host["dx"] = (5e-3 + (5e-2 - 5e-3) * torch.rand(B, 1, generator=torch.Generator().manual_seed(rank))).pin_memory()
    dev_in = [host[k].to(device, non_blocking=True) for k in KEYS]
    h2d = sum(host[k].numel() * 4 for k in KEYS)
    out_host = torch.empty((B, 1, nx, ny), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        return ms

    # ---------------- device-resident timing ("value") ----------------
    for _ in range(args.warmup):
        out = model(dev_in)
    barrier()
    sampler = ClockSampler(local); sampler.start()
    timer = ops.KernelTimer(*args.roofline_kernel)
    ops.KERNEL_TIMER = timer
    launches0 = _lib.lib.pcnn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        out = model(dev_in)
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1000.0 / args.steps   # CPU time to enqueue one step (no sync inside)
    e1.record()
    barrier()
    ops.KERNEL_TIMER = None
    launches = _lib.lib.pcnn_launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    sampler.stop_flag = True; sampler.join(2)
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1000.0)

    # ---------------- end to end through the public API with host buffers ("e2e") ----------------
    def e2e_step():
        # the public call with HOST buffers: slices of the batch are copied in, solved and copied out on three streams
        model([host[k] for k in KEYS], out=out_host)
    e2e_steps = max(1, min(args.steps, 3)) if ms_per_step > 2000 else args.steps
    e2e_step()
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * B * e2e_steps / (e2e_ms / 1000.0)

    # ---------------- roofline of the dominant kernel (live CUDA-event timing inside the timed region)
    peaks = measured_peaks()
    ks = timer.summary()
    roofline = None
    if ks:
        ach = ks["flops_per_launch"] / (ks["avg_ms"] * 1e-3) / 1e12
        kk = args.roofline_kernel[2]
        hp_mode = {"mixed": "tc2"}.get(args.precision, args.precision)      # precision mode of the HPNN, which owns the timed kernel
        issue_factor = {"tc": 1.0, "tc3": 3.0, "tc2": 2.0}.get(hp_mode, 0.0) * (kk + 3.0) / kk   # MMAs issued per algorithmic MAC
        # DRAM bytes per launch of this kernel from the ncu --set full captures (dram__bytes_read.sum + dram__bytes_write.sum
        # at 32 samples per launch), scaled to the samples per launch here.  tc2: three of the four 32->32 k15 launches of a
        # forward have no residual input (profiles/r01_conv_tc_k15_tc2_b32_full_raw.csv: 303.6 + 228.8 MB), one has
        # (profiles/r01b_conv_tc_k15_tc2_res_b32_full_raw.csv: 512.1 + 237.5 MB; unchanged in the final build,
        # profiles/r01c_k15_tc2_full_raw.csv: 512.5 + 237.3 MB)
        per_sample = {"tc": (151.436032e6 + 94.614272e6) / 32,
                      "tc2": (3 * (303.551488e6 + 228.785920e6) + (512.086272e6 + 237.487872e6)) / 4 / 32}.get(hp_mode)
        samples_per_launch = min(B, max(1, int((getattr(model, "max_microbatch", B) or B) * 65536 // (nx * ny))))
        traffic = per_sample * samples_per_launch if (per_sample and args.roofline_kernel == (32, 32, 15) and (nx, ny) == (256, 256)) else None
        roofline = {"bound": "tensor", "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s",
                    "frac": ach / peaks["tflops"], "traffic": traffic, "kernel": "conv2d %d->%d k%d (%s)" % (args.roofline_kernel + (hp_mode,)),
                    "launches_timed": ks["launches"], "avg_launch_ms": ks["avg_ms"], "peak_source": peaks["source"] + " bf16 sustained",
                    "algorithmic_flops_per_launch": ks["flops_per_launch"],
                    "mma_issued_tflops": ach * issue_factor if issue_factor else None,
                    "mma_issued_frac": ach * issue_factor / peaks["tflops"] if issue_factor else None,
                    "note": "achieved/frac count ALGORITHMIC conv FLOPs once; the kernel issues (k+3)/k x that in MMAs (row-group zero padding) and 3x in tc3 (hi/lo operand split)"}

    # ---------------- accuracy + residual of what was timed (outside the timed region) ----------------
    acc = None
    cpu_baseline = None
    if rank == 0:
        from oracle import poisson_oracle as O
        from poisson_cnn_b200.losses import linear_operator_loss
        nchk = min(B, args.check_samples)
        ref = O.pcnn_forward(hp_cfg, db_cfg, w, *[host[k][:nchk].double() for k in KEYS])
        got = out[:nchk].double().cpu()
        gs = torch.cat([dev_in[5], dev_in[5]], 1)
        res = float(linear_operator_loss(3, 2, ndims=2)(dev_in[0], out, gs))
        acc = {"rel_l2_vs_oracle_f64": float((got - ref).norm() / ref.norm()), "samples_checked": nchk, "laplacian_residual_mse": res}
        if world == 1 and not args.no_cpu_baseline:
            val, n, threads = oracle_throughput(hp_cfg, db_cfg, w, nx, ny, args.cpu_budget_s)
            cpu_baseline = {"value": val, "unit": "solutions/s", "cores": threads, "kind": "port",
                            "sample": "%d single-sample %dx%d Poisson_CNN forwards, torch-CPU fp32 oracle (TensorFlow reference not installable offline)" % (n, nx, ny)}

    # ---------------- the other precision modes, briefly (N=1 only; informational) ----------------
    other = {}
    if world == 1 and args.other_modes:
        for mode in [m for m in args.other_modes.split(",") if m and m != args.precision]:
            model.set_precision(mode)
            nst = 1 if mode == "fp32" else 2
            o = model(dev_in)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(nst):
                o = model(dev_in)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / nst
            nchk = min(B, args.check_samples)
            err = float((o[:nchk].double().cpu() - ref).norm() / ref.norm())
            other[mode] = {"value": B / (ms / 1000.0), "ms_per_step": ms, "rel_l2_vs_oracle_f64": err, "steps": nst}
        model.set_precision(args.precision)

    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "solutions/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp32": "f32", "tc": "f16", "tc3": "f16 hi+lo split (3 MMAs), f32 accumulate", "tc2": "f16 + e4m3 correction MMA (K=32), f32 accumulate",
                                     "mixed": "f16 + e4m3 correction MMA (K=32) in the HPNN, single-pass f16 in the DBCNN, f32 accumulate"}.get(args.precision, args.precision), "data": "synthetic",
            "config": {"workload": "Poisson_CNN_Legacy forward (HPNN + 4x DBCNN merged), batch %d per GPU, %dx%d grids, pcnn_end_to_end architecture, precision mode %s" % (B, nx, ny, args.precision),
                       "per_gpu_batch": B, "global_batch": B * world, "grid": [nx, ny], "parallelism": "batch-sharded x%d" % world,
                       "l2": "inputs+activations per step (%.1f GB) exceed the 126 MB L2" % (B * nx * ny * 4 * 32 / 1e9),
                       "flop_per_solution": pcnn_flops(nx, ny)},
            "e2e": {"value": e2e_value, "unit": "solutions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": out_host.numel() * 4, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "host_enqueue_ms_per_step": host_enqueue_ms,
            "clocks": sampler.result(),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "accuracy": acc,
            "other_modes": other,
            "model_tflops": value / world * pcnn_flops(nx, ny) / 1e12,
            "frac_of_bf16_sustained_peak": value / world * pcnn_flops(nx, ny) / 1e12 / peaks["tflops"],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="samples per GPU per step (BASELINE configs[1]: 256)")
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--precision", default=os.environ.get("PCNN_PRECISION", "mixed"), choices=["fp32", "tc", "tc2", "tc3", "mixed"],
                    help="mixed (default): tc2 in the HPNN + single-pass tc in the DBCNN, holds the 2e-3 budget with a 6x margin; tc2: fp16 main MMA + one e4m3 correction MMA everywhere; tc3: hi/lo fp16 split; tc: single FP16 pass; fp32: strict CUDA-core path")
    ap.add_argument("--other-modes", default="tc2,tc,tc3,fp32", help="comma list of extra precision modes timed briefly at N=1 (reported under other_modes)")
    ap.add_argument("--check-samples", type=int, default=2)
    ap.add_argument("--cpu-budget-s", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--roofline-kernel", type=lambda s: tuple(int(v) for v in s.split(",")), default=(32, 32, 15),
                    help="Cin,Cout,k of the conv whose launches are timed for the roofline (dominant: 32->32 k15)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
