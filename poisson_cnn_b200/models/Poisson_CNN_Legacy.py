"""Poisson_CNN_Legacy: homogeneous-RHS network + Dirichlet-BC network on all four edges, merged.

Mirrors poisson_CNN/models/Poisson_CNN_Legacy.py:5-51: model([rhs, left, top, right, bottom, dx])
with rhs [B,1,nx,ny], left/right [B,1,ny], top/bottom [B,1,nx], dx [B,1] -> [B,1,nx,ny].
The DBCNN weights are shared by the four boundaries, so the four calls are batched (4B when
nx == ny, otherwise 2B + 2B); the rot90/flip of flip_and_rotate_tensor and the 1/scaling-factor
rescales are folded into the final merge kernel.

Like the reference, every input is divided by its per-sample max|.| (set_max_magnitude_in_batch_and_return_scaling_factors,
Poisson_CNN_Legacy.py:23-28): an all-zero right-hand side or boundary (a homogeneous problem) gives 1/0 = inf and a NaN
prediction for that sample -- the reference's behaviour, kept deliberately; pass the homogeneous part to the HPNN alone.
"""
import torch

from .. import ops
from ._base import WeightedModel


class Poisson_CNN_Legacy(WeightedModel):
    COMPLIANT_PRECISIONS = ("fp32", "tc2", "tc3", "mixed")     # 'tc': 4.4e-3 at 256x256, outside the 2e-3 budget

    def __init__(self, hpnn, dbcnn, jacobi_iterations=0, max_microbatch=128):
        super().__init__()
        self.hpnn = hpnn
        self.dbcnn = dbcnn
        self.data_format = hpnn.data_format
        self.jacobi_iterations = jacobi_iterations
        # samples are independent end to end, so a large batch is processed in slices: bounds activation memory
        # (the DBCNN sees 4x the slice) without changing any result.  The bound is stated in 256x256-grid samples
        # (128 -> ~63 GB of pooled activations in 'mixed' mode) and scales inversely with the grid size.
        self.max_microbatch = max_microbatch
        self.microbatch_samples = None
        self._io_streams = None     # (host->device, device->host) copy streams of the host-buffer call path

    def weight_specs(self, prefix=""):
        hs, hm = self.hpnn.weight_specs(prefix + "hpnn/")
        ds, dm = self.dbcnn.weight_specs(prefix + "dbcnn/")
        return {**hs, **ds}, {**hm, **dm}

    def keras_key_map(self, prefix=""):
        from .. import tf_checkpoint as T
        return T.pcnn_key_map(self.hpnn._cfg, self.dbcnn._cfg, prefix)

    def load_weights(self, source, prefix="", device=None):
        if isinstance(source, str):
            source = self._read_weight_file(source, prefix)     # TF checkpoint prefix (reference format) or .npz
        self.hpnn.load_weights(source, prefix + "hpnn/", device)
        self.dbcnn.load_weights(source, prefix + "dbcnn/", device)
        self.device = self.hpnn.device
        self._on_weights_loaded()
        return self

    def get_weights_dict(self, prefix=""):
        return {**self.hpnn.get_weights_dict(prefix + "hpnn/"), **self.dbcnn.get_weights_dict(prefix + "dbcnn/")}

    def __call__(self, inp, out=None, non_blocking=False):
        """inp = [rhs, left, top, right, bottom, dx].  CUDA tensors -> CUDA result on the current stream (no sync).
        HOST tensors (what a Keras user passes) -> the batch is streamed through the device in slices, the
        host->device copy of slice i+1 and the device->host copy of slice i-1 overlapping the kernels of slice i on two
        copy streams; returns a pinned host tensor (`out` if given: pass a pinned [B,1,nx,ny] float32 buffer to avoid the
        page-locking cost per call).  The host result is COMPLETE on return, like the reference's model(x).numpy();
        non_blocking=True returns (out, event) instead, without waiting: `out` may be read after event.synchronize()."""
        rhs, left, top, right, bottom, dx = inp
        if rhs.dim() != 4 or rhs.shape[1] != 1:
            raise ValueError("rhs must be [batch, 1, nx, ny] (channels_first)")
        B, _, nx, ny = rhs.shape
        for t, n, name in ((left, ny, "left"), (right, ny, "right"), (top, nx, "top"), (bottom, nx, "bottom")):
            if tuple(t.shape) != (B, 1, n):
                raise ValueError("%s boundary must be [batch, 1, %d], got %s" % (name, n, tuple(t.shape)))
        mb = self.microbatch_samples            # explicit slice size, if set
        if mb is None and self.max_microbatch:
            mb = max(1, int(self.max_microbatch * 65536 // (nx * ny)))
        if mb:
            mb = min(mb, 65535 // 4)            # the DBCNN runs 4x the slice; several kernels put the batch in gridDim.y/z
        if not rhs.is_cuda:
            return self._run_host(list(inp), mb, out, non_blocking)
        if out is not None or non_blocking:
            raise ValueError("out= / non_blocking= apply to host inputs only")
        with torch.cuda.device(rhs.device):     # launches go to the current device: make it the tensors' device
            try:
                return self._run(rhs, left, top, right, bottom, dx, mb)
            except torch.OutOfMemoryError:
                # the activation-buffer pool keeps one set of buffers per tensor shape ever seen in this process; when a
                # new grid shape does not fit next to the idle ones, they are released and the call is repeated once
                from .. import ops
                ops.blk8_pool_clear()
                torch.cuda.empty_cache()
                return self._run(rhs, left, top, right, bottom, dx, mb)

    def capture(self, example_inputs):
        """CUDA-graph the forward pass for the (batch, grid) shape of `example_inputs` (CUDA tensors, batch no larger
        than one micro-batch): returns a callable with the same [rhs, left, top, right, bottom, dx] signature whose
        cost on the host is one graph launch instead of ~165 kernel launches.  See poisson_cnn_b200/graph.py."""
        from ..graph import GraphedCall
        rhs = example_inputs[0]
        B, _, nx, ny = rhs.shape
        mb = self.microbatch_samples
        if mb is None and self.max_microbatch:
            mb = max(1, int(self.max_microbatch * 65536 // (nx * ny)))
        if mb and B > mb:
            raise ValueError("capture: batch %d exceeds one micro-batch (%d); graphs are for small batches" % (B, mb))
        return GraphedCall(self._forward, list(example_inputs))

    def _run_host(self, host, mb, out, non_blocking=False):
        if self.device is None:
            raise ValueError("load_weights() first: the model does not know its device yet")
        dev = self.device
        if any(t.is_cuda for t in host):
            raise ValueError("inputs must be all on the host or all on the device")
        host = [t.contiguous().float() for t in host]
        host = [t if t.is_pinned() else t.pin_memory() for t in host]
        B, _, nx, ny = host[0].shape
        if out is None:
            out = torch.empty((B, 1, nx, ny), dtype=torch.float32).pin_memory()
        elif tuple(out.shape) != (B, 1, nx, ny) or out.dtype != torch.float32 or out.is_cuda or not out.is_contiguous():
            raise ValueError("out must be a contiguous host float32 tensor of shape [batch, 1, nx, ny]")
        mb = mb or B
        # only the first slice's host->device copy and the last slice's device->host copy are exposed
        with torch.cuda.device(dev):
            if self._io_streams is None:
                self._io_streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
            s_in, s_out = self._io_streams
            cur = torch.cuda.current_stream(dev)
            slices = [(lo, min(lo + mb, B)) for lo in range(0, B, mb)]
            bufs = [[torch.empty((hi - lo,) + tuple(t.shape[1:]), device=dev, dtype=torch.float32) for t in host] for lo, hi in slices]
            alloc = torch.cuda.Event(); alloc.record(cur)
            s_in.wait_event(alloc)              # recycled blocks may still be read by kernels queued earlier
            ready = []
            with torch.cuda.stream(s_in):
                for (lo, hi), bs in zip(slices, bufs):
                    for b, t in zip(bs, host):
                        b.copy_(t[lo:hi], non_blocking=True)
                    ev = torch.cuda.Event(); ev.record(s_in); ready.append(ev)
            for (lo, hi), bs, ev in zip(slices, bufs, ready):
                cur.wait_event(ev)
                try:
                    o = self._forward(bs)
                except torch.OutOfMemoryError:
                    ops.blk8_pool_clear()
                    torch.cuda.empty_cache()
                    o = self._forward(bs)
                done = torch.cuda.Event(); done.record(cur)
                s_out.wait_event(done)
                with torch.cuda.stream(s_out):
                    out[lo:hi].copy_(o, non_blocking=True)
                o.record_stream(s_out)
            fin = torch.cuda.Event(); fin.record(s_out)
            cur.wait_event(fin)                 # the caller's stream order covers the last device->host copy
        if non_blocking:
            return out, fin
        fin.synchronize()                       # stream order does not order the HOST: wait for the last copy
        return out

    def _engine_config(self):
        return {"hpnn_model": self.hpnn._cfg, "dbcnn_model": self.dbcnn._cfg, "jacobi_iterations": int(self.jacobi_iterations)}

    def _engine_weights(self):
        return self.get_weights_dict()

    def _engine_microbatch(self, nx, ny):
        """Explicit slice size for the engine, or 0 for its automatic rule (128 * 65536 / (nx*ny): this class's default)."""
        if self.microbatch_samples:
            return int(self.microbatch_samples)
        if self.max_microbatch and self.max_microbatch != 128:
            return max(1, int(self.max_microbatch * 65536 // (nx * ny)))
        return 0 if self.max_microbatch else 65535 // 4

    def _run(self, rhs, left, top, right, bottom, dx, mb):
        B, _, nx, ny = rhs.shape
        if self.use_engine:                     # the engine slices the batch itself, inside one workspace
            return self._forward([rhs, left, top, right, bottom, dx])
        if mb and B > mb:
            out = torch.empty((B, 1, nx, ny), device=rhs.device, dtype=torch.float32)
            for lo in range(0, B, mb):
                out[lo:lo + mb] = self._forward([t[lo:lo + mb] for t in (rhs, left, top, right, bottom, dx)])
            return out
        return self._forward([rhs, left, top, right, bottom, dx])

    def _forward(self, inp):
        rhs, left, top, right, bottom, dx = inp
        B, _, nx, ny = rhs.shape
        if self.use_engine:
            e = self.engine()
            e.set_microbatch(self._engine_microbatch(nx, ny))
            return e.forward(rhs, left, top, right, bottom, dx)

        # per-sample max-normalisation of the five inputs (set_max_magnitude_in_batch_and_return_scaling_factors)
        mrhs = ops.maxabs(rhs)
        rhs_n = ops.scale_inv(rhs, mrhs)
        ml, mt, mr, mb = ops.maxabs(left), ops.maxabs(top), ops.maxabs(right), ops.maxabs(bottom)

        hp = self.hpnn([rhs_n, dx])

        if nx == ny:
            bcs = torch.empty((4 * B, 1, ny), device=rhs.device, dtype=torch.float32)
            for i, (t, m) in enumerate(((left, ml), (top, mt), (right, mr), (bottom, mb))):
                ops.scale_inv(t, m, out=bcs[i * B:(i + 1) * B])
            res = self.dbcnn([bcs, dx.repeat(4, 1), nx])
            L, T, R, Bt = (res[i * B:(i + 1) * B] for i in range(4))
        else:
            lr = torch.empty((2 * B, 1, ny), device=rhs.device, dtype=torch.float32)
            ops.scale_inv(left, ml, out=lr[:B]); ops.scale_inv(right, mr, out=lr[B:])
            tb = torch.empty((2 * B, 1, nx), device=rhs.device, dtype=torch.float32)
            ops.scale_inv(top, mt, out=tb[:B]); ops.scale_inv(bottom, mb, out=tb[B:])
            dx2 = dx.repeat(2, 1)
            res_lr = self.dbcnn([lr, dx2, nx])
            res_tb = self.dbcnn([tb, dx2, ny])
            L, R = res_lr[:B], res_lr[B:]
            T, Bt = res_tb[:B], res_tb[B:]

        pred = ops.merge(hp, L, T, R, Bt, dx, mrhs, ml, mt, mr, mb)
        if self.jacobi_iterations > 0:
            # the reference passes the max-normalised rhs here (it rebinds `rhs`, Poisson_CNN_Legacy.py:23,49)
            pred = ops.jacobi(pred, rhs_n, torch.cat([dx, dx], 1), self.jacobi_iterations)
        return pred

    call = __call__
