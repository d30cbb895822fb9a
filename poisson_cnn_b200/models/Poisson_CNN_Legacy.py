"""Poisson_CNN_Legacy: homogeneous-RHS network + Dirichlet-BC network on all four edges, merged.

Mirrors poisson_CNN/models/Poisson_CNN_Legacy.py:5-51: model([rhs, left, top, right, bottom, dx])
with rhs [B,1,nx,ny], left/right [B,1,ny], top/bottom [B,1,nx], dx [B,1] -> [B,1,nx,ny].
The DBCNN weights are shared by the four boundaries, so the four calls are batched (4B when
nx == ny, otherwise 2B + 2B); the rot90/flip of flip_and_rotate_tensor and the 1/scaling-factor
rescales are folded into the final merge kernel.
"""
import torch

from .. import ops
from ._base import WeightedModel


class Poisson_CNN_Legacy(WeightedModel):
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
        return self

    def get_weights_dict(self, prefix=""):
        return {**self.hpnn.get_weights_dict(prefix + "hpnn/"), **self.dbcnn.get_weights_dict(prefix + "dbcnn/")}

    def __call__(self, inp):
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
        try:
            return self._run(rhs, left, top, right, bottom, dx, mb)
        except torch.OutOfMemoryError:
            # the activation-buffer pool keeps one set of buffers per tensor shape ever seen in this process; when a
            # new grid shape does not fit next to the idle ones, they are released and the call is repeated once
            from .. import ops
            ops.blk8_pool_clear()
            torch.cuda.empty_cache()
            return self._run(rhs, left, top, right, bottom, dx, mb)

    def _run(self, rhs, left, top, right, bottom, dx, mb):
        B, _, nx, ny = rhs.shape
        if mb and B > mb:
            out = torch.empty((B, 1, nx, ny), device=rhs.device, dtype=torch.float32)
            for lo in range(0, B, mb):
                out[lo:lo + mb] = self._forward([t[lo:lo + mb] for t in (rhs, left, top, right, bottom, dx)])
            return out
        return self._forward([rhs, left, top, right, bottom, dx])

    def _forward(self, inp):
        rhs, left, top, right, bottom, dx = inp
        B, _, nx, ny = rhs.shape

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
