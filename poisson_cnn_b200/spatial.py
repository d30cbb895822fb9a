"""Single-grid spatial decomposition of the HPNN across GPUs (first half).

The use case is the reference's Navier-Stokes pressure projection (Navier_Stokes_2D/solvers.py:29-33,204-334): ONE large grid
per time step, batch 1 -- batch sharding has nothing to split.  The grid is cut into P horizontal bands of H/P rows, one per
GPU (torch.distributed over NCCL), and every full-resolution convolution of the trunk runs on its band:

  * The BLK8 layout materialises a 7-row halo above and below every tensor and the tcgen05 convolution kernel is
    padding-agnostic ("the padding mode is whatever the halo holds").  At a band edge that is not a physical boundary the
    halo rows are simply filled with the neighbour's edge rows -- the kernels run unchanged and compute, bit for bit, what
    the monolithic convolution computes for those rows.  One exchange of <= 7 rows x 2 neighbours per layer (peer copies of
    (W+14) x 16 B x planes per row; 1.8 MB per direction at 2048^2 in `tc2`).
  * Global couplings: the pooled pyramids of the bottleneck branches need the whole map -> the band features are
    all-gathered once (32 channels) and the low-resolution branches (9 % of the FLOPs) plus the fused upsample-merge run
    REPLICATED on every rank; each rank keeps its band of the merged map.  The Scaling block and the boundary ring need the
    whole 1-channel output -> all-gather, replicated tail.  The dx MLP is replicated.  No other exchange.

The result is bit-identical to the single-GPU tensor-core program (tests/test_gpu_spatial.py), because every kernel sees
exactly the operands it sees there.  Expected speed-up = 1 / (0.86 / P + 0.14) minus the exchanges (the replicated branches
and merge are ~14 % of a forward at 2048^2).

`world=P, comm=None` emulates the P bands inside ONE process on one GPU (the exchange is a device copy): the test of the
decomposition logic; `comm=torch.distributed group` runs one band per rank.

Not built: the DBCNN / merged-model spatial split (the top/bottom boundary networks need the transposed, column-band
exchange); the HPNN is what the pressure-projection caller uses.
"""
import torch

from . import ops
from .config import PAD_CONSTANT, PAD_SYMMETRIC, ACT_LEAKY_RELU, ACT_LINEAR

HALO = 7


def _view(t, which="buf"):
    """[B, planes, H+14, W+14, 8] view of a Blk8 buffer (fp16 elements; a q buffer has the same byte geometry)."""
    buf = getattr(t, which)
    planes = (t.C + 15) // 16 * 2
    n = t.B * planes * (t.H + 2 * HALO) * (t.W + 2 * HALO) * 8
    return buf[:n].view(t.B, planes, t.H + 2 * HALO, t.W + 2 * HALO, 8)


class SpatialHPNN:
    """model([rhs, dx]) of a Homogeneous_Poisson_NN_Legacy with the grid split into row bands.

    rhs [B,1,H,W] and dx [B,1] are given in full on every rank and the full result is returned on every rank; H must be a
    multiple of the number of bands and every band at least 16 rows.  Tensor-core precisions only ('mixed', 'tc2', 'tc3',
    'tc'): the strict FP32 kernels resolve padding while loading tiles and have no materialised halo to exchange."""

    def __init__(self, hpnn, world=None, group=None):
        import torch.distributed as dist
        self.m = hpnn
        self.group = group
        self.dist = dist if (group is not None or (world is None and dist.is_available() and dist.is_initialized())) else None
        if self.dist is not None:
            self.world = dist.get_world_size(group)
            self.rank = dist.get_rank(group)
            self.local = [self.rank]
        else:
            self.world = int(world or 1)
            self.rank = 0
            self.local = list(range(self.world))       # emulation: every band lives in this process
        self.pools = {i: {} for i in self.local}        # private BLK8 pools: band buffers carry neighbour rows in their halos
        self.full_pool = {}
        self.profile = None                             # set to {} to collect CUDA-event times per phase (ms, last call)

    def _tick(self, name):
        if self.profile is None:
            return
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self._marks.append((name, e))

    def _report(self):
        if self.profile is None:
            return
        torch.cuda.synchronize()
        self.profile.clear()
        for (n0, e0), (n1, e1) in zip(self._marks[:-1], self._marks[1:]):
            self.profile[n1] = self.profile.get(n1, 0.0) + e0.elapsed_time(e1)

    # ------------------------------------------------------------------ exchange
    def _exchange(self, ts):
        """Fill the halo rows at interior band edges with the neighbour's edge rows.  ts: {band: Blk8}."""
        for which in ("buf", "lo"):
            if any(getattr(t, which) is None for t in ts.values()):
                continue
            if self.dist is None:
                for i in self.local[:-1]:
                    up, dn = _view(ts[i], which), _view(ts[i + 1], which)
                    Hp = up.shape[2]
                    dn[:, :, 0:HALO].copy_(up[:, :, Hp - 2 * HALO:Hp - HALO])        # my bottom rows -> lower band's top halo
                    up[:, :, Hp - HALO:Hp].copy_(dn[:, :, HALO:2 * HALO])            # lower band's top rows -> my bottom halo
            else:
                v = _view(ts[self.rank], which)
                Hp = v.shape[2]
                reqs, recv = [], {}
                if self.rank > 0:
                    send_up = v[:, :, HALO:2 * HALO].contiguous()
                    recv["top"] = torch.empty_like(send_up)
                    reqs += [self.dist.P2POp(self.dist.isend, send_up, self._peer(self.rank - 1), self.group),
                             self.dist.P2POp(self.dist.irecv, recv["top"], self._peer(self.rank - 1), self.group)]
                if self.rank < self.world - 1:
                    send_dn = v[:, :, Hp - 2 * HALO:Hp - HALO].contiguous()
                    recv["bottom"] = torch.empty_like(send_dn)
                    reqs += [self.dist.P2POp(self.dist.isend, send_dn, self._peer(self.rank + 1), self.group),
                             self.dist.P2POp(self.dist.irecv, recv["bottom"], self._peer(self.rank + 1), self.group)]
                if reqs:
                    for r in self.dist.batch_isend_irecv(reqs):
                        r.wait()
                if "top" in recv:
                    v[:, :, 0:HALO].copy_(recv["top"])
                if "bottom" in recv:
                    v[:, :, Hp - HALO:Hp].copy_(recv["bottom"])
        for t in ts.values():
            t.halo = (t.halo[0], HALO)        # the halo now holds what the next layer must see: no refill

    def _peer(self, r):
        return r if self.group is None else self.dist.get_global_rank(self.group, r)

    def _gather_rows(self, parts):
        """{band: [B,C,h,W] fp32} -> [B,C,H,W] on every rank."""
        if self.dist is None:
            return torch.cat([parts[i] for i in self.local], 2)
        x = parts[self.rank].contiguous()
        out = torch.empty((self.world,) + tuple(x.shape), device=x.device, dtype=x.dtype)
        self.dist.all_gather_into_tensor(out, x, group=self.group)
        return out.permute(1, 2, 0, 3, 4).reshape(x.shape[0], x.shape[1], self.world * x.shape[2], x.shape[3]).contiguous()

    # ------------------------------------------------------------------ band-wise layers
    def _each(self, fn):
        out = {}
        for i in self.local:
            with ops.blk8_pool_scope(self.pools[i]):
                out[i] = fn(i)
        return out

    def _conv(self, xs, name, act, pad, bn_name=None, residual=None, out_scale=None, out=None, next_pad=PAD_CONSTANT):
        m = self.m
        ys = self._each(lambda i: m._conv_tc(xs[i], name, act, pad, bn_name, None if residual is None else residual[i],
                                             out_scale, None if out is None else out[i], 0, next_pad))
        if out is None:          # a conv writing part of a wider tensor is exchanged by the caller once the tensor is complete
            self._exchange(ys)
        return ys

    def _resnet(self, xs, name, act, pad, use_bn, out_scale=None, next_pad=PAD_CONSTANT):
        t = self._conv(xs, name + "/conv0", act, pad, name + "/bn0" if use_bn else None, next_pad=pad)
        t = self._conv(t, name + "/conv1", act, pad, name + "/bn1" if use_bn else None, residual=xs, next_pad=pad)
        return self._conv(t, name + "/conv2", act, pad, out_scale=out_scale, next_pad=next_pad)

    # ------------------------------------------------------------------ forward
    def __call__(self, inp):
        rhs, dx = inp
        m = self.m
        if m.precision not in ("tc", "tc2", "tc3"):
            raise ValueError("the spatial decomposition runs the tensor-core program: set_precision('mixed' | 'tc2' | 'tc3')")
        if not m._tc_supported():
            raise NotImplementedError("config not supported by the tensor-core program")
        B, _, H, Wd = rhs.shape
        P = self.world
        if H % P or H // P < 16:
            raise ValueError("the grid height (%d) must be a multiple of the %d bands and every band at least 16 rows" % (H, P))
        h = H // P
        F = m.filters
        dev = rhs.device
        split = m.tc_split
        rhs = rhs.contiguous()
        with torch.cuda.device(dev):
            self._marks = []
            self._tick("start")
            posx, posy = ops.position_table(dev, H), ops.position_table(dev, Wd)

            def first(i):
                band = rhs[:, :, i * h:(i + 1) * h].contiguous()
                if m.use_positional_embeddings:
                    x = torch.empty((B, 3, h, Wd), device=dev, dtype=torch.float32)
                    ops.check(ops.lib.pcnn_hpnn_input_f32(band.data_ptr(), posx[i * h:].data_ptr(), posy.data_ptr(), x.data_ptr(), B, h, Wd,
                                                          ops._stream()), "hpnn_input")
                else:
                    x = band
                return ops.to_blk8(x, split=split, halo=m.pre_pad)
            t = self._each(first)
            self._exchange(t)
            for k in range(m.n_pre):
                t = self._conv(t, "pre_bottleneck/%d" % k, m.pre_act, m.pre_pad, "pre_bottleneck/%d/bn" % k if m.use_batchnorm else None,
                               next_pad=m.pre_pad if k + 1 < m.n_pre else PAD_CONSTANT)
            x0 = t
            self._tick("pre_bottleneck (banded)")
            # the bottleneck branches need the whole map: gather the band features, run the low-resolution branches and the
            # fused upsample-merge replicated, keep this rank's rows of the merged map
            x0_full = self._gather_rows(self._each(lambda i: ops.from_blk8(x0[i])))
            self._tick("gather x0")
            with ops.blk8_pool_scope(self.full_pool):
                branches = m._branches_tc(x0_full, H, Wd, split)
                self._tick("branches (replicated)")
                cat_full = ops.Blk8(B, 2 * F, H, Wd, dev, split=split)
                m._merge_tc(branches, cat_full, B, H, Wd, dev)
                self._tick("upsample-merge (replicated)")
            cat = self._each(lambda i: ops.Blk8(B, 2 * F, h, Wd, dev, split=split))
            self._conv(x0, "non_bottleneck_conv", ACT_LEAKY_RELU, PAD_CONSTANT, out=cat)
            p0 = (F // 16) * 2                       # first plane of channels [F, 2F)
            for which in ("buf", "lo"):
                if getattr(cat_full, which) is None:
                    continue
                vf = _view(cat_full, which)
                for i in self.local:
                    _view(cat[i], which)[:, p0:, HALO:HALO + h].copy_(vf[:, p0:, HALO + i * h:HALO + (i + 1) * h])
            self._exchange(cat)
            del cat_full, branches
            self._tick("non_bottleneck_conv + band copy")
            y = self._conv(cat, "post_merge_conv", ACT_LEAKY_RELU, PAD_CONSTANT)
            d = ops.dense_input(dx, H, Wd)
            d = ops.dense(d, *m.conv("dx_dense/0"), ACT_LEAKY_RELU)
            d = ops.dense(d, *m.conv("dx_dense/1"), ACT_LEAKY_RELU)
            d = ops.dense(d, *m.conv("dx_dense/2"), ACT_LINEAR)
            y = self._resnet(y, "post_merge_resnet", ACT_LEAKY_RELU, PAD_CONSTANT, False, out_scale=d)
            S, nreg = m.n_final, m.final_regular_conv_stages
            for k in range(S - nreg):
                y = self._conv(y, "final/%d/conv" % k, m.final_act, m.final_pad)
                y = self._resnet(y, "final/%d/resnet" % k, m.final_act, PAD_CONSTANT, False)
            for k in range(S - nreg, S):
                y = self._conv(y, "final/%d/conv" % k, ACT_LINEAR, PAD_CONSTANT)
            self._tick("trunk after the merge (banded)")
            # Scaling, boundary ring, post-smoother need the whole (single-channel) map: gather, replicated tail
            y_full = self._gather_rows(self._each(lambda i: ops.from_blk8(y[i], C=y[i].C)))
            out = m._tail(y_full, rhs, dx, S, None)
            self._tick("gather + tail (replicated)")
            self._report()
            return out
