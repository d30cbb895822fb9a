"""Single-grid spatial decomposition of the HPNN across GPUs (first half).

The use case is the reference's Navier-Stokes pressure projection (Navier_Stokes_2D/solvers.py:29-33,204-334): ONE large grid
per time step, batch 1 -- batch sharding has nothing to split.  The grid is cut into P horizontal bands of H/P rows, one per
GPU (torch.distributed over NCCL), and every full-resolution convolution of the trunk runs on its band:

  * The BLK8 layout materialises a 7-row halo above and below every tensor and the tcgen05 convolution kernel is
    padding-agnostic ("the padding mode is whatever the halo holds").  At a band edge that is not a physical boundary the
    halo rows are simply filled with the neighbour's edge rows -- the kernels run unchanged and compute, bit for bit, what
    the monolithic convolution computes for those rows.  One exchange of <= 7 rows x 2 neighbours per layer (peer copies of
    (W+14) x 16 B x planes per row; 1.8 MB per direction at 2048^2 in `tc2`).
  * Bottleneck branches whose pooling windows and transpose-conv phases align with the bands (down-sampling factor divides
    the band height, >= 16 low-resolution rows per band: ds = 2, 4, 8, 16 at 2048^2 on 8 GPUs) are split the same way: pooled
    per band, convolved per band with their own halo exchange, up-sampled into the band of the merged map by the fused
    merge kernel with band-local inputs.  The others (ds = 3: 2048 is not a multiple of 3; the 2x2 .. 64x64 maps of the
    multilinear branches) run REPLICATED from one all-gather of the 32-channel features; the merge takes their rows /
    the full-map interpolation tables entered at the band's first row.
  * The Scaling block and the boundary ring need the whole 1-channel output -> all-gather, replicated tail.  The dx MLP is
    replicated.  No other exchange.

The result is bit-identical to the single-GPU tensor-core program (tests/test_gpu_spatial.py), because every kernel sees
exactly the operands it sees there (including the hierarchical pooling chain, which follows the full map's divisibility).

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


def band_bounds(H, P, strides):
    """Row boundaries [r_0 = 0, ..., r_P = H] of P bands: as even as possible on multiples of lcm(strides), so that the
    low-resolution rows and the up-sampling phases of every transpose-conv branch split with the bands; a plain equal
    split when the grid is too small for aligned bands of >= 16 rows."""
    import math
    L = 1
    for s_ in strides:
        L = L * int(s_) // math.gcd(L, int(s_))
    bnd = [0] + [int(round(i * H / P / L)) * L for i in range(1, P)] + [H]
    if any(bnd[i + 1] - bnd[i] < 16 for i in range(P)):
        bnd = [(i * H) // P for i in range(P + 1)]
    if any(bnd[i + 1] - bnd[i] < 16 for i in range(P)):
        raise ValueError("every band of the grid height (%d) over %d bands must have at least 16 rows" % (H, P))
    return bnd


class SpatialHPNN:
    """model([rhs, dx]) of a Homogeneous_Poisson_NN_Legacy with the grid split into row bands.

    rhs [B,1,H,W] and dx [B,1] are given in full on every rank and the full result is returned on every rank; every band has
    at least 16 rows.  Tensor-core precisions only ('mixed', 'tc2', 'tc3',
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
        self.min_band_pixels = 128 * 1024               # low-resolution pixels per band below which a branch runs replicated
        self._stage = {}                                # send / receive staging buffers of the halo exchange

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
        whiches = [w for w in ("buf", "lo") if all(getattr(t, w) is not None for t in ts.values())]
        if self.dist is None:
            for which in whiches:
                for i in self.local[:-1]:
                    up, dn = _view(ts[i], which), _view(ts[i + 1], which)
                    Hp = up.shape[2]
                    dn[:, :, 0:HALO].copy_(up[:, :, Hp - 2 * HALO:Hp - HALO])        # my bottom rows -> lower band's top halo
                    up[:, :, Hp - HALO:Hp].copy_(dn[:, :, HALO:2 * HALO])            # lower band's top rows -> my bottom halo
        else:
            reqs, fills = [], []
            for which in whiches:                      # one NCCL group for both buffers and both neighbours
                v = _view(ts[self.rank], which)
                Hp = v.shape[2]
                for side, peer, rows_out, rows_in in (("up", self.rank - 1, slice(HALO, 2 * HALO), slice(0, HALO)),
                                                      ("dn", self.rank + 1, slice(Hp - 2 * HALO, Hp - HALO), slice(Hp - HALO, Hp))):
                    if peer < 0 or peer >= self.world:
                        continue
                    # staging buffers are kept per (buffer, side, geometry): no allocation in the steady state
                    key = (which, side, tuple(v.shape), v.dtype)
                    if key not in self._stage:
                        shp = (v.shape[0], v.shape[1], HALO, v.shape[3], v.shape[4])
                        self._stage[key] = (torch.empty(shp, device=v.device, dtype=v.dtype), torch.empty(shp, device=v.device, dtype=v.dtype))
                    send, got = self._stage[key]
                    send.copy_(v[:, :, rows_out])
                    reqs += [self.dist.P2POp(self.dist.isend, send, self._peer(peer), self.group),
                             self.dist.P2POp(self.dist.irecv, got, self._peer(peer), self.group)]
                    fills.append((v[:, :, rows_in], got))
            if reqs:
                for r in self.dist.batch_isend_irecv(reqs):
                    r.wait()
            for dst, got in fills:
                dst.copy_(got)
        for t in ts.values():
            t.halo = (t.halo[0], HALO)        # the halo now holds what the next layer must see: no refill

    def _peer(self, r):
        return r if self.group is None else self.dist.get_global_rank(self.group, r)

    def _gather_rows(self, parts, div=1):
        """{band: [B,C,h_i/div,W'] fp32} -> the full map on every rank (bands may differ in height; div: the pooling level
        of the parts, whose band i holds ceil(h_i / div) rows)."""
        if self.dist is None:
            return torch.cat([parts[i] for i in self.local], 2)
        x = parts[self.rank]
        rows = [-(-(self.bounds[i + 1] - self.bounds[i]) // div) for i in range(self.world)]
        assert x.shape[2] == rows[self.rank]
        pad = torch.zeros((x.shape[0], x.shape[1], max(rows), x.shape[3]), device=x.device, dtype=x.dtype)
        pad[:, :, :x.shape[2]].copy_(x)
        out = torch.empty((self.world,) + tuple(pad.shape), device=x.device, dtype=x.dtype)
        if x.is_cuda:
            self.dist.all_gather_into_tensor(out, pad, group=self.group)
        else:                      # gloo (the CPU test of this logic) takes the per-rank list form
            self.dist.all_gather(list(out.unbind(0)), pad, group=self.group)
        return torch.cat([out[i, :, :, :rows[i]] for i in range(self.world)], 2)

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

    # ------------------------------------------------------------------ bottleneck branches
    def _pool_chain(self, H, Wd, factors):
        """{s: (source level t or 0 for the features, factor)} exactly as Homogeneous_Poisson_NN_Legacy._pool_pyramid decides
        for the FULL map (means of means differ from direct means in the last bit: the bands must follow the same chain)."""
        chain, done = {}, []
        for s in sorted(set(factors)):
            src, f = 0, s
            if H % s == 0 and Wd % s == 0:
                for t in sorted(done, reverse=True):
                    if s % t == 0 and H % t == 0 and Wd % t == 0:
                        src, f = t, s // t
                        break
            chain[s] = (src, f)
            done.append(s)
        return chain

    def _branches(self, x0, H, Wd, split):
        """Low-resolution branch outputs for the banded merge: {'dc': [(per-band Blk8 dict, kernel, bias, stride, act)],
        'rs': [(full fp32 source, method)], 'um': 1|2} or None when the banded merge does not apply (then everything runs
        replicated through the model's own _branches_tc / _merge_tc)."""
        m = self.m
        F = m.filters
        blocks = m.bottleneck_deconv_blocks + m.bottleneck_multilinear_blocks
        bsplit = 1 if m.requested_precision == "mixed" else split
        deconv = [b for b in blocks if b.kind == "deconv"]
        strides = [b.upsampling_factor for b in deconv]
        rs_hw = [(-(-H // b.downsampling_factor), -(-Wd // b.downsampling_factor)) for b in blocks if b.kind != "deconv"]
        bnd = self.bounds
        hs = [bnd[i + 1] - bnd[i] for i in range(self.world)]
        # band boundaries are multiples of every transpose-conv stride (self.align), so every branch's low-resolution rows and
        # up-sampling phases split cleanly; SAME pooling must not start before row 0 (pad_before = 0)
        ok = (bsplit == 1 and F == 32 and bool(deconv) and len(blocks) <= 16 and all(s <= 32 for s in strides) and len(deconv) <= 8
              and len(rs_hw) <= 8 and all(b.upsampling_factor == b.downsampling_factor and all(r % b.upsampling_factor == 0 for r in bnd[:-1])
                                          and (-(-H // b.downsampling_factor) * b.downsampling_factor - H) // 2 == 0 for b in deconv))
        if not ok:
            return None
        um = 1 if ops.upsample_merge_tc_fits(strides, rs_hw) else (2 if (rs_hw and ops.upsample_merge_tc_fits(strides, [])) else 0)
        if um == 0 or (um == 1 and any(F * a * b > 8192 for a, b in rs_hw)):
            return None
        for b in blocks:
            m._branch_out_hw(b, H, Wd)
        chain = self._pool_chain(H, Wd, [b.downsampling_factor for b in blocks])
        # a branch is split with the bands when a band of its low-resolution map is large enough for the convolution time
        # saved to exceed the ~80 us an exchange costs per layer (measured: splitting every branch at N = 2 took 5.6 ms where
        # the replicated branches take 3); smaller maps run replicated
        banded = {b.downsampling_factor for b in deconv
                  if min(hs) // b.downsampling_factor >= 16 and -(-Wd // b.downsampling_factor) >= 16
                  and (min(hs) // b.downsampling_factor) * (-(-Wd // b.downsampling_factor)) >= self.min_band_pixels}
        self._tick("branch setup")

        def aligned(t):       # level t can be pooled per band exactly as the full map pools it
            return all(r % t == 0 for r in bnd[:-1]) and (-(-H // t) * t - H) // 2 == 0 and min(hs) >= t
        rep = [b for b in blocks if b.downsampling_factor not in banded]
        # every replicated level is derived, along the full map's chain, from the SMALLEST map that can be pooled per band and
        # gathered (level 16 for the 32/64/128 branches at 2048^2: 2 MB instead of the 537 MB of the features themselves)
        gather_levels, need_x0 = set(), False
        for b in rep:
            t = b.downsampling_factor
            while t and not aligned(t):
                t = chain[t][0]
            if t:
                gather_levels.add(t)
            else:
                need_x0 = True
        needed = set(banded) | gather_levels      # per-band levels and, transitively, the levels they are pooled from
        grew = True
        while grew:
            grew = False
            for q in list(needed):
                if chain[q][0] and chain[q][0] not in needed:
                    needed.add(chain[q][0])
                    grew = True
        band_pool = {}
        x0f = self._each(lambda i: ops.from_blk8(x0[i])) if (needed or need_x0) else None
        for s_ in sorted(needed):
            src, f = chain[s_]
            band_pool[s_] = self._each(lambda i: ops.avgpool_same(x0f[i] if src == 0 else band_pool[src][i], f))
        full = {0: self._gather_rows(x0f)} if need_x0 else {}
        for t in sorted(gather_levels):
            full[t] = self._gather_rows(band_pool[t], div=t)
        rep_out = {}
        with ops.blk8_pool_scope(self.full_pool):
            def full_level(t):
                if t not in full:
                    src, f = chain[t]
                    full[t] = ops.avgpool_same(full_level(src), f)
                return full[t]
            for b in rep:
                ph, pw = -(-H // b.downsampling_factor), -(-Wd // b.downsampling_factor)
                name = "bottleneck_%s/%d" % (b.kind, b.index)
                pooled = full_level(b.downsampling_factor)
                if m._branch_on_tc(b, ph, pw):
                    t = ops.to_blk8(pooled, split=bsplit, halo=b.pad)
                    t = m._conv_tc(t, name + "/conv0", b.act, b.pad, next_pad=b.pad)
                    for r in range(1, b.n_convs):
                        t = m._resnet_tc(t, "%s/resnet%d" % (name, r), b.act, b.pad, b.use_batchnorm,
                                         next_pad=b.pad if r + 1 < b.n_convs else PAD_CONSTANT)
                    if b.kind != "deconv":
                        t = ops.from_blk8(t)
                else:
                    t = m._bottleneck_lowres(b, None, pooled)
                    if b.kind == "deconv":
                        t = ops.to_blk8(t)
                rep_out[(b.kind, b.index)] = t
        del full
        self._tick("branches (replicated)")
        dc, rs = [], []
        for b in blocks:
            name = "bottleneck_%s/%d" % (b.kind, b.index)
            if b.kind != "deconv":
                rs.append((rep_out[(b.kind, b.index)], b.resize_method))
                continue
            s_ = b.downsampling_factor
            if s_ in banded:
                t = self._each(lambda i: ops.to_blk8(band_pool[s_][i], split=bsplit, halo=b.pad))
                self._exchange(t)
                t = self._conv(t, name + "/conv0", b.act, b.pad, next_pad=b.pad)
                for r in range(1, b.n_convs):
                    t = self._resnet(t, "%s/resnet%d" % (name, r), b.act, b.pad, b.use_batchnorm,
                                     next_pad=b.pad if r + 1 < b.n_convs else PAD_CONSTANT)
            else:                                   # this band's rows of the replicated low-resolution output
                full = rep_out[(b.kind, b.index)]
                vf = _view(full)

                def rows(i):
                    lo_, hl = bnd[i] // s_, -(-hs[i] // s_)
                    t_ = ops.Blk8(full.B, full.C, hl, full.W, full.device, split=1)
                    _view(t_)[:, :, HALO:HALO + hl].copy_(vf[:, :, HALO + lo_:HALO + lo_ + hl])
                    return t_
                t = self._each(rows)
            dk, db = m.conv(name + "/deconv")
            key = ("deconv_packed_tc", dk.data_ptr())
            if key not in m._tc:
                m._tc[key] = ops.pack_deconv_kernel_tc(dk)
            dc.append((t, m._tc[key], db, b.upsampling_factor, b.deconv_act))
        self._tick("branches (banded)")
        return {"dc": dc, "rs": rs, "um": um, "alpha": 1.0 / float(len(blocks) * F)}

    def _merge_banded(self, br, cat, H, Wd):
        F = self.m.filters

        def one(i):
            r0, hb = self.bounds[i], self.bounds[i + 1] - self.bounds[i]
            packed = [(t[i], k, b_, s_, a_) for t, k, b_, s_, a_ in br["dc"]]
            ops.upsample_merge_tc_blk8(packed, br["rs"] if br["um"] == 1 else [], br["alpha"], cat[i], F, hb, Wd, row_offset=r0, full_H=H)
            if br["um"] == 2:
                ops.resize_add_blk8(br["rs"], br["alpha"], cat[i], F, hb, Wd, row_offset=r0, full_H=H)
            return None
        self._each(one)

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
        bnd = band_bounds(H, P, [b_.upsampling_factor for b_ in m.bottleneck_deconv_blocks])
        self.bounds = bnd
        F = m.filters
        dev = rhs.device
        split = m.tc_split
        rhs = rhs.contiguous()
        with torch.cuda.device(dev):
            self._marks = []
            self._tick("start")
            posx, posy = ops.position_table(dev, H), ops.position_table(dev, Wd)

            def first(i):
                r0, h = bnd[i], bnd[i + 1] - bnd[i]
                band = rhs[:, :, r0:r0 + h].contiguous()
                if m.use_positional_embeddings:
                    x = torch.empty((B, 3, h, Wd), device=dev, dtype=torch.float32)
                    ops.check(ops.lib.pcnn_hpnn_input_f32(band.data_ptr(), posx[r0:].data_ptr(), posy.data_ptr(), x.data_ptr(), B, h, Wd,
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
            br = self._branches(x0, H, Wd, split)
            cat = self._each(lambda i: ops.Blk8(B, 2 * F, bnd[i + 1] - bnd[i], Wd, dev, split=split))
            self._conv(x0, "non_bottleneck_conv", ACT_LEAKY_RELU, PAD_CONSTANT, out=cat)
            if br is not None:
                self._merge_banded(br, cat, H, Wd)
                self._tick("upsample-merge (banded)")
            else:
                # general merge kernels (no fused tensor-core merge for this config / shape): everything replicated from the
                # gathered features, each rank keeps its rows of the merged map
                x0_full = self._gather_rows(self._each(lambda i: ops.from_blk8(x0[i])))
                with ops.blk8_pool_scope(self.full_pool):
                    branches = m._branches_tc(x0_full, H, Wd, split)
                    cat_full = ops.Blk8(B, 2 * F, H, Wd, dev, split=split)
                    m._merge_tc(branches, cat_full, B, H, Wd, dev)
                p0 = (F // 16) * 2                       # first plane of channels [F, 2F)
                for which in ("buf", "lo"):
                    if getattr(cat_full, which) is None:
                        continue
                    vf = _view(cat_full, which)
                    for i in self.local:
                        _view(cat[i], which)[:, p0:, HALO:HALO + bnd[i + 1] - bnd[i]].copy_(vf[:, p0:, HALO + bnd[i]:HALO + bnd[i + 1]])
                del cat_full, branches, x0_full
                self._tick("branches + merge (replicated)")
            self._exchange(cat)
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
