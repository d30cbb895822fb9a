"""Homogeneous_Poisson_NN_Legacy on hand-written CUDA kernels.

Same constructor kwargs, call convention and error behaviour as the reference class
(poisson_CNN/models/Homogeneous_Poisson_NN_Legacy.py:10-257); every tensor op goes through
libpcnn.so.  model([rhs, dx]) -> [B,1,H,W] on the device/stream of the inputs, no host sync.
"""
import copy

import os

import torch

from .. import ops
from .. import weights as W
from ..config import (activation_enum, padding_enum, resize_enum, get_init_arguments_from_config,
                      process_normalizations, process_output_scaling_modes, bottleneck_output_size,
                      ACT_LEAKY_RELU, ACT_LINEAR, PAD_CONSTANT)
from ._base import WeightedModel


class _Bottleneck:
    """One resolution branch (poisson_CNN/blocks/bottleneck_block.py:8-118), config only."""

    def __init__(self, kind, index, ndims, downsampling_factor, filters, conv_kernel_size, deconv_kernel_size=None,
                 data_format="channels_first", conv_activation=None, conv_use_bias=True, use_resnet=False,
                 padding_mode="constant", constant_padding_value=0.0, n_convs=1, upsampling_factor=None,
                 conv_initializer_constraint_regularizer_options=None, downsampling_method="conv",
                 conv_downsampling_kernel_size=None, pool_downsampling_method="max", use_batchnorm=False,
                 batchnorm_trainable=True, resize_method="bilinear", deconv_activation=None, deconv_use_bias=True,
                 deconv_initializer_constraint_regularizer_options=None):
        self.kind, self.index = kind, index
        self.downsampling_factor = downsampling_factor
        self.upsampling_factor = downsampling_factor if upsampling_factor is None else upsampling_factor
        self.filters = filters
        self.ksize = conv_kernel_size
        self.deconv_ksize = deconv_kernel_size
        self.act = activation_enum(conv_activation)
        self.deconv_act = activation_enum(deconv_activation)
        self.pad = padding_enum(padding_mode)
        self.pad_value = constant_padding_value
        self.n_convs = n_convs
        self.use_batchnorm = use_batchnorm
        self.resize_method = resize_enum(resize_method) if kind == "multilinear" else None
        method = downsampling_method.lower()
        if method not in ("conv", "pool"):
            raise ValueError("Downsampling method can only be conv or pool")
        if method != "pool" or not use_resnet:
            raise NotImplementedError("only downsampling_method='pool' with use_resnet=True (every shipped config) is built")
        if pool_downsampling_method.lower() not in ("average", "avg"):
            raise NotImplementedError("only average pooling (every shipped config) is built")


class Homogeneous_Poisson_NN_Legacy(WeightedModel):
    # the single-pass 'tc' mode misses the 2e-3 budget on this 45-convolution-deep network (measured 4.4e-3 at 256x256)
    COMPLIANT_PRECISIONS = ("fp32", "tc2", "tc3", "mixed")

    def __init__(self, data_format="channels_first", final_convolutions_config=None,
                 pre_bottleneck_convolutions_config=None, bottleneck_deconv_config=None,
                 bottleneck_multilinear_config=None, input_normalization=None, output_scaling=None,
                 use_batchnorm=False, postsmoother_iterations=5, use_scaling=False,
                 use_positional_embeddings=True, scaling_config=None, gradient_accumulation_steps=None,
                 bc_type="dirichlet"):
        super().__init__()
        self.ndims = 2
        if data_format != "channels_first":
            raise NotImplementedError("the CUDA path is channels_first (every shipped config)")
        self.data_format = data_format
        self.gradient_accumulation_steps = gradient_accumulation_steps
        self.input_normalization = process_normalizations(input_normalization)
        self.output_scaling = process_output_scaling_modes(output_scaling)
        self.use_batchnorm = use_batchnorm
        self.use_positional_embeddings = use_positional_embeddings

        if pre_bottleneck_convolutions_config is None:
            raise ValueError("Provide a config for pre bottleneck convolutions")
        if (bottleneck_deconv_config is None) or (bottleneck_multilinear_config is None):
            raise ValueError("Provide a config for bottleneck blocks")
        if final_convolutions_config is None:
            raise ValueError("Provide a config for final convolutions")
        if use_scaling and scaling_config is None:
            raise ValueError("use_scaling=True needs a scaling_config")
        if bc_type.lower() not in ("dirichlet", "neumann"):
            raise ValueError("bc_type can only be neumann or dirichlet.")
        self.bc_type = ops.BC_DIRICHLET if bc_type.lower() == "dirichlet" else ops.BC_NEUMANN

        # keep a JSON-like copy for weight_specs()
        self._cfg = copy.deepcopy({
            "use_batchnorm": use_batchnorm, "use_scaling": use_scaling,
            "use_positional_embeddings": use_positional_embeddings,
            "pre_bottleneck_convolutions_config": pre_bottleneck_convolutions_config,
            "bottleneck_deconv_config": bottleneck_deconv_config,
            "bottleneck_multilinear_config": bottleneck_multilinear_config,
            "final_convolutions_config": final_convolutions_config, "scaling_config": scaling_config,
            "postsmoother_iterations": postsmoother_iterations, "bc_type": bc_type.lower(), "data_format": data_format})

        pre = copy.deepcopy(pre_bottleneck_convolutions_config)
        self.pre_pad = padding_enum(pre.pop("padding_mode", "CONSTANT"))
        self.pre_pad_value = pre.pop("constant_padding_value", 0.0)
        self.pre_act = activation_enum(pre.get("activation"))
        self.n_pre = len(pre["filters"])

        assert bottleneck_deconv_config["filters"] == bottleneck_multilinear_config["filters"]
        self.filters = bottleneck_deconv_config["filters"]
        f_cfg = ["downsampling_factors", "upsampling_factors", "conv_kernel_sizes", "deconv_kernel_sizes", "n_convs"]
        f_arg = ["downsampling_factor", "upsampling_factor", "conv_kernel_size", "deconv_kernel_size", "n_convs"]
        dcfg = bottleneck_deconv_config
        self.bottleneck_deconv_blocks = [
            _Bottleneck("deconv", k, ndims=2, data_format=data_format, use_batchnorm=use_batchnorm,
                        **get_init_arguments_from_config(dcfg, k, f_cfg, f_arg))
            for k in range(len(dcfg["downsampling_factors"]))]
        self.bottleneck_deconv_blocks.sort(key=lambda b: b.downsampling_factor, reverse=True)
        mcfg = bottleneck_multilinear_config
        f_cfg = ["downsampling_factors", "upsampling_factors", "conv_kernel_sizes", "n_convs"] + (["resize_methods"] if "resize_methods" in mcfg else [])
        f_arg = ["downsampling_factor", "upsampling_factor", "conv_kernel_size", "n_convs"] + (["resize_method"] if "resize_methods" in mcfg else [])
        self.bottleneck_multilinear_blocks = [
            _Bottleneck("multilinear", k, ndims=2, data_format=data_format, use_batchnorm=use_batchnorm,
                        **get_init_arguments_from_config(mcfg, k, f_cfg, f_arg))
            for k in range(len(mcfg["downsampling_factors"]))]
        self.bottleneck_multilinear_blocks.sort(key=lambda b: b.downsampling_factor, reverse=True)

        fin = copy.deepcopy(final_convolutions_config)
        self.final_pad = padding_enum(fin.pop("padding_mode", "CONSTANT"))
        self.final_pad_value = fin.pop("constant_padding_value", 0.0)
        self.final_regular_conv_stages = fin.pop("final_regular_conv_stages", 2)
        self.final_act = activation_enum(fin.get("activation"))
        self.n_final = len(fin["filters"])

        self.postsmoother_iterations = postsmoother_iterations
        self.use_scaling = use_scaling
        if use_scaling:
            sc = dict(scaling_config)
            self.scaling_stages = sc.get("stages", 2)
            self.scaling_ratio = sc.get("downsampling_ratio_per_stage", 2)
            self.scaling_levels = copy.deepcopy(sc.get("spp_levels", [[2, 2], 3, 5]))
            self.scaling_act = activation_enum(sc.get("activation"))
            if sc.get("kernel_size", 3) % 2 == 0:
                raise NotImplementedError("Scaling: even kernel sizes with Keras 'same' padding are not built")

    def keras_key_map(self, prefix=""):
        from .. import tf_checkpoint as T
        return T.hpnn_key_map(self._cfg, "", prefix)

    # layer-name prefixes whose tc2 correction pass is skipped (precision experiments / 'mixed' mode)
    tc_uncorrected = ()

    def weight_specs(self, prefix=""):
        return W.hpnn_weight_specs(self._cfg, prefix)

    # ------------------------------------------------------------------ building blocks
    def _resnet(self, x, name, act, pad, pad_value, use_bn, out_scale=None):
        """blocks/resnet.py:29-39 with BN and the residual add fused into the conv epilogues."""
        k0, b0 = self.conv(name + "/conv0")
        k1, b1 = self.conv(name + "/conv1")
        k2, b2 = self.conv(name + "/conv2")
        t = ops.conv2d(x, k0, b0, act, pad, pad_value, bn=self.bn(name + "/bn0") if use_bn else None)
        t = ops.conv2d(t, k1, b1, act, pad, pad_value, bn=self.bn(name + "/bn1") if use_bn else None, residual=x)
        return ops.conv2d(t, k2, b2, act, pad, pad_value, out_scale=out_scale)

    def _branch_layers(self, blk, name):
        """Layer program of a bottleneck branch for ops.smallmap_stack (flags: 1 save input, 2 add saved)."""
        key = ("branch_stack", name)
        if key not in self._tc:                  # cleared whenever weights are (re)loaded
            k, b = self.conv(name + "/conv0")
            layers = [{"kernel": k, "bias": b, "bn": None, "flags": 0}]
            for r in range(1, blk.n_convs):
                rn = "%s/resnet%d" % (name, r)
                for c, fl in enumerate((1, 2, 0)):
                    kc, bc_ = self.conv("%s/conv%d" % (rn, c))
                    bn = self.bn("%s/bn%d" % (rn, c)) if (blk.use_batchnorm and c < 2) else None
                    layers.append({"kernel": kc, "bias": bc_, "bn": bn, "flags": fl})
            self._tc[key] = layers
        return self._tc[key]

    def _branch_out_hw(self, blk, H, Wd):
        out_hw = (bottleneck_output_size(H, blk.downsampling_factor, blk.upsampling_factor),
                  bottleneck_output_size(Wd, blk.downsampling_factor, blk.upsampling_factor))
        if out_hw != (H, Wd):
            raise ValueError("bottleneck branch ds=%d would produce %s for a %s grid; the merge needs equal shapes"
                             % (blk.downsampling_factor, out_hw, (H, Wd)))
        return out_hw

    @staticmethod
    def _pool_pyramid(x0, factors):
        """{s: AveragePooling2D(s, 'same')(x0)} for every downsampling factor.  Where the windows nest exactly (the
        grid is divisible by s, so SAME pooling needs no padding at either level) level s is pooled from the largest
        already computed level that divides it -- means of equally sized blocks compose exactly -- so x0 is read
        once or twice instead of once per branch."""
        H, W = x0.shape[2], x0.shape[3]
        levels = {}
        for s in sorted(set(factors)):
            src, f = x0, s
            if H % s == 0 and W % s == 0:
                for t in sorted(levels, reverse=True):
                    if s % t == 0 and H % t == 0 and W % t == 0:
                        src, f = levels[t], s // t
                        break
            levels[s] = ops.avgpool_same(src, f)
        return levels

    def _bottleneck_lowres(self, blk, x0, pooled=None):
        """pool -> conv -> resnets of one branch (blocks/bottleneck_block.py:36-50), FP32 kernels."""
        name = "bottleneck_%s/%d" % (blk.kind, blk.index)
        h = pooled if pooled is not None else ops.avgpool_same(x0, blk.downsampling_factor)
        stack = self._branch_layers(blk, name)
        if ops.smallmap_stack_supported(h.shape[2], h.shape[3], stack):
            # 2x2 .. 8x8 maps: the whole conv + resnet chain in ONE kernel, activations in shared memory
            return ops.smallmap_stack(h, stack, blk.act, blk.pad, blk.pad_value)
        k, b = self.conv(name + "/conv0")
        h = ops.conv2d(h, k, b, blk.act, blk.pad, blk.pad_value)
        for r in range(1, blk.n_convs):
            h = self._resnet(h, "%s/resnet%d" % (name, r), blk.act, blk.pad, blk.pad_value, blk.use_batchnorm)
        return h

    def _bottleneck(self, blk, x0, merged, first, alpha):
        name = "bottleneck_%s/%d" % (blk.kind, blk.index)
        out_hw = self._branch_out_hw(blk, x0.shape[2], x0.shape[3])
        h = self._bottleneck_lowres(blk, x0)
        if blk.kind == "deconv":
            dk, db = self.conv(name + "/deconv")
            ops.deconv_same(h, dk, db, out_hw, blk.upsampling_factor, blk.deconv_act, alpha, out=merged, accumulate=not first)
        else:
            ops.resize(h, out_hw, blk.resize_method, alpha, out=merged, accumulate=not first)

    # ------------------------------------------------------------------ tensor-core building blocks
    def _conv_tc(self, x, name, act, pad, bn_name=None, residual=None, out_scale=None, out=None, out_c_offset=0,
                 next_pad=PAD_CONSTANT):
        """next_pad: padding mode of the layer that reads the result (a SYMMETRIC ring is written by this conv's epilogue).
        The precision mode follows the input tensor (bottleneck branches may run single-pass inside a tc2 network)."""
        wp, bias = self.tc_conv(name, mode=x.mode)
        corr = not any(name.startswith(pfx) for pfx in self.tc_uncorrected)
        return ops.conv2d_tc(x, wp, bias, act, pad, bn=self.bn(bn_name) if bn_name else None, residual=residual,
                             out_scale=out_scale, out=out, out_c_offset=out_c_offset, out_halo=next_pad, correction=corr)

    def _resnet_tc(self, x, name, act, pad, use_bn, out_scale=None, next_pad=PAD_CONSTANT):
        t = self._conv_tc(x, name + "/conv0", act, pad, name + "/bn0" if use_bn else None, next_pad=pad)
        t = self._conv_tc(t, name + "/conv1", act, pad, name + "/bn1" if use_bn else None, residual=x, next_pad=pad)
        return self._conv_tc(t, name + "/conv2", act, pad, out_scale=out_scale, next_pad=next_pad)

    def _tc_ok(self, ksize, pad_value=0.0):
        return ksize % 2 == 1 and ksize <= 15 and float(pad_value) == 0.0

    def _branch_on_tc(self, blk, ph, pw):
        """A branch's conv + resnet chain runs on the tensor cores when its pooled map is at least 16 pixels a side (the
        transpose-conv branches at every shipped size; the multilinear ones from ~512-pixel grids on: 64x64 ... 16x16 maps at
        2048^2, which the FP32 kernels ran at 8 CTAs per launch); smaller maps take the fused FP32 stack kernels."""
        return min(ph, pw) >= 16 and self._tc_ok(blk.ksize, blk.pad_value) and blk.pad in (PAD_CONSTANT, 1)

    def _branches_tc(self, x0_f32, H, Wd, split):
        """pool -> conv -> resnets of every bottleneck branch from the full-resolution NCHW fp32 features (the pooling pyramid
        reads them); returns the low-res branch outputs for _merge_tc."""
        F = self.filters
        blocks = self.bottleneck_deconv_blocks + self.bottleneck_multilinear_blocks
        alpha = 1.0 / float(len(blocks) * F)
        dc, rs = [], []                          # low-res branch outputs, upsampled and summed by ONE fused kernel
        # 'mixed': the branches are averaged with weight 1/(8*F) before they re-enter the trunk -- measured: running
        # them single-pass changes the merged model's error by < 1e-6 (tests/probes/layer_sensitivity_probe.py)
        bsplit = 1 if self.requested_precision == "mixed" else split
        pools = self._pool_pyramid(x0_f32, [blk.downsampling_factor for blk in blocks])
        # single-pass branches with 32 filters feed the tensor-core upsample-merge kernel as BLK8 fp16 (no fp32 copy)
        um_tc = (bsplit == 1 and F == 32 and os.environ.get("PCNN_UM_TC", "1") != "0"
                 and any(b_.kind == "deconv" for b_ in blocks) and len(blocks) <= 16
                 and all(b_.upsampling_factor <= 32 for b_ in blocks if b_.kind == "deconv"))
        if um_tc:
            # 1: everything in the fused kernel (its resize sources are staged whole in shared memory: grids up to ~400 pixels a
            # side); 2: larger grids -- the transpose-conv branches in the fused kernel, the resize branches added by a second
            # pass (ops.resize_add_blk8); 0: general kernels
            strides = [b_.upsampling_factor for b_ in blocks if b_.kind == "deconv"]
            rs_hw = [(-(-H // b_.downsampling_factor), -(-Wd // b_.downsampling_factor)) for b_ in blocks if b_.kind != "deconv"]
            um_tc = 1 if ops.upsample_merge_tc_fits(strides, rs_hw) else (2 if (rs_hw and ops.upsample_merge_tc_fits(strides, [])) else 0)
        for blk in blocks:
            self._branch_out_hw(blk, H, Wd)
            ph, pw = -(-H // blk.downsampling_factor), -(-Wd // blk.downsampling_factor)
            name = "bottleneck_%s/%d" % (blk.kind, blk.index)
            if self._branch_on_tc(blk, ph, pw):
                h = ops.to_blk8(pools[blk.downsampling_factor], split=bsplit, halo=blk.pad)
                h = self._conv_tc(h, name + "/conv0", blk.act, blk.pad, next_pad=blk.pad)
                for r in range(1, blk.n_convs):
                    h = self._resnet_tc(h, "%s/resnet%d" % (name, r), blk.act, blk.pad, blk.use_batchnorm,
                                        next_pad=blk.pad if r + 1 < blk.n_convs else PAD_CONSTANT)
                if not um_tc or blk.kind != "deconv":
                    h = ops.from_blk8(h)         # the resize branches and the non-tensor-core merges read NCHW fp32
            else:
                h = self._bottleneck_lowres(blk, x0_f32, pools[blk.downsampling_factor])
                if um_tc and blk.kind == "deconv":
                    h = ops.to_blk8(h)           # tiny map (< 16 pixels a side) computed by the FP32 stack kernel
            if blk.kind == "deconv":
                dk, db = self.conv(name + "/deconv")
                dc.append((h, dk, db, blk.upsampling_factor, blk.deconv_act))
            else:
                rs.append((h, blk.resize_method))

        return dc, rs, um_tc, alpha

    def _merge_tc(self, branches, cat, B, H, Wd, dev):
        """All upsamplings + the branch sum written into channels [F, 2F) of the BLK8 tensor `cat` (one fused kernel)."""
        dc, rs, um_tc, alpha = branches
        F = self.filters
        fused_dc = (F % 8 == 0 and len(dc) <= 8 and len(rs) <= 8
                    and all(tuple(k.shape) == (s_, s_, F, F) and s_ <= 32 for _, k, _, s_, _ in dc))
        fused = fused_dc and all(F * h.shape[2] * h.shape[3] <= 8192 for h, _ in rs)
        if (um_tc == 1 and fused) or (um_tc == 2 and fused_dc):
            packed = []
            for h, k, bias_, s_, act_ in dc:
                key = ("deconv_packed_tc", k.data_ptr())
                if key not in self._tc:
                    self._tc[key] = ops.pack_deconv_kernel_tc(k)
                packed.append((h, self._tc[key], bias_, s_, act_))
            ops.upsample_merge_tc_blk8(packed, rs if um_tc == 1 else [], alpha, cat, F, H, Wd)
            if um_tc == 2:
                ops.resize_add_blk8(rs, alpha, cat, F, H, Wd)
        elif fused:
            dc = [(ops.from_blk8(h) if isinstance(h, ops.Blk8) else h, k, b_, s_, a_) for h, k, b_, s_, a_ in dc]
            packed = []
            for h, k, bias_, s_, act_ in dc:
                key = ("deconv_packed", k.data_ptr())
                if key not in self._tc:          # packed once per layer; the cache dies with the weights (_on_weights_loaded)
                    self._tc[key] = ops.pack_deconv_kernel(k)
                packed.append((h, self._tc[key], bias_, s_, act_))
            ops.upsample_merge_blk8(packed, rs, alpha, cat, F, H, Wd)
        else:                                    # general kernels: fp32 merge buffer, one read-modify-write per branch
            merged = torch.empty((B, F, H, Wd), device=dev, dtype=torch.float32)
            for i, (h, dk, db, s_, act_) in enumerate(dc):
                h = ops.from_blk8(h) if isinstance(h, ops.Blk8) else h
                ops.deconv_same(h, dk, db, (H, Wd), s_, act_, alpha, out=merged, accumulate=(i != 0))
            for i, (h, method) in enumerate(rs):
                ops.resize(h, (H, Wd), method, alpha, out=merged, accumulate=(i != 0 or bool(dc)))
            ops.to_blk8(merged, out=cat, c_offset=F)

    def _call_tc(self, rhs, dx):
        """Same graph as the FP32 path with every heavy convolution on tcgen05 (BLK8 fp16 activations).
        FP32 kernels keep: pooling / upsampling / merge, the 2..8-pixel multilinear branches, Scaling and
        the boundary ring."""
        B, _, H, Wd = rhs.shape
        F = self.filters
        dev = rhs.device
        x = ops.hpnn_input(rhs) if self.use_positional_embeddings else rhs
        split = self.tc_split
        t = ops.to_blk8(x, split=split, halo=self.pre_pad)
        for k in range(self.n_pre):
            t = self._conv_tc(t, "pre_bottleneck/%d" % k, self.pre_act, self.pre_pad,
                              "pre_bottleneck/%d/bn" % k if self.use_batchnorm else None,
                              next_pad=self.pre_pad if k + 1 < self.n_pre else PAD_CONSTANT)
        x0 = t                                   # BLK8, F channels
        x0_f32 = ops.from_blk8(x0)               # pooling pyramid reads NCHW fp32

        branches = self._branches_tc(x0_f32, H, Wd, split)
        cat = ops.Blk8(B, 2 * F, H, Wd, dev, split=split)
        self._conv_tc(x0, "non_bottleneck_conv", ACT_LEAKY_RELU, PAD_CONSTANT, out=cat)
        self._merge_tc(branches, cat, B, H, Wd, dev)
        y = self._conv_tc(cat, "post_merge_conv", ACT_LEAKY_RELU, PAD_CONSTANT)

        d = ops.dense_input(dx, H, Wd)
        d = ops.dense(d, *self.conv("dx_dense/0"), ACT_LEAKY_RELU)
        d = ops.dense(d, *self.conv("dx_dense/1"), ACT_LEAKY_RELU)
        d = ops.dense(d, *self.conv("dx_dense/2"), ACT_LINEAR)
        y = self._resnet_tc(y, "post_merge_resnet", ACT_LEAKY_RELU, PAD_CONSTANT, False, out_scale=d)

        S, nreg = self.n_final, self.final_regular_conv_stages
        for k in range(S - nreg):
            y = self._conv_tc(y, "final/%d/conv" % k, self.final_act, self.final_pad)
            y = self._resnet_tc(y, "final/%d/resnet" % k, self.final_act, PAD_CONSTANT, False)
        for k in range(S - nreg, S):             # the last linear convs: 16 output rows x 8 channel slots per tile
            y = self._conv_tc(y, "final/%d/conv" % k, ACT_LINEAR, PAD_CONSTANT)
        cat2 = torch.empty((B, 2, H, Wd), device=dev, dtype=torch.float32) if self.use_scaling and y.C == 1 else None
        y = ops.from_blk8(y, C=y.C, out=None if cat2 is None else cat2[:, 0:1])
        return self._tail(y, rhs, dx, S, cat2)

    def _tail(self, y, rhs, dx, first_regular, cat2=None):
        """The last linear convs (strict mode), Scaling, boundary ring and post-smoother (FP32 kernels)."""
        B, _, H, Wd = rhs.shape
        S = self.n_final
        if cat2 is None and self.use_scaling:
            cat2 = torch.empty((B, 2, H, Wd), device=rhs.device, dtype=torch.float32)
        for k in range(first_regular, S):
            kk, bb = self.conv("final/%d/conv" % k)
            last = (k == S - 1) and self.use_scaling and kk.shape[3] == 1
            y = ops.conv2d(y, kk, bb, ACT_LINEAR, PAD_CONSTANT, 0.0, out=cat2[:, 0:1] if last else None)
        s = None
        if self.use_scaling:
            if y.data_ptr() != cat2.data_ptr():
                cat2[:, 0:1].copy_(y)
                y = cat2[:, 0:1]
            cat2[:, 1:2].copy_(rhs)
            h = cat2
            for st in range(self.scaling_stages):
                kk, bb = self.conv("scaling/conv%d" % st)
                h = ops.conv2d(h, kk, bb, self.scaling_act, PAD_CONSTANT, 0.0)
                h = ops.avgpool_same(h, self.scaling_ratio)
            v = ops.spatial_pyramid_pool(h, self.scaling_levels, ops.POOL_MAX, ndims=2)
            v = ops.dense(v, *self.conv("scaling/dense0"), ACT_LEAKY_RELU)
            v = ops.dense(v, *self.conv("scaling/dense1"), ACT_LEAKY_RELU)
            s = ops.dense(v, *self.conv("scaling/dense2"), ACT_LINEAR).reshape(B)
        out = ops.hpnn_finalize(y, s, self.bc_type)
        if self.postsmoother_iterations > 0:
            out = ops.jacobi(out, rhs, torch.cat([dx, dx], 1), self.postsmoother_iterations)
        return out

    def _tc_supported(self):
        ks = list(self._cfg["pre_bottleneck_convolutions_config"]["kernel_sizes"]) + \
            list(self._cfg["final_convolutions_config"]["kernel_sizes"]) + \
            list(self._cfg["bottleneck_deconv_config"]["conv_kernel_sizes"])
        return (all(self._tc_ok(k) for k in ks) and self.filters <= 32 and float(self.pre_pad_value) == 0.0
                and float(self.final_pad_value) == 0.0 and self.pre_pad in (0, 1) and self.final_pad == 0
                and max(self._cfg["final_convolutions_config"]["filters"]) <= 32 and self.n_pre >= 2)

    # ------------------------------------------------------------------ forward
    def __call__(self, inp):
        rhs, dx = inp
        if rhs.dim() != 4 or rhs.shape[1] != 1:
            raise ValueError("rhs must be [batch, 1, nx, ny] (channels_first)")
        if dx.dim() != 2 or dx.shape[1] != 1 or dx.shape[0] != rhs.shape[0]:
            raise ValueError("dx must be [batch, 1]")
        if not rhs.is_cuda:
            raise ValueError("rhs must be a CUDA tensor (the hot path has no CPU implementation)")
        with torch.cuda.device(rhs.device):     # launches go to the current device: make it the tensors' device
            return self._forward(rhs, dx)

    def _engine_config(self):
        return {"hpnn_model": self._cfg}

    def _engine_weights(self):
        return self.get_weights_dict("hpnn/")

    def _forward(self, rhs, dx):
        if self.use_engine and not self.tc_uncorrected:
            return self.engine().hpnn_forward(rhs, dx)
        B, _, H, Wd = rhs.shape
        F = self.filters
        if self.precision in ("tc", "tc2", "tc3"):
            if not self._tc_supported():
                raise NotImplementedError("precision='tc' covers odd kernels <= 15, <= 32 filters, zero CONSTANT / SYMMETRIC padding")
            return self._call_tc(rhs, dx)

        x = ops.hpnn_input(rhs) if self.use_positional_embeddings else rhs
        for k in range(self.n_pre):
            kk, bb = self.conv("pre_bottleneck/%d" % k)
            x = ops.conv2d(x, kk, bb, self.pre_act, self.pre_pad, self.pre_pad_value,
                           bn=self.bn("pre_bottleneck/%d/bn" % k) if self.use_batchnorm else None)
        x0 = x

        # concat(non_bottleneck_conv(x0), merged) is assembled in place: channels [0,F) and [F,2F)
        cat = torch.empty((B, 2 * F, H, Wd), device=rhs.device, dtype=torch.float32)
        merged = cat[:, F:]
        blocks = self.bottleneck_deconv_blocks + self.bottleneck_multilinear_blocks
        alpha = 1.0 / float(len(blocks) * F)
        for i, blk in enumerate(blocks):
            self._bottleneck(blk, x0, merged, i == 0, alpha)
        kk, bb = self.conv("non_bottleneck_conv")
        ops.conv2d(x0, kk, bb, ACT_LEAKY_RELU, PAD_CONSTANT, 0.0, out=cat[:, :F])
        kk, bb = self.conv("post_merge_conv")
        y = ops.conv2d(cat, kk, bb, ACT_LEAKY_RELU, PAD_CONSTANT, 0.0)

        # dx MLP -> per-(sample, channel) scale, applied in the epilogue of post_merge_resnet's last conv
        d = ops.dense_input(dx, H, Wd)
        d = ops.dense(d, *self.conv("dx_dense/0"), ACT_LEAKY_RELU)
        d = ops.dense(d, *self.conv("dx_dense/1"), ACT_LEAKY_RELU)
        d = ops.dense(d, *self.conv("dx_dense/2"), ACT_LINEAR)
        y = self._resnet(y, "post_merge_resnet", ACT_LEAKY_RELU, PAD_CONSTANT, 0.0, False, out_scale=d)

        S, nreg = self.n_final, self.final_regular_conv_stages
        for k in range(S - nreg):
            kk, bb = self.conv("final/%d/conv" % k)
            y = ops.conv2d(y, kk, bb, self.final_act, self.final_pad, self.final_pad_value)
            y = self._resnet(y, "final/%d/resnet" % k, self.final_act, PAD_CONSTANT, 0.0, False)
        return self._tail(y, rhs, dx, S - nreg)

    call = __call__
