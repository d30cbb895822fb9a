"""Dirichlet_BC_NN_Legacy_2 on hand-written CUDA kernels.

Same constructor kwargs / call convention as poisson_CNN/models/Dirichlet_BC_NN_Legacy.py:14-166:
model([bc, dx, x_output_resolution]) with bc [B,1,n], dx [B,1] -> [B,1,x_res,n].
"""
import copy
import warnings

import os

import torch

from .. import ops
from .. import weights as W
from ..config import activation_enum, padding_enum, PAD_CONSTANT, ACT_TANH
from ._base import WeightedModel


class Dirichlet_BC_NN_Legacy_2(WeightedModel):
    _MIXED_AS = "tc"       # precision "mixed": this network runs single-pass FP16 (see WeightedModel.set_precision)

    def __init__(self, data_format="channels_first", boundary_conv_config=None, spp_config=None,
                 domain_info_mlp_config=None, final_convolutions_config=None, postsmoother_iterations=0,
                 use_batchnorm=False):
        super().__init__()
        self.ndims = 2
        if data_format != "channels_first":
            raise NotImplementedError("the CUDA path is channels_first (every shipped config)")
        self.data_format = data_format
        self.use_batchnorm = use_batchnorm
        if boundary_conv_config is None:
            raise ValueError("Provide a config for the boundary convolutions.")
        if spp_config is None:
            raise ValueError("Provide a config for the Spatial Pyramid Pooling.")
        if final_convolutions_config is None:
            raise ValueError("Provide a config for the domain convolutions.")
        if domain_info_mlp_config is None:
            raise ValueError("Provide a config for the domain info MLP.")
        assert boundary_conv_config["filters"][-1] == domain_info_mlp_config["units"][-1]
        if boundary_conv_config["filters"][-1] > 27:
            warnings.warn(str(boundary_conv_config["filters"][-1]) + " sinh modes chosen may lead to NaN values with float32 precision. Consider using fewer than 28 when using float32.")
        self.x_dir_nmodes = domain_info_mlp_config["units"][-1]

        self._cfg = copy.deepcopy({
            "use_batchnorm": use_batchnorm, "boundary_conv_config": boundary_conv_config, "spp_config": spp_config,
            "domain_info_mlp_config": domain_info_mlp_config, "final_convolutions_config": final_convolutions_config,
            "postsmoother_iterations": postsmoother_iterations, "data_format": data_format})

        bcfg = copy.deepcopy(boundary_conv_config)
        self.boundary_pad = padding_enum(bcfg.pop("padding_mode", "CONSTANT"))
        self.boundary_pad_value = bcfg.pop("constant_padding_value", 0.0)
        self.boundary_act = activation_enum(bcfg.get("activation"))
        self.n_boundary = len(bcfg["filters"])

        self.spp_levels = copy.deepcopy(spp_config["levels"])
        ptype = spp_config.get("pooling_type", "average").lower()
        if ptype in ("average", "avg"):
            self.spp_mode = ops.POOL_AVG
        elif ptype == "max":
            self.spp_mode = ops.POOL_MAX
        else:
            raise ValueError("unknown SPP pooling_type " + ptype)

        self.mlp_acts = [activation_enum(a) for a in domain_info_mlp_config["activations"]]
        self.n_mlp = len(domain_info_mlp_config["units"])

        fin = copy.deepcopy(final_convolutions_config)
        self.final_pad = padding_enum(fin.pop("padding_mode", "CONSTANT"))
        self.final_pad_value = fin.pop("constant_padding_value", 0.0)
        self.final_regular_conv_stages = fin.pop("final_regular_conv_stages", 2)
        self.final_act = activation_enum(fin.get("activation"))
        self.n_final = len(fin["filters"])
        self.postsmoother_iterations = postsmoother_iterations

    def weight_specs(self, prefix=""):
        return W.dbcnn_weight_specs(self._cfg, prefix)

    def keras_key_map(self, prefix=""):
        from .. import tf_checkpoint as T
        return T.dbcnn_key_map(self._cfg, "", prefix)

    def _engine_config(self):
        return {"dbcnn_model": self._cfg}

    def _engine_weights(self):
        return self.get_weights_dict("dbcnn/")

    def _resnet1d(self, x, name, act, pad, pad_value, use_bn):
        k0, b0 = self.conv(name + "/conv0")
        k1, b1 = self.conv(name + "/conv1")
        k2, b2 = self.conv(name + "/conv2")
        t = ops.conv1d(x, k0, b0, act, pad, pad_value, bn=self.bn(name + "/bn0") if use_bn else None)
        t = ops.conv1d(t, k1, b1, act, pad, pad_value, bn=self.bn(name + "/bn1") if use_bn else None, residual=x)
        return ops.conv1d(t, k2, b2, act, pad, pad_value)

    def _boundary_layers(self):
        """Layer program of the boundary stack for ops.boundary_stack (flags: 1 save input, 2 add saved)."""
        key = ("boundary_stack",)
        if key not in self._tc:                  # cleared whenever weights are (re)loaded
            layers = []
            for k in range(self.n_boundary):
                kk, bb = self.conv("boundary/%d/conv" % k)
                layers.append({"kernel": kk, "bias": bb, "bn": self.bn("boundary/%d/bn" % k) if self.use_batchnorm else None, "flags": 0})
                name = "boundary/%d/resnet" % k
                for c, fl in enumerate((1, 2, 0)):
                    kc, bc_ = self.conv("%s/conv%d" % (name, c))
                    bn = self.bn("%s/bn%d" % (name, c)) if (self.use_batchnorm and c < 2) else None
                    layers.append({"kernel": kc, "bias": bc_, "bn": bn, "flags": fl})
            self._tc[key] = layers
        return self._tc[key]

    def _resnet2d(self, x, name, act):
        k0, b0 = self.conv(name + "/conv0")
        k1, b1 = self.conv(name + "/conv1")
        k2, b2 = self.conv(name + "/conv2")
        t = ops.conv2d(x, k0, b0, act, PAD_CONSTANT, 0.0)
        t = ops.conv2d(t, k1, b1, act, PAD_CONSTANT, 0.0, residual=x)
        return ops.conv2d(t, k2, b2, act, PAD_CONSTANT, 0.0)

    def _tc_supported(self):
        fin = self._cfg["final_convolutions_config"]
        return (all(k % 2 == 1 and k <= 15 for k in fin["kernel_sizes"]) and max(fin["filters"]) <= 32
                and self.x_dir_nmodes + 2 <= 32 and self.final_pad == PAD_CONSTANT and float(self.final_pad_value) == 0.0)

    def raw_forward(self, bc, dx, x_res):
        """Everything up to (not including) the final max-normalisation; returns (raw [B,1,x_res,n], max|raw| [B])."""
        if bc.dim() != 3 or bc.shape[1] != 1:
            raise ValueError("bc must be [batch, 1, n] (channels_first)")
        if dx.dim() != 2 or dx.shape[1] != 1 or dx.shape[0] != bc.shape[0]:
            raise ValueError("dx must be [batch, 1]")
        x_res = int(x_res)
        B, _, n = bc.shape
        h = ops.dbcnn_input(bc, x_res)
        stack = self._boundary_layers()
        if ops.boundary_stack_supported(n, stack):
            # the whole 1-D stack (n_boundary x (conv [+BN] + resnet)) in ONE kernel, activations in shared memory
            h = ops.boundary_stack(h, stack, self.boundary_act, self.boundary_pad, self.boundary_pad_value)
        else:
            for k in range(self.n_boundary):
                kk, bb = self.conv("boundary/%d/conv" % k)
                h = ops.conv1d(h, kk, bb, self.boundary_act, self.boundary_pad, self.boundary_pad_value,
                               bn=self.bn("boundary/%d/bn" % k) if self.use_batchnorm else None)
                h = self._resnet1d(h, "boundary/%d/resnet" % k, self.boundary_act, self.boundary_pad,
                                   self.boundary_pad_value, self.use_batchnorm)
        spp = ops.spatial_pyramid_pool(h, self.spp_levels, self.spp_mode, ndims=1)
        v = ops.dense_input(dx, x_res, n, extra=spp, normalize=True)
        for i in range(self.n_mlp):
            v = ops.dense(v, *self.conv("mlp/%d" % i), self.mlp_acts[i])
        S, nreg = self.n_final, self.final_regular_conv_stages
        if self.precision in ("tc", "tc2", "tc3"):
            if not self._tc_supported():
                raise NotImplementedError("precision='tc' covers odd kernels <= 15, <= 32 filters, zero CONSTANT padding")
            # Single-pass mode: the first 2-D convolution sees a SEPARABLE input (every channel of the mode expansion is
            # h[b,m,y] * S[m,x]; the position channels are 1 * posx[x] and posy[y] * 1), so its row taps fold into per-row
            # weights and the [B,29,x_res,n] expansion is never written (ops.pack_rowweights_tc): 14 MMAs per tile, not 121.
            sep = (self.tc_split == 1 and S - nreg >= 1 and os.environ.get("PCNN_DBCNN_SEPARABLE", "1") != "0")
            if sep:
                M = h.shape[1]
                key = ("rowweights", x_res)
                if key not in self._tc:      # row basis = (M sinh modes, posx, 1): built by the library's host-side packer
                    self._tc[key] = ops.pack_rowweights_tc(self._w["final/0/conv/kernel"], x_res=x_res)
                t = ops.dbcnn_signal_blk8(h, v)
            else:
                # the [B,29,x_res,n] mode expansion is produced directly in the tensor-core operand layout
                t = ops.dbcnn_expand_blk8(h, v, x_res, split=self.tc_split)
            for k in range(S - nreg):
                if sep and k == 0:
                    t = ops.conv2d_tc_rowweights(t, self._tc[("rowweights", x_res)], self._w.get("final/0/conv/bias"), self.final_act)
                else:
                    wp, bb = self.tc_conv("final/%d/conv" % k)
                    t = ops.conv2d_tc(t, wp, bb, self.final_act, PAD_CONSTANT)
                name = "final/%d/resnet" % k
                (w0, b0), (w1, b1), (w2, b2) = (self.tc_conv(name + "/conv%d" % i) for i in range(3))
                u = ops.conv2d_tc(t, w0, b0, self.final_act, PAD_CONSTANT)
                u = ops.conv2d_tc(u, w1, b1, self.final_act, PAD_CONSTANT, residual=t)
                t = ops.conv2d_tc(u, w2, b2, self.final_act, PAD_CONSTANT)
            for k in range(S - nreg, S):
                wp, bb = self.tc_conv("final/%d/conv" % k)
                t = ops.conv2d_tc(t, wp, bb, ACT_TANH, PAD_CONSTANT)
            out = ops.from_blk8(t, C=t.C)
            return out, ops.maxabs(out)
        out = ops.dbcnn_expand(h, v, x_res)
        for k in range(S - nreg):
            kk, bb = self.conv("final/%d/conv" % k)
            out = ops.conv2d(out, kk, bb, self.final_act, self.final_pad, self.final_pad_value)
            out = self._resnet2d(out, "final/%d/resnet" % k, self.final_act)
        for k in range(S - nreg, S):
            kk, bb = self.conv("final/%d/conv" % k)
            out = ops.conv2d(out, kk, bb, ACT_TANH, PAD_CONSTANT, 0.0)
        return out, ops.maxabs(out)

    def __call__(self, inp):
        bc, dx, x_res = inp
        if not isinstance(bc, torch.Tensor) or not bc.is_cuda:
            raise ValueError("bc must be a CUDA tensor (the hot path has no CPU implementation)")
        with torch.cuda.device(bc.device):      # launches go to the current device: make it the tensors' device
            if self.use_engine:
                if bc.dim() != 3 or bc.shape[1] != 1:
                    raise ValueError("bc must be [batch, 1, n] (channels_first)")
                if dx.dim() != 2 or dx.shape[1] != 1 or dx.shape[0] != bc.shape[0]:
                    raise ValueError("dx must be [batch, 1]")
                return self.engine().dbcnn_forward(bc, dx, x_res)
            raw, m = self.raw_forward(bc, dx, x_res)
            out = ops.dbcnn_finalize(raw, m, bc)
            if self.postsmoother_iterations > 0:
                out = ops.jacobi(out, torch.zeros_like(out), torch.cat([dx, dx], 1), self.postsmoother_iterations)
            return out

    call = __call__
