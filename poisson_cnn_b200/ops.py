"""Thin torch-tensor front end over the C ABI (include/pcnn.h).

torch is used only for device memory and streams; every function here enqueues hand-written CUDA
kernels from libpcnn.so on torch's current stream and returns without synchronising.
"""
import math

import numpy as np
import torch

from . import _lib
from ._lib import lib, check
from .config import (ACT_LINEAR, PAD_CONSTANT, PAD_SYMMETRIC, RESIZE_NEAREST, RESIZE_BILINEAR, RESIZE_BICUBIC,
                     RESIZE_BICUBIC_LEGACY_AC)

POOL_AVG, POOL_MAX = 0, 1
BC_DIRICHLET, BC_NEUMANN = 0, 1


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _chk(t, name, dtype=torch.float32):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (the hot path has no CPU implementation)" % name)
    if t.dtype != dtype:
        raise ValueError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if t.device.index != torch.cuda.current_device():
        # every launch goes to torch's CURRENT device and stream: a tensor living elsewhere would be dereferenced on the
        # wrong GPU.  The model classes switch to their own device; direct users of ops wrap calls in torch.cuda.device().
        raise ValueError("%s lives on %s but the current CUDA device is %d (wrap the call in torch.cuda.device(...))"
                         % (name, t.device, torch.cuda.current_device()))
    return t


def _nchw_bstride(t, name):
    """Accept dense NCHW tensors or channel-slices of one (batch stride may exceed C*H*W)."""
    _chk(t, name)
    B, C, H, W = t.shape
    sb, sc, sh, sw = t.stride()
    if not ((sw == 1 or W == 1) and (sh == W or H == 1) and (sc == H * W or C == 1)):
        raise ValueError("%s must be NCHW-contiguous within a sample (strides %s)" % (name, t.stride()))
    return sb if B > 1 else C * H * W


def conv2d(x, kernel, bias=None, act=ACT_LINEAR, pad_mode=PAD_CONSTANT, pad_value=0.0, bn=None,
           residual=None, out_scale=None, out=None):
    """pad + VALID conv + bias + act [+ BN affine] [+ residual] [* out_scale].  x [B,Cin,H,W],
    kernel [kh,kw,Cin,Cout] (Keras layout)."""
    in_bs = _nchw_bstride(x, "x")
    _chk(kernel, "kernel")
    B, Cin, H, W = x.shape
    kh, kw, kcin, Cout = kernel.shape
    if kcin != Cin:
        raise ValueError("conv2d: kernel expects %d input channels, got %d" % (kcin, Cin))
    if out is None:
        out = torch.empty((B, Cout, H, W), device=x.device, dtype=torch.float32)
    out_bs = _nchw_bstride(out, "out")
    res_bs = _nchw_bstride(residual, "residual") if residual is not None else 0
    bn_s, bn_t = (bn if bn is not None else (None, None))
    timed = KERNEL_TIMER is not None and KERNEL_TIMER.match(Cin, Cout, kh, kw, H, W)
    if timed:
        KERNEL_TIMER.start()
    check(lib.pcnn_conv2d_f32(_p(x), _p(kernel.contiguous()), _p(bias), _p(bn_s), _p(bn_t), _p(residual),
                              _p(out_scale), _p(out), B, Cin, Cout, H, W, kh, kw, int(pad_mode),
                              float(pad_value), int(act), in_bs, out_bs, res_bs, _stream()), "conv2d")
    if timed:
        KERNEL_TIMER.stop(2.0 * B * H * W * kh * kw * Cin * Cout)
    return out


class KernelTimer:
    """CUDA-event timer around the launches of one conv shape (bench.py's live roofline measurement).
    Events are recorded on the launching stream; durations are read after the timed region."""

    def __init__(self, cin, cout, k):
        self.key = (cin, cout, k)
        self.events, self.flops = [], []

    def match(self, cin, cout, kh, kw, H, W):
        return (cin, cout, kh) == self.key and kh == kw

    def start(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self._s = e

    def stop(self, flops):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.events.append((self._s, e))
        self.flops.append(flops)

    def summary(self):
        ms = [a.elapsed_time(b) for a, b in self.events]
        if not ms:
            return None
        return {"launches": len(ms), "avg_ms": sum(ms) / len(ms), "flops_per_launch": sum(self.flops) / len(self.flops)}


KERNEL_TIMER = None


def conv1d(x, kernel, bias=None, act=ACT_LINEAR, pad_mode=PAD_CONSTANT, pad_value=0.0, bn=None, residual=None):
    """Conv1D as the H == 1 case.  x [B,C,n], kernel [k,Cin,Cout]."""
    r = None if residual is None else residual.unsqueeze(2)
    return conv2d(x.unsqueeze(2), kernel.unsqueeze(0), bias, act, pad_mode, pad_value, bn, r).squeeze(2)


def boundary_stack_supported(n, layers):
    """The fused 1-D stack keeps three [28, n+18] activation buffers and one layer's weights in shared memory."""
    if not layers or len(layers) > 48:
        return False
    wmax = max(l["kernel"].shape[0] * l["kernel"].shape[1] * 32 for l in layers)
    ok = all(l["kernel"].shape[0] % 2 == 1 and l["kernel"].shape[0] <= 19 and max(l["kernel"].shape[1:]) <= 28 for l in layers)
    return ok and (3 * 28 * ((n + 18 + 3) // 4 * 4) + wmax) * 4 <= 220 * 1024


def boundary_stack(x, layers, act, pad_mode=PAD_CONSTANT, pad_value=0.0):
    """Fused Conv1D stack.  x [B,C,n]; layers: [{"kernel": [k,Cin,Cout], "bias": t|None, "bn": (scale, shift)|None,
    "flags": 0|1 (save input)|2 (add saved)}] -> [B, Cout_last, n]."""
    import ctypes
    _chk(x, "x")
    x = x.contiguous()
    B, C, n = x.shape
    L = len(layers)
    keep = [l["kernel"].contiguous() for l in layers]

    def arr_p(vals):
        return (ctypes.c_void_p * L)(*vals)

    def arr_i(vals):
        return (ctypes.c_int * L)(*[int(v) for v in vals])
    for l, kern in zip(layers, keep):
        _chk(kern, "kernel")
    out = torch.empty((B, keep[-1].shape[2], n), device=x.device, dtype=torch.float32)
    check(lib.pcnn_boundary_stack_f32(
        _p(x), _p(out), B, n, C, L, arr_p([k.data_ptr() for k in keep]), arr_p([_p(l.get("bias")) for l in layers]),
        arr_p([None if l.get("bn") is None else l["bn"][0].data_ptr() for l in layers]),
        arr_p([None if l.get("bn") is None else l["bn"][1].data_ptr() for l in layers]),
        arr_i([k.shape[0] for k in keep]), arr_i([k.shape[1] for k in keep]), arr_i([k.shape[2] for k in keep]),
        arr_i([l.get("flags", 0) for l in layers]), int(act), int(pad_mode), float(pad_value), _stream()), "boundary_stack")
    return out


def smallmap_stack_supported(H, W, layers):
    """The fused 2-D small-map stack: maps of <= 64 pixels, odd kernels <= 7, <= 32 channels, everything in shared memory."""
    if not layers or len(layers) > 24 or H * W > 64:
        return False
    ok = all(l["kernel"].shape[0] == l["kernel"].shape[1] and l["kernel"].shape[0] % 2 == 1 and l["kernel"].shape[0] <= 7
             and max(l["kernel"].shape[2:]) <= 32 for l in layers)
    if not ok:
        return False
    pm = max(l["kernel"].shape[0] // 2 for l in layers)
    wmax = max(l["kernel"].shape[0] ** 2 * l["kernel"].shape[2] * 32 for l in layers)
    return (3 * 32 * (H + 2 * pm) * (W + 2 * pm) + wmax) * 4 <= 220 * 1024


def smallmap_stack(x, layers, act, pad_mode=PAD_CONSTANT, pad_value=0.0):
    """Fused Conv2D stack on a tiny map.  x [B,C,H,W] (H*W <= 64); layers as in boundary_stack with Keras Conv2D kernels."""
    import ctypes
    _chk(x, "x")
    x = x.contiguous()
    B, C, H, W = x.shape
    L = len(layers)
    keep = [_chk(l["kernel"], "kernel").contiguous() for l in layers]

    def arr_p(vals):
        return (ctypes.c_void_p * L)(*vals)

    def arr_i(vals):
        return (ctypes.c_int * L)(*[int(v) for v in vals])
    out = torch.empty((B, keep[-1].shape[3], H, W), device=x.device, dtype=torch.float32)
    check(lib.pcnn_smallmap_stack_f32(
        _p(x), _p(out), B, H, W, C, L, arr_p([k.data_ptr() for k in keep]), arr_p([_p(l.get("bias")) for l in layers]),
        arr_p([None if l.get("bn") is None else l["bn"][0].data_ptr() for l in layers]),
        arr_p([None if l.get("bn") is None else l["bn"][1].data_ptr() for l in layers]),
        arr_i([k.shape[0] for k in keep]), arr_i([k.shape[2] for k in keep]), arr_i([k.shape[3] for k in keep]),
        arr_i([l.get("flags", 0) for l in layers]), int(act), int(pad_mode), float(pad_value), _stream()), "smallmap_stack")
    return out


def avgpool_same(x, s):
    in_bs = _nchw_bstride(x, "x")
    B, C, H, W = x.shape
    out = torch.empty((B, C, -(-H // s), -(-W // s)), device=x.device, dtype=torch.float32)
    check(lib.pcnn_avgpool_same_f32(_p(x), _p(out), B, C, H, W, int(s), in_bs, _stream()), "avgpool_same")
    return out


def deconv_same(x, kernel, bias, out_hw, stride, act=ACT_LINEAR, alpha=1.0, out=None, accumulate=False):
    _chk(x, "x"); _chk(kernel, "kernel")
    x = x.contiguous()
    B, Cin, ih, iw = x.shape
    kh, kw, Cout, kcin = kernel.shape
    if kcin != Cin:
        raise ValueError("deconv_same: kernel expects %d input channels, got %d" % (kcin, Cin))
    oh, ow = int(out_hw[0]), int(out_hw[1])
    if out is None:
        out = torch.empty((B, Cout, oh, ow), device=x.device, dtype=torch.float32)
        accumulate = False
    out_bs = _nchw_bstride(out, "out")
    check(lib.pcnn_deconv_same_f32(_p(x), _p(kernel.contiguous()), _p(bias), _p(out), B, Cin, Cout, ih, iw, oh, ow,
                                   kh, kw, int(stride), int(act), float(alpha), int(bool(accumulate)), out_bs,
                                   _stream()), "deconv_same")
    return out


# ---- host-built tables (tiny, cached per device) -------------------------------------------------
_cache = {}


def _cached(key, build):
    if key not in _cache:
        _cache[key] = build()
    return _cache[key]


def _bicubic_table(a=-0.5):
    # TF resize_bicubic CPU kernel: 1024-step cubic-convolution table in float32 (a = -0.5 Keys with half-pixel
    # centres, a = -0.75 for the legacy align_corners path)
    n = 1024
    a = np.float32(a)
    x = np.arange(n + 1, dtype=np.float32) / np.float32(n)
    t0 = ((a + np.float32(2)) * x - (a + np.float32(3))) * x * x + np.float32(1)
    x1 = x + np.float32(1)
    t1 = ((a * x1 - np.float32(5) * a) * x1 + np.float32(8) * a) * x1 - np.float32(4) * a
    tab = np.empty(2 * (n + 1), dtype=np.float32)
    tab[0::2], tab[1::2] = t0, t1
    return tab


def resize_axis_table(n_in, n_out, method):
    """Per-axis gather indices / weights of tf.image.resize (half-pixel centres, antialias=False), or of the
    legacy tf.compat.v1 resize_images(BICUBIC, align_corners=True) used by dataset/utils/image_resize.py:20."""
    o = np.arange(n_out, dtype=np.float32)
    if method == RESIZE_BICUBIC_LEGACY_AC:
        # scale = (in-1)/(out-1), src = dst*scale, a = -0.75, out-of-range taps CLAMP to the edge (no renormalisation)
        scale = np.float32(n_in - 1) / np.float32(n_out - 1) if n_out > 1 else np.float32(n_in) / np.float32(n_out)
        src = o * scale
        fl = np.floor(src)
        tab = _cached("bicubic_tab_legacy", lambda: _bicubic_table(-0.75))
        n = 1024
        loc = fl.astype(np.int64)
        off = np.rint((src - fl) * np.float32(n)).astype(np.int64)
        w = np.stack([tab[off * 2 + 1], tab[off * 2], tab[(n - off) * 2], tab[(n - off) * 2 + 1]], 1).astype(np.float32)
        idx = np.clip(np.stack([loc - 1, loc, loc + 1, loc + 2], 1), 0, n_in - 1)
        return idx.astype(np.int32), w
    scale = np.float32(n_in) / np.float32(n_out)
    if method == RESIZE_NEAREST:
        idx = np.clip(np.floor((o + np.float32(0.5)) * scale).astype(np.int64), 0, n_in - 1)
        return idx[:, None].astype(np.int32), np.ones((n_out, 1), np.float32)
    src = (o + np.float32(0.5)) * scale - np.float32(0.5)
    fl = np.floor(src)
    if method == RESIZE_BILINEAR:
        lo = np.maximum(fl.astype(np.int64), 0)
        hi = np.minimum(np.ceil(src).astype(np.int64), n_in - 1)
        lerp = (src - fl).astype(np.float32)
        return np.stack([lo, hi], 1).astype(np.int32), np.stack([np.float32(1) - lerp, lerp], 1).astype(np.float32)
    if method == RESIZE_BICUBIC:
        tab = _cached("bicubic_tab", _bicubic_table)
        n = 1024
        loc = fl.astype(np.int64)
        off = np.rint((src - fl) * np.float32(n)).astype(np.int64)
        w = np.stack([tab[off * 2 + 1], tab[off * 2], tab[(n - off) * 2], tab[(n - off) * 2 + 1]], 1)
        raw = np.stack([loc - 1, loc, loc + 1, loc + 2], 1)
        idx = np.clip(raw, 0, n_in - 1)
        w = np.where(idx == raw, w, np.float32(0)).astype(np.float32)
        s = w.sum(1, dtype=np.float32)
        ok = np.abs(s) >= 1000.0 * np.finfo(np.float32).tiny
        w = np.where(ok[:, None], w * (np.float32(1) / np.where(ok, s, np.float32(1)))[:, None], w).astype(np.float32)
        return idx.astype(np.int32), w
    raise ValueError("resize method %r" % (method,))


def _resize_tables(device, n_in, n_out, method):
    def build():
        idx, w = resize_axis_table(n_in, n_out, method)
        return (torch.from_numpy(np.ascontiguousarray(idx)).to(device), torch.from_numpy(np.ascontiguousarray(w)).to(device))
    return _cached(("resize", str(device), n_in, n_out, method), build)


def resize(x, out_hw, method, alpha=1.0, out=None, accumulate=False):
    _chk(x, "x")
    x = x.contiguous()
    B, C, ih, iw = x.shape
    oh, ow = int(out_hw[0]), int(out_hw[1])
    iy, wy = _resize_tables(x.device, ih, oh, method)
    ix, wx = _resize_tables(x.device, iw, ow, method)
    if out is None:
        out = torch.empty((B, C, oh, ow), device=x.device, dtype=torch.float32)
        accumulate = False
    out_bs = _nchw_bstride(out, "out")
    check(lib.pcnn_resize_f32(_p(x), _p(iy), _p(wy), _p(ix), _p(wx), iy.shape[1], _p(out), B, C, ih, iw, oh, ow,
                              float(alpha), int(bool(accumulate)), out_bs, _stream()), "resize")
    return out


def split_indices(n, sections):
    """poisson_CNN/dataset/utils/split_indices.py:4-26 (numpy.array_split boundaries)."""
    per, extra = n // sections, n % sections
    return np.cumsum([0] + [per + 1] * extra + [per] * (sections - extra))


def spp_boxes(levels, H, W, ndims):
    """Bin boxes (y0,y1,x0,x1) in the order SpatialPyramidPool.call emits them (meshgrid 'ij')."""
    boxes = []
    for lv in levels:
        if isinstance(lv, int):
            lv = [lv] * ndims
        elif len(lv) == 1:
            lv = [lv[0]] * ndims
        elif len(lv) != ndims:
            raise ValueError("Each SPP level must have a pool size with ndims or 1 element(s). Got " + str(len(lv)))
        if ndims == 1:
            ex = split_indices(W, lv[0])
            boxes += [(0, H, int(ex[i]), int(ex[i + 1])) for i in range(lv[0])]
        else:
            ey, ex = split_indices(H, lv[0]), split_indices(W, lv[1])
            boxes += [(int(ey[i]), int(ey[i + 1]), int(ex[j]), int(ex[j + 1])) for i in range(lv[0]) for j in range(lv[1])]
    return np.asarray(boxes, dtype=np.int32)


def spatial_pyramid_pool(x, levels, mode, ndims=2):
    """x [B,C,H,W] (ndims=2) or [B,C,n] (ndims=1) -> [B, sum(bins)]."""
    _chk(x, "x")
    if ndims == 1:
        x = x.unsqueeze(2)
    x = x.contiguous()
    B, C, H, W = x.shape
    key = ("spp", str(x.device), repr(levels), H, W, ndims)
    boxes = _cached(key, lambda: torch.from_numpy(spp_boxes(levels, H, W, ndims)).to(x.device))
    out = torch.empty((B, boxes.shape[0]), device=x.device, dtype=torch.float32)
    check(lib.pcnn_spp_f32(_p(x), _p(boxes), _p(out), B, C, H, W, boxes.shape[0], int(mode), _stream()), "spp")
    return out


def dense(x, kernel, bias, act=ACT_LINEAR):
    _chk(x, "x"); _chk(kernel, "kernel")
    x = x.contiguous()
    B, nin = x.shape
    if kernel.shape[0] != nin:
        raise ValueError("dense: kernel expects %d inputs, got %d" % (kernel.shape[0], nin))
    y = torch.empty((B, kernel.shape[1]), device=x.device, dtype=torch.float32)
    check(lib.pcnn_dense_f32(_p(x), _p(kernel.contiguous()), _p(bias), _p(y), B, nin, kernel.shape[1], int(act), _stream()), "dense")
    return y


def maxabs(x):
    """max |x| per sample -> [B]."""
    _chk(x, "x")
    x = x.contiguous()
    B = x.shape[0]
    out = torch.empty((B,), device=x.device, dtype=torch.float32)
    check(lib.pcnn_maxabs_f32(_p(x), _p(out), B, x.numel() // B, _stream()), "maxabs")
    return out


def scale_inv(x, m, out=None):
    """x * (1/m[b]) per sample (set_max_magnitude with max_magnitude 1.0)."""
    _chk(x, "x"); _chk(m, "m")
    x = x.contiguous()
    y = torch.empty_like(x) if out is None else out
    if not y.is_contiguous() or y.shape != x.shape:
        raise ValueError("scale_inv: out must be contiguous and shaped like x")
    B = x.shape[0]
    check(lib.pcnn_scale_inv_f32(_p(x), _p(m), _p(y), B, x.numel() // B, _stream()), "scale_inv")
    return y


def tf_linspace01(n):
    """tf.linspace(0., 1., n) in float32: start + i*step, last element := stop."""
    if n == 1:
        return np.zeros(1, np.float32)
    step = np.float32(1.0) / np.float32(n - 1)
    v = np.arange(n, dtype=np.float32) * step
    v[-1] = np.float32(1.0)
    return v


def host_table(what, count, a, b=0, c=0, dtype=np.float32):
    """A host-built table from the library's own builders (csrc/engine.cu: the model-level C ABI uploads the same bytes)."""
    import ctypes
    out = np.zeros(int(count), dtype=dtype)
    n = lib.pcnn_host_table(what.encode(), int(a), int(b), int(c), out.ctypes.data_as(ctypes.c_void_p), out.nbytes)
    if n != count:
        check(int(n) if n < 0 else -1, "pcnn_host_table(%s)" % what)
    return out


def position_table(device, n):
    """cos(pi * linspace(0,1,n)), float32 (generate_position_embeddings)."""
    return _cached(("pos", str(device), n), lambda: torch.from_numpy(host_table("pos", n, n)).to(device))


def sinh_basis_table(device, n_modes, x_res):
    """build_series_x_dir_components (poisson_CNN/models/Dirichlet_BC_NN_Legacy.py:106-112), float32."""
    return _cached(("sinh", str(device), n_modes, x_res),
                   lambda: torch.from_numpy(host_table("sinh", n_modes * x_res, n_modes, x_res).reshape(n_modes, x_res)).to(device))


def hpnn_input(rhs):
    _chk(rhs, "rhs")
    rhs = rhs.contiguous()
    B, _, H, W = rhs.shape
    out = torch.empty((B, 3, H, W), device=rhs.device, dtype=torch.float32)
    check(lib.pcnn_hpnn_input_f32(_p(rhs), _p(position_table(rhs.device, H)), _p(position_table(rhs.device, W)),
                                  _p(out), B, H, W, _stream()), "hpnn_input")
    return out


def dbcnn_input(bc, x_res):
    _chk(bc, "bc")
    bc = bc.contiguous()
    B, _, n = bc.shape
    out = torch.empty((B, 3, n), device=bc.device, dtype=torch.float32)
    posx0 = float(np.cos(np.float32(math.pi) * tf_linspace01(x_res)[:1])[0])
    check(lib.pcnn_dbcnn_input_f32(_p(bc), posx0, _p(position_table(bc.device, n)), _p(out), B, n, _stream()), "dbcnn_input")
    return out


def dbcnn_expand(h, modew, x_res):
    _chk(h, "h"); _chk(modew, "modew")
    h, modew = h.contiguous(), modew.contiguous()
    B, M, n = h.shape
    S = sinh_basis_table(h.device, M, x_res)
    out = torch.empty((B, M + 2, x_res, n), device=h.device, dtype=torch.float32)
    check(lib.pcnn_dbcnn_expand_f32(_p(h), _p(S), _p(modew), _p(position_table(h.device, x_res)),
                                    _p(position_table(h.device, n)), _p(out), B, M, x_res, n, _stream()), "dbcnn_expand")
    return out


def dbcnn_finalize(raw, m, bc):
    raw, bc = raw.contiguous(), bc.contiguous()
    B, _, xres, n = raw.shape
    out = torch.empty_like(raw)
    check(lib.pcnn_dbcnn_finalize_f32(_p(raw), _p(m), _p(bc), _p(out), B, xres, n, _stream()), "dbcnn_finalize")
    return out


def hpnn_finalize(y, s, bc_type):
    """y [B,1,H,W] (may be channel 0 of a wider buffer); s [B] or None."""
    ybs = _nchw_bstride(y, "y")
    B, _, H, W = y.shape
    out = torch.empty((B, 1, H, W), device=y.device, dtype=torch.float32)
    check(lib.pcnn_hpnn_finalize_f32(_p(y), _p(s), _p(out), B, H, W, int(bc_type), ybs, _stream()), "hpnn_finalize")
    return out


def dense_input(dx, n0, n1, extra=None, normalize=False):
    """[dx, dx*(n0-1), dx*(n1-1), extra...] -> [B, 3+nextra]; normalize divides the sizes by their max."""
    _chk(dx, "dx")
    dx = dx.contiguous()
    B = dx.shape[0]
    ne = 0 if extra is None else extra.shape[1]
    if extra is not None:
        extra = _chk(extra, "extra").contiguous()
    out = torch.empty((B, 3 + ne), device=dx.device, dtype=torch.float32)
    check(lib.pcnn_dense_input_f32(_p(dx), _p(extra), _p(out), B, int(n0), int(n1), ne, int(bool(normalize)), _stream()), "dense_input")
    return out


def merge(hp, L, T, R, Bt, dx, mrhs, ml, mt, mr, mb):
    B, _, nx, ny = hp.shape
    out = torch.empty_like(hp)
    args = [t.contiguous() for t in (hp, L, T, R, Bt, dx, mrhs, ml, mt, mr, mb)]
    check(lib.pcnn_merge_f32(*[_p(t) for t in args], _p(out), B, nx, ny, _stream()), "merge")
    return out


def laplacian_residual(rhs, sol, grid_spacings, stencil=3, rhs_maxabs=None):
    """Per-sample sum of squared residuals (float64) [B]."""
    _chk(rhs, "rhs"); _chk(sol, "sol"); _chk(grid_spacings, "grid_spacings")
    rhs, sol, gs = rhs.contiguous(), sol.contiguous(), grid_spacings.contiguous()
    if rhs.shape != sol.shape or rhs.shape[1] != 1:
        raise ValueError("laplacian_residual: rhs and solution must both be [B,1,H,W]")
    B, _, H, W = rhs.shape
    out = torch.empty((B,), device=rhs.device, dtype=torch.float64)
    check(lib.pcnn_laplacian_residual_f32(_p(rhs), _p(sol), _p(gs), _p(rhs_maxabs), _p(out), B, H, W, int(stencil),
                                          _stream()), "laplacian_residual")
    return out


def jacobi(guess, rhs, grid_spacings, n_iter):
    cur = guess.contiguous()
    rhs, gs = rhs.contiguous(), grid_spacings.contiguous()
    B, _, H, W = cur.shape
    for _ in range(n_iter):
        nxt = torch.empty_like(cur)
        check(lib.pcnn_jacobi_sweep_f32(_p(cur), _p(rhs), _p(gs), _p(nxt), B, H, W, _stream()), "jacobi_sweep")
        cur = nxt
    return cur


def dst_solve(rhs, left, top, right, bottom, dx, method="fft", dtype=torch.float64):
    """DST-I direct solve of the reference's ground-truth system.  rhs [B,1,nx,ny] -> [B,1,nx,ny].
    method "fft" (default): DST-I along y (Bluestein chirp-z transforms on shared-memory FFTs, arithmetic in `dtype`: float64
    like the reference's solver, or float32), float64 tridiagonal solves along x, inverse DST along y: three passes over the
    grid (csrc/dst_fft.cu); "gemm": the dense sine-matrix form in float64 (O(N^3), kept as an independent cross-check and
    for rows beyond 2050 points)."""
    for t, nme in ((rhs, "rhs"), (left, "left"), (top, "top"), (right, "right"), (bottom, "bottom"), (dx, "dx")):
        _chk(t, nme)
    if method not in ("fft", "gemm"):
        raise ValueError("dst_solve: method must be 'fft' or 'gemm'")
    B, _, nx, ny = rhs.shape
    dev = rhs.device
    args = [t.contiguous() for t in (rhs, left, top, right, bottom, dx)]
    out = torch.empty_like(args[0])
    if method == "fft" and ny <= 2050:
        dbl = 1 if dtype == torch.float64 else 0
        if dtype not in (torch.float64, torch.float32):
            raise ValueError("dst_solve: dtype must be float64 or float32")

        def build():
            p = torch.empty((lib.pcnn_dst_fft_plan_bytes(ny - 2, dbl),), device=dev, dtype=torch.uint8)
            check(lib.pcnn_dst_fft_plan_init(_p(p), ny - 2, dbl, _stream()), "dst_fft_plan_init")
            return p
        py = _cached(("dst_plan", str(dev), ny - 2, dbl), build)
        work = torch.empty((lib.pcnn_dst_fft_workspace_bytes(B, nx, ny),), device=dev, dtype=torch.uint8)
        check(lib.pcnn_dst_solve_fft(*[_p(t) for t in args], _p(py), _p(work), _p(out), B, nx, ny, dbl, _stream()),
              "dst_solve_fft")
        return out

    def sine(m):
        def build():
            s = torch.empty((m, m), device=dev, dtype=torch.float64)
            check(lib.pcnn_dst_sine_matrix(_p(s), m, _stream()), "dst_sine_matrix")
            return s
        return _cached(("dst_sine", str(dev), m), build)
    sx, sy = sine(nx - 2), sine(ny - 2)
    work = torch.empty((lib.pcnn_dst_workspace_bytes(B, nx, ny) // 8,), device=dev, dtype=torch.float64)
    check(lib.pcnn_dst_solve(*[_p(t) for t in args], _p(sx), _p(sy), _p(work), _p(out), B, nx, ny, _stream()), "dst_solve")
    return out


# =================================================================================================
# tensor-core path: BLK8 fp16 activations + tcgen05 convolution (csrc/conv_tc.cu)
# =================================================================================================
_BLK8_POOL = {}     # (device, B, C, H, W) -> [buffer]: recycled buffers keep their zero halo / channel padding


def _blk8_buffer(key, nbytes, device):
    """(buffer, halo state).  Recycled buffers come back with whatever their halo held when they were released
    (zeros from allocation, or a mirrored ring): the consumer-side check in blk8_halo_fill refreshes it only if
    the next user needs something else, so steady-state inference neither allocates nor refills."""
    pool = _BLK8_POOL.get(key)
    if pool:
        return pool.pop()
    try:
        buf = torch.zeros(nbytes // 2, dtype=torch.float16, device=device)
    except torch.OutOfMemoryError:
        # the pool keeps one set of buffers per tensor shape ever seen; under memory pressure (many grid
        # shapes in one process) the idle ones are released and the allocation is retried once
        blk8_pool_clear()
        torch.cuda.empty_cache()
        buf = torch.zeros(nbytes // 2, dtype=torch.float16, device=device)
    return buf, (PAD_CONSTANT, 7)


class Blk8:
    """fp16 activation in the BLK8 layout [B][Cpad/8][H+14][W+14][8] (7-pixel halo in memory).
    split selects the precision mode of the tensor: False/1 fp16 only; True/2 a second buffer `lo` with the
    fp16 rounding remainder (x = hi + lo, ~22 bits); 3 a second buffer of e4m3 planes (e4m3(x) and
    e4m3((x-hi)*2^11), 16 channels per plane) feeding the single fp8 correction MMA.
    `halo` records what the halo currently holds: (PAD_CONSTANT, 7) after allocation, (mode, pad) after a
    halo fill or a producer that wrote the mirrored ring itself, (mode, -1) when a mirrored halo went stale.
    Buffers are recycled (with their halo state) through a pool keyed by the exact shape, so steady-state
    inference neither allocates nor refills them."""

    __slots__ = ("buf", "lo", "mode", "B", "C", "H", "W", "halo", "_key", "_key_lo", "_pool")

    def __init__(self, B, C, H, W, device, split=False, sym=False):
        nbytes = lib.pcnn_blk8_bytes(B, C, H, W)
        if nbytes == 0:
            raise ValueError("Blk8: bad shape")
        self.mode = (2 if split else 1) if isinstance(split, bool) else int(split)
        if self.mode not in (1, 2, 3):
            raise ValueError("Blk8: precision mode must be 1, 2 or 3")
        # hi and lo/q buffers never trade places: bytes one role leaves untouched (channel padding) must stay zero
        # ... and tensors whose producer writes a mirrored ring recycle among themselves (no refills in steady state)
        self._key = (str(device), B, C, H, W, "hi", bool(sym))
        self._key_lo = (str(device), B, C, H, W, "lo%d" % self.mode, bool(sym))
        self._pool = _BLK8_POOL          # buffers go back to the pool they came from (a CUDA graph owns a private one)
        self.buf, st = _blk8_buffer(self._key, nbytes, device)
        self.lo = None
        if self.mode >= 2:
            self.lo, st_lo = _blk8_buffer(self._key_lo, nbytes, device)
            if st_lo != st:
                st = (st[0] if st[0] != PAD_CONSTANT else st_lo[0], -1)      # the two buffers disagree: stale
        self.B, self.C, self.H, self.W = B, C, H, W
        self.halo = st

    def __del__(self):
        try:
            self._pool.setdefault(self._key, []).append((self.buf, self.halo))
            if self.lo is not None:
                self._pool.setdefault(self._key_lo, []).append((self.lo, self.halo))
        except Exception:
            pass

    @property
    def device(self):
        return self.buf.device

    @property
    def split(self):
        return self.lo is not None

    def _plane_off(self, c_offset):
        if c_offset % 8:
            raise ValueError("channel offset must be a multiple of 8")
        return (c_offset // 8) * (self.H + 14) * (self.W + 14) * 16

    def plane_ptr(self, c_offset):
        """Device address of channel plane c_offset/8 of sample 0 (in-place concat)."""
        return self.buf.data_ptr() + self._plane_off(c_offset)

    def plane_ptr_lo(self, c_offset):
        return None if self.lo is None else self.lo.data_ptr() + self._plane_off(c_offset)


def blk8_pool_clear():
    _BLK8_POOL.clear()


class blk8_pool_scope:
    """Context manager: BLK8 buffers are drawn from (and return to) `pool` instead of the process-wide one.
    graph.GraphedCall gives every captured CUDA graph its own pool: the buffers a graph was captured with, and the
    halo contents it leaves in them, must never be seen by eager calls or by other graphs."""

    def __init__(self, pool):
        self.pool = pool

    def __enter__(self):
        global _BLK8_POOL
        self._saved = _BLK8_POOL
        _BLK8_POOL = self.pool
        return self.pool

    def __exit__(self, *exc):
        global _BLK8_POOL
        _BLK8_POOL = self._saved
        return False


def blk8_pool_canonicalize(pool=None):
    """Sort every free list by buffer address: the next pass then hands the same buffer to the same tensor as the
    previous canonicalised pass did, whatever order the buffers came back in."""
    pool = _BLK8_POOL if pool is None else pool
    for v in pool.values():
        v.sort(key=lambda e: e[0].data_ptr())


def blk8_pool_snapshot(pool=None):
    """{key: [(address, halo state), ...]}: what the next pass will find (used to detect the steady state)."""
    pool = _BLK8_POOL if pool is None else pool
    return {k: [(b.data_ptr(), st) for b, st in v] for k, v in pool.items() if v}


def _fusable_halo(mode, H, W):
    """SYMMETRIC rings can be written by the producer itself when the map is at least 7x7."""
    return int(mode) == PAD_SYMMETRIC and H >= 7 and W >= 7


def to_blk8(x, out=None, c_total=None, c_offset=0, split=False, halo=PAD_CONSTANT):
    """NCHW fp32 -> BLK8 (optionally into channels [c_offset, c_offset+C) of a wider buffer).
    halo=PAD_SYMMETRIC also writes the mirrored 7-wide ring (the padding of the layer that reads the tensor)."""
    in_bs = _nchw_bstride(x, "x")
    B, C, H, W = x.shape
    if out is None:
        out = Blk8(B, c_total or C, H, W, x.device, split=split, sym=_fusable_halo(halo, H, W))
    if (out.B, out.H, out.W) != (B, H, W):
        raise ValueError("to_blk8: destination shape mismatch")
    fused = _fusable_halo(halo, H, W) and (out.C + 15) // 16 == (C + 15) // 16 and c_offset == 0
    check(lib.pcnn_to_blk8(_p(x), _p(out.buf), _p(out.lo), out.mode, B, C, H, W, out.C, int(c_offset), in_bs,
                           PAD_SYMMETRIC if fused else PAD_CONSTANT, _stream()), "to_blk8")
    if fused:
        out.halo = (PAD_SYMMETRIC, 7)
    elif out.halo[0] != PAD_CONSTANT:
        out.halo = (out.halo[0], -1)
    return out


def from_blk8(t, C=None, c_offset=0, out=None):
    C = t.C if C is None else C
    if out is None:
        out = torch.empty((t.B, C, t.H, t.W), device=t.device, dtype=torch.float32)
    out_bs = _nchw_bstride(out, "out")
    check(lib.pcnn_from_blk8(_p(t.buf), _p(t.lo), t.mode, _p(out), t.B, C, t.H, t.W, t.C, int(c_offset), out_bs, _stream()), "from_blk8")
    return out


def blk8_halo_fill(t, pad, mode):
    """Make the halo of `t` hold tf.pad(mode) of width pad (no-op if it already does)."""
    pad, mode = int(pad), int(mode)
    cur_mode, cur_pad = t.halo
    if mode == PAD_CONSTANT and cur_mode == PAD_CONSTANT:
        return t                      # zero halo is 7 wide from allocation and never written
    if cur_mode == mode and cur_pad >= pad:
        return t                      # a wider mirrored ring contains the narrower one
    if mode == PAD_CONSTANT:
        pad = 7
    for b in (t.buf, t.lo):
        if b is not None:
            check(lib.pcnn_blk8_halo_fill(_p(b), t.B, t.C, t.H, t.W, pad, mode, _stream()), "blk8_halo_fill")
    t.halo = (mode, pad)
    return t


def pack_conv_weights_tc(kernel, nsplit=1):
    """Keras [k,k,Cin,Cout] fp32 -> packed fp16 operand image (done once per layer); nsplit=2 adds W_lo."""
    _chk(kernel, "kernel")
    kh, kw, Cin, Cout = kernel.shape
    n = lib.pcnn_conv_tc_packed_weight_bytes(kh, kw, Cin, Cout, int(nsplit))
    packed = torch.empty(n // 2, dtype=torch.float16, device=kernel.device)
    # power-of-two pre-scale so that max|W| lands near 2^9: W_hi and the remainder W_lo (~2^-12 |W|) both stay in
    # fp16's normal range; the kernel multiplies the accumulator by 1/scale (exact).  One host sync at pack time.
    wmax = float(kernel.abs().max())
    scale = 2.0 ** max(-24, min(24, math.floor(math.log2(512.0 / wmax)))) if wmax > 0 and math.isfinite(wmax) else 1.0
    check(lib.pcnn_conv_tc_pack_weights(_p(kernel.contiguous()), _p(packed), kh, kw, Cin, Cout, int(nsplit), scale, _stream()), "conv_tc_pack_weights")
    return {"packed": packed, "k": kh, "cin": Cin, "cout": Cout, "nsplit": int(nsplit), "acc_scale": 1.0 / scale}


_NUM_SMS = {}


def _num_sms(device):
    if device not in _NUM_SMS:
        _NUM_SMS[device] = torch.cuda.get_device_properties(device).multi_processor_count
    return _NUM_SMS[device]


def conv2d_tc(x, wp, bias=None, act=ACT_LINEAR, pad_mode=PAD_CONSTANT, bn=None, residual=None, out_scale=None,
              out=None, out_channels_total=None, out_c_offset=0, out_halo=PAD_CONSTANT, correction=True):
    """tcgen05 convolution on BLK8 tensors.  x: Blk8 with >= wp['cin'] channels; returns a Blk8.
    Split precision is selected by the packed weights (wp['nsplit'] == 2 needs split tensors)."""
    if not isinstance(x, Blk8):
        raise ValueError("conv2d_tc: x must be a Blk8 tensor (use to_blk8)")
    k, cout, nsplit = wp["k"], wp["cout"], wp["nsplit"]
    if -(-wp["cin"] // 16) != -(-x.C // 16):
        raise ValueError("conv2d_tc: kernel expects %d input channels, tensor holds %d" % (wp["cin"], x.C))
    split = nsplit >= 2
    if x.mode != nsplit:
        raise ValueError("conv2d_tc: weights packed for precision mode %d, input tensor is mode %d" % (nsplit, x.mode))
    if nsplit == 3 and -(-wp["cin"] // 8) != -(-x.C // 8) and (-(-wp["cin"] // 8) % 2 or -(-x.C // 8) % 2):
        # a lone last 8-channel plane changes the e4m3 plane layout of the tensor and the packing of the kernel
        raise ValueError("conv2d_tc: mode-3 tensor with %d channels does not match a kernel packed for %d" % (x.C, wp["cin"]))
    blk8_halo_fill(x, k // 2, pad_mode)
    if out is None:
        out = Blk8(x.B, out_channels_total or cout, x.H, x.W, x.device, split=nsplit, sym=_fusable_halo(out_halo, x.H, x.W))
    if (out.B, out.H, out.W) != (x.B, x.H, x.W) or out.mode != nsplit:
        raise ValueError("conv2d_tc: destination shape / precision mismatch")
    if residual is not None and ((residual.B, residual.H, residual.W) != (x.B, x.H, x.W) or residual.mode != nsplit):
        raise ValueError("conv2d_tc: residual shape / precision mismatch")
    bn_s, bn_t = (bn if bn is not None else (None, None))
    # the producer writes the next layer's SYMMETRIC padding itself when it owns every channel of the tensor
    fused_halo = _fusable_halo(out_halo, x.H, x.W) and out_c_offset == 0 and (out.C + 15) // 16 == (cout + 15) // 16
    timed = KERNEL_TIMER is not None and KERNEL_TIMER.match(wp["cin"], cout, k, k, x.H, x.W)
    if timed:
        KERNEL_TIMER.start()
    check(lib.pcnn_conv2d_tc(_p(x.buf), _p(x.lo) if split else None, _p(wp["packed"]), _p(bias), _p(bn_s), _p(bn_t),
                             None if residual is None else _p(residual.buf),
                             None if (residual is None or not split) else _p(residual.lo), _p(out_scale),
                             out.plane_ptr(out_c_offset), out.plane_ptr_lo(out_c_offset) if split else None,
                             x.B, wp["cin"], cout, out.C, 0 if residual is None else residual.C, x.H, x.W, k, int(act),
                             nsplit | (0x10 if (nsplit == 3 and not correction) else 0), wp["acc_scale"],
                             PAD_SYMMETRIC if fused_halo else PAD_CONSTANT,
                             _num_sms(x.device), _stream()), "conv2d_tc")
    if timed:
        KERNEL_TIMER.stop(2.0 * x.B * x.H * x.W * k * k * wp["cin"] * cout)
    if fused_halo:
        out.halo = (PAD_SYMMETRIC, 7)         # the epilogue wrote the mirrored ring for the next layer
    elif out.halo[0] != PAD_CONSTANT:
        out.halo = (out.halo[0], -1)          # interior changed: a mirrored halo is stale
    return out


def dbcnn_expand_blk8(h, modew, x_res, split=False):
    """einsum('bmy,mx,bm->bmxy') + concat(pos) written straight into a BLK8 tensor [B, M+2, x_res, n]."""
    _chk(h, "h"); _chk(modew, "modew")
    h, modew = h.contiguous(), modew.contiguous()
    B, M, n = h.shape
    S = sinh_basis_table(h.device, M, x_res)
    out = Blk8(B, M + 2, x_res, n, h.device, split=split)
    check(lib.pcnn_dbcnn_expand_blk8(_p(h), _p(S), _p(modew), _p(position_table(h.device, x_res)),
                                     _p(position_table(h.device, n)), _p(out.buf), _p(out.lo), out.mode, B, M, x_res, n, _stream()), "dbcnn_expand_blk8")
    return out


def pack_deconv_kernel(kernel):
    """Keras deconv kernel [s,s,C,C] (k == stride) -> the bank-staggered phase-matrix image the fused
    upsample-merge kernel bulk-copies ([s][s][C/8][8*C+4] floats); done once per layer."""
    _chk(kernel, "kernel")
    s_, s2, C, C2 = kernel.shape
    n = lib.pcnn_upsample_merge_packed_floats(int(s_), int(C)) if (s_ == s2 and C == C2) else 0
    if n == 0:
        raise ValueError("pack_deconv_kernel: needs a [s,s,C,C] kernel with C a multiple of 8")
    packed = torch.empty(n, dtype=torch.float32, device=kernel.device)
    check(lib.pcnn_upsample_merge_pack_kernel(_p(kernel.contiguous()), _p(packed), int(s_), int(C), _stream()), "upsample_merge_pack_kernel")
    return packed


def upsample_merge_blk8(deconv_branches, resize_branches, alpha, out, c_offset, H, W):
    """Fused upsample + branch sum written into channels [c_offset, c_offset+C) of the Blk8 tensor `out`.
    deconv_branches: [(x [B,C,ih,iw] fp32, packed kernel (pack_deconv_kernel), bias or None, stride, act)];
    resize_branches: [(x [B,C,ih,iw] fp32, method)]."""
    import ctypes
    if not isinstance(out, Blk8) or (out.H, out.W) != (int(H), int(W)):
        raise ValueError("upsample_merge_blk8: destination must be a Blk8 tensor of the output size")
    B, C = out.B, None
    keep = []

    def arr_p(vals):
        return (ctypes.c_void_p * max(len(vals), 1))(*[v for v in vals])

    def arr_i(vals):
        return (ctypes.c_int * max(len(vals), 1))(*[int(v) for v in vals])
    d_in, d_k, d_b, d_s, d_ih, d_iw, d_act = [], [], [], [], [], [], []
    for x, kern, bias, stride, act in deconv_branches:
        _chk(x, "x"); _chk(kern, "kernel")
        x = x.contiguous()
        C = x.shape[1] if C is None else C
        if x.shape[0] != B or x.shape[1] != C or kern.numel() != lib.pcnn_upsample_merge_packed_floats(int(stride), int(C)):
            raise ValueError("upsample_merge_blk8: deconv branch needs x [B,C,ih,iw] and pack_deconv_kernel() of a [s,s,C,C] kernel with s == stride")
        keep += [x, kern]
        d_in.append(x.data_ptr()); d_k.append(kern.data_ptr()); d_b.append(_p(bias)); d_s.append(stride)
        d_ih.append(x.shape[2]); d_iw.append(x.shape[3]); d_act.append(act)
    r_in, r_iy, r_wy, r_ix, r_wx, r_t, r_ih, r_iw = [], [], [], [], [], [], [], []
    for x, method in resize_branches:
        _chk(x, "x")
        x = x.contiguous()
        C = x.shape[1] if C is None else C
        if x.shape[0] != B or x.shape[1] != C:
            raise ValueError("upsample_merge_blk8: resize branch needs x [B,C,ih,iw]")
        iy, wy = _resize_tables(x.device, x.shape[2], int(H), method)
        ix, wx = _resize_tables(x.device, x.shape[3], int(W), method)
        keep += [x]
        r_in.append(x.data_ptr()); r_iy.append(iy.data_ptr()); r_wy.append(wy.data_ptr()); r_ix.append(ix.data_ptr())
        r_wx.append(wx.data_ptr()); r_t.append(iy.shape[1]); r_ih.append(x.shape[2]); r_iw.append(x.shape[3])
    check(lib.pcnn_upsample_merge_blk8(len(d_in), arr_p(d_in), arr_p(d_k), arr_p(d_b), arr_i(d_s), arr_i(d_ih), arr_i(d_iw),
                                       arr_i(d_act), len(r_in), arr_p(r_in), arr_p(r_iy), arr_p(r_wy), arr_p(r_ix), arr_p(r_wx),
                                       arr_i(r_t), arr_i(r_ih), arr_i(r_iw), float(alpha), _p(out.buf), _p(out.lo), out.mode,
                                       B, int(C or 0), int(H), int(W), out.C, int(c_offset), _stream()), "upsample_merge_blk8")
    if out.halo[0] != PAD_CONSTANT:
        out.halo = (out.halo[0], -1)
    return out


def pack_deconv_kernel_tc(kernel):
    """Keras deconv kernel [s,s,32,32] (k == stride) -> fp16 phase matrices [s][s][32][36] for upsample_merge_tc_blk8."""
    _chk(kernel, "kernel")
    s_, s2, C, C2 = kernel.shape
    if s_ != s2 or C != 32 or C2 != 32 or lib.pcnn_upsample_merge_tc_packed_bytes(int(s_)) == 0:
        raise ValueError("pack_deconv_kernel_tc: needs a [s,s,32,32] kernel with s <= 32")
    packed = torch.empty(lib.pcnn_upsample_merge_tc_packed_bytes(int(s_)) // 2, dtype=torch.float16, device=kernel.device)
    check(lib.pcnn_upsample_merge_tc_pack_kernel(_p(kernel.contiguous()), _p(packed), int(s_), _stream()), "upsample_merge_tc_pack_kernel")
    return packed


def upsample_merge_tc_fits(strides, resize_hw):
    """True if the tensor-core upsample-merge kernel can stage these branches in shared memory."""
    import ctypes
    st = (ctypes.c_int * max(len(strides), 1))(*[int(v) for v in strides])
    ih = (ctypes.c_int * max(len(resize_hw), 1))(*[int(h) for h, _ in resize_hw])
    iw = (ctypes.c_int * max(len(resize_hw), 1))(*[int(w) for _, w in resize_hw])
    n = lib.pcnn_upsample_merge_tc_smem_bytes(len(strides), st, len(resize_hw), ih, iw)
    return 0 < n <= 227 * 1024


def upsample_merge_tc_blk8(deconv_branches, resize_branches, alpha, out, c_offset, H, W, row_offset=0, full_H=None):
    """upsample_merge_blk8 with the transpose convolutions on the tensor cores, reading the branch outputs as BLK8 fp16.
    deconv_branches: [(Blk8 x with 32 channels [ih,iw], packed kernel (pack_deconv_kernel_tc), bias or None, stride, act)];
    resize_branches: [(x [B,32,ih,iw] fp32, method)].
    row_offset / full_H: `out` is a band of rows [row_offset, row_offset + H) of a full_H-row map (spatial decomposition):
    the resize tables are those of the full map, entered at row_offset; the deconv inputs are the band's own low-res rows."""
    import ctypes
    if not isinstance(out, Blk8) or (out.H, out.W) != (int(H), int(W)):
        raise ValueError("upsample_merge_tc_blk8: destination must be a Blk8 tensor of the output size")
    B = out.B
    keep = []

    def arr_p(vals):
        return (ctypes.c_void_p * max(len(vals), 1))(*[v for v in vals])

    def arr_i(vals):
        return (ctypes.c_int * max(len(vals), 1))(*[int(v) for v in vals])
    d_in, d_k, d_b, d_s, d_ih, d_iw, d_act = [], [], [], [], [], [], []
    for x, kern, bias, stride, act in deconv_branches:
        if not isinstance(x, Blk8) or x.C != 32 or x.B != B:
            raise ValueError("upsample_merge_tc_blk8: deconv branch inputs are 32-channel Blk8 tensors of the same batch")
        if kern.dtype != torch.float16 or kern.numel() * 2 != lib.pcnn_upsample_merge_tc_packed_bytes(int(stride)):
            raise ValueError("upsample_merge_tc_blk8: kernel must come from pack_deconv_kernel_tc() with s == stride")
        keep += [x, kern]
        d_in.append(x.buf.data_ptr()); d_k.append(kern.data_ptr()); d_b.append(_p(bias)); d_s.append(stride)
        d_ih.append(x.H); d_iw.append(x.W); d_act.append(act)
    r_in, r_iy, r_wy, r_ix, r_wx, r_t, r_ih, r_iw = [], [], [], [], [], [], [], []
    for x, method in resize_branches:
        _chk(x, "x")
        x = x.contiguous()
        if x.shape[0] != B or x.shape[1] != 32:
            raise ValueError("upsample_merge_tc_blk8: resize branch needs x [B,32,ih,iw]")
        iy, wy = _resize_tables(x.device, x.shape[2], int(full_H or H), method)
        ix, wx = _resize_tables(x.device, x.shape[3], int(W), method)
        keep += [x]
        iy, wy = iy[int(row_offset):], wy[int(row_offset):]
        r_in.append(x.data_ptr()); r_iy.append(iy.data_ptr()); r_wy.append(wy.data_ptr()); r_ix.append(ix.data_ptr())
        r_wx.append(wx.data_ptr()); r_t.append(iy.shape[1]); r_ih.append(x.shape[2]); r_iw.append(x.shape[3])
    check(lib.pcnn_upsample_merge_tc_blk8(len(d_in), arr_p(d_in), arr_p(d_k), arr_p(d_b), arr_i(d_s), arr_i(d_ih), arr_i(d_iw),
                                          arr_i(d_act), len(r_in), arr_p(r_in), arr_p(r_iy), arr_p(r_wy), arr_p(r_ix), arr_p(r_wx),
                                          arr_i(r_t), arr_i(r_ih), arr_i(r_iw), float(alpha), _p(out.buf), _p(out.lo), out.mode,
                                          B, int(H), int(W), out.C, int(c_offset), _stream()), "upsample_merge_tc_blk8")
    if out.halo[0] != PAD_CONSTANT:
        out.halo = (out.halo[0], -1)
    return out


def resize_add_blk8(resize_branches, alpha, out, c_offset, H, W, row_offset=0, full_H=None):
    """In place: channels [c_offset, c_offset+C) of the Blk8 tensor `out` += alpha * sum_r resize_r(x_r).
    resize_branches: [(x [B,C,ih,iw] fp32, method)].  The second pass of the merge on grids whose resize sources do not
    fit the fused kernel's shared memory (see pcnn.h)."""
    import ctypes
    if not isinstance(out, Blk8) or (out.H, out.W) != (int(H), int(W)) or not resize_branches:
        raise ValueError("resize_add_blk8: destination must be a Blk8 tensor of the output size and at least one branch")
    B, C, keep = out.B, None, []

    def arr_p(vals):
        return (ctypes.c_void_p * len(vals))(*vals)

    def arr_i(vals):
        return (ctypes.c_int * len(vals))(*[int(v) for v in vals])
    r_in, r_iy, r_wy, r_ix, r_wx, r_t, r_ih, r_iw = [], [], [], [], [], [], [], []
    for x, method in resize_branches:
        _chk(x, "x")
        x = x.contiguous()
        C = x.shape[1] if C is None else C
        if x.shape[0] != B or x.shape[1] != C:
            raise ValueError("resize_add_blk8: resize branches need x [B,C,ih,iw] with one C")
        iy, wy = _resize_tables(x.device, x.shape[2], int(full_H or H), method)      # row_offset / full_H: as in upsample_merge_tc_blk8
        ix, wx = _resize_tables(x.device, x.shape[3], int(W), method)
        keep += [x]
        iy, wy = iy[int(row_offset):], wy[int(row_offset):]
        r_in.append(x.data_ptr()); r_iy.append(iy.data_ptr()); r_wy.append(wy.data_ptr()); r_ix.append(ix.data_ptr())
        r_wx.append(wx.data_ptr()); r_t.append(iy.shape[1]); r_ih.append(x.shape[2]); r_iw.append(x.shape[3])
    check(lib.pcnn_resize_add_blk8(len(r_in), arr_p(r_in), arr_p(r_iy), arr_p(r_wy), arr_p(r_ix), arr_p(r_wx), arr_i(r_t), arr_i(r_ih),
                                   arr_i(r_iw), float(alpha), _p(out.buf), _p(out.lo), out.mode, B, int(C), int(H), int(W), out.C,
                                   int(c_offset), _stream()), "resize_add_blk8")
    return out


def rowweights_image(kernel, row_basis, cp, rt, T):
    """The fp16 operand image of pcnn_conv2d_tc_rowweights (layout in include/pcnn.h) and the power-of-two pre-scale:
    [ceil(Cin/16)][k][2][T][cp][8], slot t = output row (t//rt)*rt + rt-1 - t%rt (zeros beyond H), K half = channels
    16c + 8*half + [0,8).  Pure tensor algebra (any device): A_x[b,m,co] = sum_a kernel[a,b,m,co] * row_basis[m, x+a-k/2]."""
    k, _, Cin, Cout = kernel.shape
    H = row_basis.shape[1]
    pad = k // 2
    sp = torch.nn.functional.pad(row_basis, (pad, pad))                       # [Cin, H + 2 pad], zero-extended
    U = sp.unfold(1, k, 1)                                                    # [Cin, H, k]: U[m,x,a] = S[m, x+a-pad]
    A = torch.einsum("mxa,abmc->xbmc", U.double(), kernel.double()).float()   # [H, k, Cin, Cout]
    amax = float(A.abs().max())
    scale = 2.0 ** max(-24, min(24, math.floor(math.log2(512.0 / amax)))) if amax > 0 and math.isfinite(amax) else 1.0
    c16 = -(-Cin // 16)
    Ap = torch.zeros((H + 1, k, c16 * 16, cp), device=kernel.device, dtype=torch.float32)     # row H = zeros (slots beyond H)
    Ap[:H, :, :Cin, :Cout] = A * scale
    t = torch.arange(T, device=kernel.device)
    x_of_t = (t // rt) * rt + (rt - 1 - t % rt)
    x_of_t = torch.where(x_of_t < H, x_of_t, torch.full_like(x_of_t, H))
    P = Ap[x_of_t]                                                            # [T, k, c16*16, cp]
    P = P.reshape(T, k, c16, 2, 8, cp).permute(2, 1, 3, 0, 5, 4).contiguous() # [c16, k, 2, T, cp, 8]
    return P.half().contiguous(), scale


def pack_rowweights_tc(kernel, row_basis=None, x_res=None):
    """Row weights of a convolution whose input is separable, in[b,m,x,y] = h[b,m,y] * row_basis[m,x] (zero padding):
    A_x[b,m,co] = sum_a kernel[a,b,m,co] * row_basis[m, x+a-k/2], packed for pcnn_conv2d_tc_rowweights.  Done once per
    (layer, grid height).  With x_res (the DBCNN's case: Cin-2 sinh modes, posx, 1) the image comes from the library's
    host-side builder pcnn_host_rowweights -- the bytes the model-level C ABI uploads; with an explicit row_basis it is
    computed by rowweights_image (torch einsum, then the slot re-ordering)."""
    import ctypes
    _chk(kernel, "kernel")
    k, k2, Cin, Cout = kernel.shape
    if k != k2 or k % 2 == 0:
        raise ValueError("pack_rowweights_tc: needs an odd square kernel")
    H = int(x_res) if row_basis is None else row_basis.shape[1]
    cp = lib.pcnn_conv_tc_channel_slots(int(Cout), int(k))
    T = lib.pcnn_conv_tc_rowweight_slots(int(Cout), int(k), int(H))
    if cp == 0 or T == 0:
        raise ValueError("pack_rowweights_tc: unsupported layer (Cout <= 32, odd k <= 15)")
    rt = 5 if cp == 24 else 128 // cp
    if row_basis is None:
        kk = np.ascontiguousarray(kernel.detach().cpu().numpy(), dtype=np.float32)
        img = np.zeros(-(-Cin // 16) * k * 2 * T * cp * 8, dtype=np.float16)
        sc = ctypes.c_float(0.0)
        n = lib.pcnn_host_rowweights(kk.ctypes.data_as(ctypes.c_void_p), int(k), int(Cin), int(Cout), H,
                                     img.ctypes.data_as(ctypes.c_void_p), img.nbytes, ctypes.byref(sc))
        if n != img.size:
            check(int(n) if n < 0 else -1, "pcnn_host_rowweights")
        return {"packed": torch.from_numpy(img).to(kernel.device), "k": int(k), "cin": int(Cin), "cout": int(Cout), "H": H,
                "acc_scale": float(sc.value)}
    _chk(row_basis, "row_basis")
    if row_basis.shape[0] != Cin:
        raise ValueError("pack_rowweights_tc: needs a [Cin, H] row basis")
    packed, scale = rowweights_image(kernel, row_basis, cp, rt, T)
    return {"packed": packed, "k": int(k), "cin": int(Cin), "cout": int(Cout), "H": int(H), "acc_scale": 1.0 / scale}


def conv2d_tc_rowweights(x_row, wp, bias=None, act=ACT_LINEAR):
    """out[b,co,x,y] = act(sum A_x[b_tap,m,co] * x_row[b,m,y+b_tap-k/2] + bias): x_row is a Blk8 tensor with H = 1 (the 1-D
    signals), wp comes from pack_rowweights_tc.  Returns a Blk8 [B, Cout, wp['H'], W] (single fp16 pass)."""
    if not isinstance(x_row, Blk8) or x_row.H != 1 or x_row.mode != 1:
        raise ValueError("conv2d_tc_rowweights: x_row must be a single-pass Blk8 tensor with H = 1")
    if -(-wp["cin"] // 16) != -(-x_row.C // 16):
        raise ValueError("conv2d_tc_rowweights: weights expect %d input channels, tensor holds %d" % (wp["cin"], x_row.C))
    blk8_halo_fill(x_row, wp["k"] // 2, PAD_CONSTANT)
    out = Blk8(x_row.B, wp["cout"], wp["H"], x_row.W, x_row.device, split=1)
    check(lib.pcnn_conv2d_tc_rowweights(_p(x_row.buf), _p(wp["packed"]), _p(bias), _p(out.buf), x_row.B, wp["cin"], wp["cout"],
                                        out.C, wp["H"], x_row.W, wp["k"], int(act), wp["acc_scale"], _num_sms(x_row.device),
                                        _stream()), "conv2d_tc_rowweights")
    if out.halo[0] != PAD_CONSTANT:
        out.halo = (out.halo[0], -1)
    return out


def dbcnn_signal_blk8(h, modew):
    """The DBCNN boundary features as the one-row operand of the separable first 2-D convolution: Blk8 [B, M+2, 1, n] with
    channel m < M = h[b,m,y] * modew[b,m], channel M = 1, channel M+1 = posy[y]."""
    _chk(h, "h"); _chk(modew, "modew")
    h, modew = h.contiguous(), modew.contiguous()
    B, M, n = h.shape
    out = Blk8(B, M + 2, 1, n, h.device, split=1)
    check(lib.pcnn_dbcnn_signal_blk8(_p(h), _p(modew), _p(position_table(h.device, n)), _p(out.buf), B, M, n, _stream()), "dbcnn_signal_blk8")
    return out
