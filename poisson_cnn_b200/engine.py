"""ctypes front end of the model-level C ABI (include/pcnn.h: pcnn_create ... pcnn_forward; csrc/engine.cu).

The layer program of the three reference calls (model([rhs, dx]), model([bc, dx, x_res]), model([rhs, left, top, right,
bottom, dx])) runs inside libpcnn.so; this module only owns device memory (torch) and hands raw pointers across the ABI:
the weights once, then inputs, output and ONE workspace per forward call.  torch is plumbing (allocation, streams).
"""
import ctypes
import json

import numpy as np
import torch

from ._lib import lib, check

PRECISION_CODE = {"fp32": 0, "tc": 1, "tc3": 2, "tc2": 3, "mixed": 4}


class Engine:
    """One pcnn_handle: config + weights + packed operands on one device."""

    MAX_WORKSPACES = 3      # workspaces are kept per shape (the handle caches tables and halo state in them)

    def __init__(self, config, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("the engine runs on a CUDA device (there is no CPU path)")
        self._h = ctypes.c_void_p()
        text = json.dumps(config)
        check(lib.pcnn_create(text.encode("utf-8"), self.device.index or 0, ctypes.byref(self._h)), "pcnn_create")
        self.precision = None
        self._ws = {}           # shape key -> uint8 tensor (insertion-ordered: oldest first)
        self._microbatch = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib.pcnn_destroy(self._h)
            self._h = ctypes.c_void_p()
        self._ws = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ setup
    def set_weights(self, weights):
        """weights: {name: array} with the names of poisson_cnn_b200/weights.py ('hpnn/...', 'dbcnn/...')."""
        for name, a in weights.items():
            a = np.ascontiguousarray(a)
            if a.dtype == np.float64:
                code = 1
            else:
                a = np.ascontiguousarray(a, dtype=np.float32)
                code = 0
            shape = (ctypes.c_int64 * a.ndim)(*a.shape)
            check(lib.pcnn_set_weight(self._h, name.encode("utf-8"), a.ctypes.data_as(ctypes.c_void_p), shape, a.ndim, code),
                  "pcnn_set_weight(%s)" % name)
        self.precision = None
        return self

    def finalize(self, precision):
        if precision not in PRECISION_CODE:
            raise ValueError("precision must be one of %s" % (tuple(PRECISION_CODE),))
        with torch.cuda.device(self.device):
            check(lib.pcnn_finalize_weights(self._h, PRECISION_CODE[precision]), "pcnn_finalize_weights")
        self.precision = precision
        self._ws = {}
        return self

    def set_microbatch(self, samples):
        samples = int(samples or 0)
        if samples != self._microbatch:
            check(lib.pcnn_set_microbatch(self._h, samples), "pcnn_set_microbatch")
            self._microbatch = samples
            self._ws = {}

    # ------------------------------------------------------------------ workspace
    def workspace_bytes(self, kind, B, d0, d1):
        n = ctypes.c_size_t(0)
        fn = {"pcnn": lib.pcnn_workspace_bytes, "hpnn": lib.pcnn_hpnn_workspace_bytes, "dbcnn": lib.pcnn_dbcnn_workspace_bytes}[kind]
        check(fn(self._h, int(B), int(d0), int(d1), ctypes.byref(n)), "pcnn_workspace_bytes")
        return int(n.value)

    def _workspace(self, kind, B, d0, d1):
        key = (kind, int(B), int(d0), int(d1))
        ws = self._ws.get(key)
        if ws is None:
            need = self.workspace_bytes(kind, B, d0, d1)
            while len(self._ws) >= self.MAX_WORKSPACES:
                self._ws.pop(next(iter(self._ws)))
            try:
                ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            except torch.OutOfMemoryError:
                self._ws = {}
                torch.cuda.empty_cache()
                ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        else:
            self._ws[key] = self._ws.pop(key)      # most recently used last
        return ws

    def release_workspaces(self):
        self._ws = {}

    # ------------------------------------------------------------------ forward calls
    @staticmethod
    def _f32(t, name, shape=None):
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise ValueError("%s must be a CUDA tensor (the hot path has no CPU implementation)" % name)
        if t.dtype != torch.float32:
            raise ValueError("%s must be float32, got %s" % (name, t.dtype))
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError("%s must have shape %s, got %s" % (name, tuple(shape), tuple(t.shape)))
        return t.contiguous()

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def forward(self, rhs, left, top, right, bottom, dx, out=None):
        rhs = self._f32(rhs, "rhs")
        B, _, nx, ny = rhs.shape
        left, right = self._f32(left, "left", (B, 1, ny)), self._f32(right, "right", (B, 1, ny))
        top, bottom = self._f32(top, "top", (B, 1, nx)), self._f32(bottom, "bottom", (B, 1, nx))
        dx = self._f32(dx, "dx", (B, 1))
        if out is None:
            out = torch.empty((B, 1, nx, ny), device=rhs.device, dtype=torch.float32)
        ws = self._workspace("pcnn", B, nx, ny)
        with torch.cuda.device(self.device):
            check(lib.pcnn_forward(self._h, rhs.data_ptr(), left.data_ptr(), top.data_ptr(), right.data_ptr(), bottom.data_ptr(),
                                   dx.data_ptr(), out.data_ptr(), B, nx, ny, ws.data_ptr(), ws.numel(), self._stream()), "pcnn_forward")
        return out

    def hpnn_forward(self, rhs, dx, out=None):
        rhs = self._f32(rhs, "rhs")
        B, _, H, W = rhs.shape
        dx = self._f32(dx, "dx", (B, 1))
        if out is None:
            out = torch.empty((B, 1, H, W), device=rhs.device, dtype=torch.float32)
        ws = self._workspace("hpnn", B, H, W)
        with torch.cuda.device(self.device):
            check(lib.pcnn_hpnn_forward(self._h, rhs.data_ptr(), dx.data_ptr(), out.data_ptr(), B, H, W, ws.data_ptr(), ws.numel(),
                                        self._stream()), "pcnn_hpnn_forward")
        return out

    def dbcnn_forward(self, bc, dx, x_res, out=None):
        bc = self._f32(bc, "bc")
        B, _, n = bc.shape
        dx = self._f32(dx, "dx", (B, 1))
        x_res = int(x_res)
        if out is None:
            out = torch.empty((B, 1, x_res, n), device=bc.device, dtype=torch.float32)
        ws = self._workspace("dbcnn", B, n, x_res)
        with torch.cuda.device(self.device):
            check(lib.pcnn_dbcnn_forward(self._h, bc.data_ptr(), dx.data_ptr(), out.data_ptr(), B, n, x_res, ws.data_ptr(), ws.numel(),
                                         self._stream()), "pcnn_dbcnn_forward")
        return out

    # ------------------------------------------------------------------ live kernel timing (bench.py roofline)
    def profile_conv_begin(self, cin, cout, k, max_launches=256):
        check(lib.pcnn_profile_conv_begin(self._h, int(cin), int(cout), int(k), int(max_launches)), "pcnn_profile_conv_begin")

    def profile_conv_end(self):
        n, ms, fl = ctypes.c_int(0), ctypes.c_double(0.0), ctypes.c_double(0.0)
        check(lib.pcnn_profile_conv_end(self._h, ctypes.byref(n), ctypes.byref(ms), ctypes.byref(fl)), "pcnn_profile_conv_end")
        if n.value == 0:
            return None
        return {"launches": n.value, "avg_ms": ms.value, "flops_per_launch": fl.value}
