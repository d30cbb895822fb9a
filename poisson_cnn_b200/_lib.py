"""ctypes binding of libpcnn.so (include/pcnn.h).

There is deliberately NO fallback: if the shared library is missing or a symbol cannot be bound,
importing this module raises, and every product entry point that computes goes through it.
"""
import ctypes
import os
from ctypes import c_int, c_int64, c_float, c_void_p, c_size_t, c_char_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpcnn.so")

P = c_void_p  # every device pointer crosses the ABI as a plain address

# name -> (restype, argtypes); mirrors include/pcnn.h one to one
SIGNATURES = {
    "pcnn_version": (c_int, []),
    "pcnn_last_error": (c_char_p, []),
    "pcnn_launch_count": (ctypes.c_longlong, []),
    "pcnn_conv2d_f32": (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_float, c_int, c_int64, c_int64, c_int64, P]),
    "pcnn_boundary_stack_f32": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P, c_int, c_int, c_float, P]),
    "pcnn_smallmap_stack_f32": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P, c_int, c_int, c_float, P]),
    "pcnn_avgpool_same_f32": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int64, P]),
    "pcnn_deconv_same_f32": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_int, c_int, c_float, c_int, c_int64, P]),
    "pcnn_resize_f32": (c_int, [P, P, P, P, P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_float, c_int, c_int64, P]),
    "pcnn_spp_f32": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "pcnn_dense_f32": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P]),
    "pcnn_maxabs_f32": (c_int, [P, P, c_int, c_int64, P]),
    "pcnn_scale_inv_f32": (c_int, [P, P, P, c_int, c_int64, P]),
    "pcnn_hpnn_input_f32": (c_int, [P, P, P, P, c_int, c_int, c_int, P]),
    "pcnn_dbcnn_input_f32": (c_int, [P, c_float, P, P, c_int, c_int, P]),
    "pcnn_dbcnn_expand_f32": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_int, P]),
    "pcnn_dbcnn_finalize_f32": (c_int, [P, P, P, P, c_int, c_int, c_int, P]),
    "pcnn_hpnn_finalize_f32": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int64, P]),
    "pcnn_dense_input_f32": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "pcnn_merge_f32": (c_int, [P] * 12 + [c_int, c_int, c_int, P]),
    "pcnn_laplacian_residual_f32": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, P]),
    "pcnn_jacobi_sweep_f32": (c_int, [P, P, P, P, c_int, c_int, c_int, P]),
    "pcnn_dst_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pcnn_dst_sine_matrix": (c_int, [P, c_int, P]),
    "pcnn_dst_solve": (c_int, [P] * 10 + [c_int, c_int, c_int, P]),
    "pcnn_dst_fft_plan_bytes": (c_size_t, [c_int, c_int]),
    "pcnn_dst_fft_plan_init": (c_int, [P, c_int, c_int, P]),
    "pcnn_dst_fft_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pcnn_dst_fft_passes": (c_int, []),
    "pcnn_dst_solve_fft": (c_int, [P] * 9 + [c_int, c_int, c_int, c_int, P]),
    "pcnn_neumann_cg_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pcnn_neumann_laplacian_apply_f32": (c_int, [P, P, P, c_int, c_int, c_int, P]),
    "pcnn_neumann_cg_solve": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, ctypes.c_double, P, P, P]),
    "pcnn_blk8_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "pcnn_to_blk8": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int64, c_int, P]),
    "pcnn_from_blk8": (c_int, [P, P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int64, P]),
    "pcnn_blk8_halo_fill": (c_int, [P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "pcnn_conv_tc_packed_weight_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "pcnn_conv_tc_channel_slots": (c_int, [c_int, c_int]),
    "pcnn_conv_tc_pack_weights": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_float, P]),
    "pcnn_conv2d_tc": (c_int, [P] * 11 + [c_int] * 10 + [c_float, c_int, c_int, P]),
    "pcnn_conv_tc_rowweight_slots": (c_int, [c_int, c_int, c_int]),
    "pcnn_dbcnn_signal_blk8": (c_int, [P, P, P, P, c_int, c_int, c_int, P]),
    "pcnn_conv2d_tc_rowweights": (c_int, [P, P, P, P] + [c_int] * 8 + [c_float, c_int, P]),
    "pcnn_upsample_merge_blk8": (c_int, [c_int, P, P, P, P, P, P, P, c_int, P, P, P, P, P, P, P, P, c_float, P, P,
                                         c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "pcnn_upsample_merge_tc_blk8": (c_int, [c_int, P, P, P, P, P, P, P, c_int, P, P, P, P, P, P, P, P, c_float, P, P,
                                            c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "pcnn_resize_add_blk8": (c_int, [c_int, P, P, P, P, P, P, P, P, c_float, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "pcnn_upsample_merge_tc_packed_bytes": (c_size_t, [c_int]),
    "pcnn_upsample_merge_tc_smem_bytes": (c_size_t, [c_int, P, c_int, P, P]),
    "pcnn_upsample_merge_tc_pack_kernel": (c_int, [P, P, c_int, P]),
    "pcnn_upsample_merge_packed_floats": (c_size_t, [c_int, c_int]),
    "pcnn_upsample_merge_pack_kernel": (c_int, [P, P, c_int, c_int, P]),
    "pcnn_dbcnn_expand_blk8": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    # model-level API (csrc/engine.cu); a pcnn_handle crosses the ABI as a plain address
    "pcnn_create": (c_int, [c_char_p, c_int, ctypes.POINTER(c_void_p)]),
    "pcnn_destroy": (c_int, [P]),
    "pcnn_set_weight": (c_int, [P, c_char_p, P, ctypes.POINTER(c_int64), c_int, c_int]),
    "pcnn_finalize_weights": (c_int, [P, c_int]),
    "pcnn_set_microbatch": (c_int, [P, c_int]),
    "pcnn_workspace_bytes": (c_int, [P, c_int, c_int, c_int, ctypes.POINTER(c_size_t)]),
    "pcnn_hpnn_workspace_bytes": (c_int, [P, c_int, c_int, c_int, ctypes.POINTER(c_size_t)]),
    "pcnn_dbcnn_workspace_bytes": (c_int, [P, c_int, c_int, c_int, ctypes.POINTER(c_size_t)]),
    "pcnn_hpnn_forward": (c_int, [P, P, P, P, c_int, c_int, c_int, P, c_size_t, P]),
    "pcnn_dbcnn_forward": (c_int, [P, P, P, P, c_int, c_int, c_int, P, c_size_t, P]),
    "pcnn_forward": (c_int, [P] * 8 + [c_int, c_int, c_int, P, c_size_t, P]),
    "pcnn_host_table": (ctypes.c_longlong, [c_char_p, c_int, c_int, c_int, P, c_size_t]),
    "pcnn_host_rowweights": (ctypes.c_longlong, [P, c_int, c_int, c_int, c_int, P, c_size_t, ctypes.POINTER(c_float)]),
    "pcnn_profile_conv_begin": (c_int, [P, c_int, c_int, c_int, c_int]),
    "pcnn_profile_conv_end": (c_int, [P, ctypes.POINTER(c_int), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
}


class PcnnError(RuntimeError):
    pass


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "poisson_cnn_b200: %s not found. Build it with `make` (or `python -c 'import __graft_entry__ as g; "
            "g.build()'`) -- there is no CPU or PyTorch fallback for the compute path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    if lib.pcnn_version() < 200:
        raise ImportError("libpcnn.so is older than the Python host side; rebuild")
    return lib


lib = _load()


def check(status, what=""):
    """Turn a negative pcnn_status into the exception the reference would have raised:
    ValueError for bad arguments/configs, RuntimeError for CUDA failures."""
    if status == 0:
        return
    msg = lib.pcnn_last_error().decode("utf-8", "replace")
    if status == -1 or status == -3:
        raise ValueError("%s: %s" % (what or "pcnn", msg))
    raise PcnnError("%s: %s" % (what or "pcnn", msg))
