"""ctypes binding of libnsvf_b200.so (C ABI declared in include/nsvf_b200.h).

There is no CPU fallback anywhere in this package: if the shared library is missing the import of
any operator fails loudly, and every operator rejects non-CUDA tensors.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libnsvf_b200.so")

c_int, c_ll, c_float, c_size_t, c_void_p = (ctypes.c_int, ctypes.c_longlong, ctypes.c_float,
                                            ctypes.c_size_t, ctypes.c_void_p)

# name -> (restype, argtypes); must list every symbol of include/nsvf_b200.h (tests check this)
SIGNATURES = {
    "nsvf_version": (c_int, []),
    "nsvf_last_error": (ctypes.c_char_p, []),
    "nsvf_kernel_launches": (ctypes.c_ulonglong, []),
    "nsvf_profile_kernel": (c_int, [ctypes.c_char_p, c_void_p, c_void_p]),
    "nsvf_profile_begin": (c_int, [ctypes.c_char_p]),
    "nsvf_profile_end": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "nsvf_ref_rcp": (c_int, [c_void_p, c_ll, c_void_p, c_void_p]),
    "nsvf_aabb_workspace_bytes": (c_size_t, [c_int, c_int]),
    "nsvf_aabb_intersect": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p,
                                    c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t]),
    "nsvf_aabb_intersect_sorted": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_int, c_float, c_void_p, c_void_p,
                                           c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t]),
    "nsvf_sort_hits_by_depth": (c_int, [c_void_p, c_ll, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nsvf_aabb_hit_mask": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_ll,
                                   c_void_p, c_void_p, c_size_t]),
    "nsvf_aabb_prepare": (c_int, [c_void_p, c_int, c_int, c_float, c_void_p, c_ll, c_void_p, c_size_t]),
    "nsvf_aabb_intersect_prepared": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_float, c_void_p,
                                             c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_size_t]),
    "nsvf_ball_intersect": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p, c_ll,
                                    c_void_p, c_void_p, c_void_p]),
    "nsvf_triangle_intersect": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_float, c_int, c_void_p, c_void_p,
                                        c_void_p, c_ll, c_void_p, c_void_p, c_void_p]),
    "nsvf_svo_workspace_bytes": (c_size_t, [c_int, c_int]),
    "nsvf_svo_intersect": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t]),
    "nsvf_svo_sorted_workspace_bytes": (c_size_t, [c_int, c_int, c_ll]),
    "nsvf_svo_intersect_sorted": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_int, c_float, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_size_t]),
    "nsvf_svo_prepare": (c_int, [c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_ll, c_void_p, c_size_t]),
    "nsvf_svo_ray_scratch_bytes": (c_size_t, [c_int, c_ll]),
    "nsvf_svo_intersect_sorted_prepared": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_int, c_float, c_void_p,
                                                   c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_void_p,
                                                   c_void_p, c_void_p, c_size_t, c_void_p, c_size_t]),
    "nsvf_inverse_cdf_sampling": (c_int, [c_void_p, c_int, c_int, c_ll, c_int, c_int, c_int, c_float,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p]),
    "nsvf_inverse_cdf_sampling_ex": (c_int, [c_void_p, c_int, c_int, c_ll, c_int, c_int, c_int, c_float,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                             c_int]),
    "nsvf_inverse_cdf_plan": (c_int, [c_void_p, c_int, c_int, c_ll, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nsvf_inverse_cdf_block": (c_int, [c_void_p, c_ll, c_int, c_int, c_float, c_int, c_int] + [c_void_p] * 7 +
                               [c_ll, c_float, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p]),
    "nsvf_inverse_cdf_stream_state_bytes": (c_size_t, [c_ll]),
    "nsvf_inverse_cdf_stream": (c_int, [c_void_p, c_ll, c_int, c_int, c_float, c_int, c_int] + [c_void_p] * 7 +
                                [c_ll, c_float, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nsvf_march_plane_stride": (c_ll, [c_ll]),
    "nsvf_march_plan_bytes": (c_size_t, [c_ll, c_int]),
    "nsvf_march_ray_lengths": (c_int, [c_void_p, c_ll, c_int, c_ll, c_void_p, c_void_p, c_void_p]),
    "nsvf_march_transpose": (c_int, [c_void_p, c_ll, c_int, c_ll, c_int, c_int] + [c_void_p] * 8),
    "nsvf_march_begin": (c_int, [c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int]),
    "nsvf_march_compact": (c_int, [c_void_p, c_ll, c_int, c_int, c_int] + [c_void_p] * 13 + [c_int]),
    "nsvf_march_epilogue": (c_int, [c_void_p, c_ll, c_int, c_int, c_int] + [c_void_p] * 9 + [c_float, c_void_p,
                                                                                             c_void_p, c_int, c_int,
                                                                                             c_void_p, c_void_p]),
    "nsvf_march_epilogue_bwd": (c_int, [c_void_p, c_ll, c_int, c_int, c_int] + [c_void_p] * 8),
    "nsvf_march_composite_fwd": (c_int, [c_void_p, c_ll, c_int] + [c_void_p] * 13 + [c_ll, c_float, c_int]),
    "nsvf_march_composite_bwd": (c_int, [c_void_p, c_ll, c_int] + [c_void_p] * 11),
    "nsvf_uniform_ray_sampling": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float] + [c_void_p] * 8),
    "nsvf_octree_build": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_void_p, c_void_p]),
    "nsvf_octree_flatten": (c_int, [c_void_p, c_void_p, c_ll]),
    "nsvf_trilinear_embed_fwd": (c_int, [c_void_p, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_float, c_void_p]),
    "nsvf_trilinear_embed_bwd": (c_int, [c_void_p, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_float, c_void_p, c_void_p, c_void_p]),
    "nsvf_composite_fwd": (c_int, [c_void_p, c_ll, c_int] + [c_void_p] * 7),
    "nsvf_composite_bwd": (c_int, [c_void_p, c_ll, c_int] + [c_void_p] * 9),
    "nsvf_split_workspace_bytes": (c_size_t, [c_void_p]),
    "nsvf_split_mark": (c_int, [c_void_p, c_int, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_size_t]),
    "nsvf_split_emit": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t]),
    "nsvf_prune_lattice_embed": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float,
                                         c_void_p]),
    "nsvf_prune_keep": (c_int, [c_void_p, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p]),
    "nsvf_masked_col_counts": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "nsvf_fill_in_blend": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                   c_void_p, c_void_p, c_void_p]),
    "nsvf_track_voxel_probs": (c_int, [c_void_p, c_ll, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "nsvf_ln_relu_fwd": (c_int, [c_void_p, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                 c_void_p]),
    "nsvf_ln_relu_bwd_workspace_bytes": (c_size_t, [c_ll, c_int]),
    "nsvf_ln_relu_bwd": (c_int, [c_void_p, c_ll, c_int] + [c_void_p] * 11 + [c_size_t]),
    "nsvf_posenc_fwd": (c_int, [c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "nsvf_posenc_bwd": (c_int, [c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "nsvf_narrow_linear_supported": (c_int, [c_int, c_int]),
    "nsvf_narrow_linear_fwd": (c_int, [c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nsvf_narrow_linear_bwd_workspace_bytes": (c_size_t, [c_ll, c_int, c_int]),
    "nsvf_narrow_linear_bwd": (c_int, [c_void_p, c_ll, c_int, c_int] + [c_void_p] * 7 + [c_size_t]),
    "nsvf_compact_count": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "nsvf_compact_fill": (c_int, [c_void_p, c_ll, c_int, c_int, c_int] + [c_void_p] * 12),
}

_lib = None


def load():
    """Load the shared library once. Raises ImportError (never falls back) when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "nsvf_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C nsvf_b200/csrc`. There is no CPU / PyTorch fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("nsvf_b200: " + load().nsvf_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device (or host) pointer of a tensor, None -> NULL."""
    return None if t is None else t.data_ptr()


def current_stream(device):
    import torch
    return torch.cuda.current_stream(device).cuda_stream


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NO_GUARD = _NoGuard()


def device_guard(device):
    """`with torch.cuda.device(device)` only when `device` is not already current (the context manager costs several
    microseconds per call, which matters on paths that launch one small kernel per call)."""
    import torch
    idx = device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(device)
