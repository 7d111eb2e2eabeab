"""ctypes binding of libstinet_b200.so -- the C ABI declared in include/stinet_b200.h.

This is the ONLY way the Python host code reaches the CUDA kernels, and there is no fallback: if the library is
missing or the device is not an sm_100 part, importing / calling raises (the product path must fail loudly).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("STINET_B200_LIB") or os.path.join(_HERE, "libstinet_b200.so")   # override: debug builds only

P, I64, I, F, SZ = c_void_p, c_int64, c_int, c_float, c_size_t
D, U64 = ctypes.c_double, ctypes.c_uint64

# name -> (restype, argtypes); mirrors include/stinet_b200.h one to one
SIGNATURES = {
    "stinet_abi_version": (I, []),
    "stinet_last_error": (c_char_p, []),
    "stinet_device_ok": (I, []),
    "stinet_launch_count": (ctypes.c_longlong, []),
    "stinet_csr_workspace_bytes": (SZ, [I64, I64]),
    "stinet_csr_build": (I, [P, P, I64, I64, P, P, P, P, P, P, SZ, P]),
    "stinet_csr_cross_positions": (I, [P, P, I64, P, P, P]),
    "stinet_concat_i32": (I, [P, P, P, P, I, P, P]),
    "stinet_aggregate_fwd": (I, [P, I64, P, P, P, I64, I64, I64, I, P, I64, P, P]),
    "stinet_aggregate_bwd": (I, [P, I64, P, P, P, P, P, I64, I64, I, P, I64, P]),
    "stinet_edge_message_fwd": (I, [P, I64, P, I64, P, P, I64, I64, P, I64, P]),
    "stinet_edge_message_bwd_target": (I, [P, I64, P, I64, P, I64, P, P, I64, I64, P, I64, P]),
    "stinet_edge_message_bwd_source": (I, [P, I64, P, I64, P, I64, P, P, P, I64, I64, P, I64, P]),
    "stinet_edge_message_fwd_mask": (I, [P, I64, P, I64, P, P, I64, I64, P, I64, P, P]),
    "stinet_edge_message_bwd_target_mask": (I, [P, I64, P, P, I64, I64, P, I64, P]),
    "stinet_edge_message_bwd_source_mask": (I, [P, I64, P, P, P, P, P, I64, I64, P, I64, P]),
    "stinet_edgeconv_hoist_fwd": (I, [P, I64, P, I64, I64, I, P, P, P]),
    "stinet_edgeconv_hoist_bwd": (I, [P, P, I64, I64, I, P, I64, P, P]),
    "stinet_pool_max_fwd": (I, [P, I64, P, P, I64, I64, I64, P, I64, P, P]),
    "stinet_pool_max_bwd": (I, [P, I64, P, P, I64, I64, P, I64, P]),
    "stinet_pool_mean_fwd": (I, [P, I64, P, P, I64, I64, P, I64, P]),
    "stinet_pool_mean_bwd": (I, [P, I64, P, P, I64, I64, P, I64, P]),
    "stinet_pool_max_i32": (I, [P, P, P, I64, P, P]),
    "stinet_unpool_fwd": (I, [P, I64, P, I64, I64, P, I64, P]),
    "stinet_unpool_bwd": (I, [P, I64, P, P, I64, I64, P, I64, P]),
    "stinet_segnorm_workspace_bytes": (SZ, [I64, I64, I64]),
    "stinet_segnorm_stats": (I, [P, I64, I64, I64, I64, I64, P, P, P, F, P, P, P, SZ, P]),
    "stinet_segnorm_fwd": (I, [P, I64, I64, I64, I64, I64, P, P, F, P, I64, I, P, I64, P, P, P, P, P, I64, P, P, P, SZ, P]),
    "stinet_segnorm_apply": (I, [P, I64, I64, I64, P, P, P, P, I64, I, P, I64, P]),
    "stinet_segnorm_bwd": (I, [P, I64, P, I64, I64, I64, I64, I64, P, P, P, P, P, I, P, I64, P, P, P, SZ, P]),
    "stinet_affnorm_workspace_bytes": (SZ, [I64, I64, I64]),
    "stinet_affnorm_fwd": (I, [P, I64, I64, I64, I64, I64, P, P, P, I, F, P, P, P, P, I64, P, P, P, SZ, P]),
    "stinet_affnorm_apply": (I, [P, I64, I64, I64, P, P, P, P, P, P, P, I64, P]),
    "stinet_affnorm_bwd": (I, [P, I64, P, I64, I64, I64, I64, I64, P, P, P, I, P, P, P, P, P, I64, P, P, P, P, SZ, P]),
    "stinet_bn_running_update": (I, [P, P, I64, F, F, I64, P, P, P]),
    "stinet_weight_entry_bytes": (SZ, []),
    "stinet_weight_entry_fill": (I64, [P, P, I64, P, I64, I64, I, P, P, I64, P, P, P, I64]),
    "stinet_weight_planes_refresh": (I, [P, I, I64, P, P]),
    "stinet_head_workspace_bytes": (SZ, [I64, I64]),
    "stinet_head_fwd": (I, [P, I64, P, P, I64, I64, P, P]),
    "stinet_head_bwd": (I, [P, I64, P, P, P, I64, I64, P, I64, P, P, P, SZ, P]),
    "stinet_masked_l1_workspace_bytes": (SZ, [I64]),
    "stinet_masked_l1_fwd": (I, [P, P, P, I64, I64, P, P, SZ, P]),
    "stinet_masked_l1_bwd": (I, [P, P, P, P, I64, I64, P, P]),
    "stinet_sort_workspace_bytes": (SZ, [I64]),
    "stinet_sort_pairs_u64": (I, [P, P, P, P, I64, I, P, SZ, P]),
    "stinet_csr_degree_order_workspace_bytes": (SZ, [I64]),
    "stinet_csr_degree_order": (I, [P, I64, P, P, SZ, P]),
    "stinet_voxel_bins": (I, [P, I, I64, D, P, P, P]),
    "stinet_voxel_keys": (I, [P, P, I64, P, P]),
    "stinet_unique_workspace_bytes": (SZ, [I64]),
    "stinet_unique_sorted_u64": (I, [P, I64, U64, P, P, P, SZ, P]),
    "stinet_cluster_finish": (I, [P, P, I64, I64, P, P, P]),
    "stinet_cluster_centroids": (I, [P, I, P, P, I64, P, P]),
    "stinet_coarse_edge_keys": (I, [P, P, I64, P, I64, I64, P, P, P]),
    "stinet_coarse_edges_emit": (I, [P, P, I64, I64, I64, P, P]),
    "stinet_metrics_workspace_bytes": (SZ, [I64]),
    "stinet_graph_laplace": (I, [P, I64, P, P, I64, I64, P, I64, P]),
    "stinet_graph_laplace_variance": (I, [P, I64, P, P, I64, P, P, SZ, P]),
    "stinet_graph_total_variation": (I, [P, I64, P, P, I64, I64, P, P, SZ, P]),
    "stinet_psnr": (I, [P, I64, P, I64, P, I64, I64, F, P, P, SZ, P]),
    "stinet_gemm_workspace_bytes": (SZ, [I64, I64, I64, I]),
    "stinet_linear_fwd": (I, [P, I64, P, I64, P, P, P, I64, I64, I64, I64, I, P, SZ, P]),
    "stinet_linear_dgrad": (I, [P, I64, P, I64, P, I64, I64, I64, I64, I, P, SZ, P]),
    "stinet_linear_wgrad": (I, [P, I64, P, I64, P, P, I64, P, I64, I64, I64, I, P, SZ, P]),
    "stinet_f16_amax": (I, [P, I64, I64, I64, P, P]),
    "stinet_f16_split": (I, [P, I64, I64, I64, P, P, P, I64, P, P]),
    "stinet_linear_fwd_f16": (I, [P, P, I64, P, P, P, I64, P, P, P, P, I64, P, I64, I64, I64, I, P, SZ, P]),
    "stinet_linear_dgrad_f16": (I, [P, P, I64, P, P, P, I64, P, P, I64, P, I64, I64, I64, I, P, SZ, P]),
    "stinet_linear_wgrad_f16": (I, [P, P, I64, P, P, P, I64, P, P, I64, I64, I64, I64, I, P, SZ, P]),
    "stinet_colsum": (I, [P, I64, P, I64, I64, P, P, SZ, P]),
    "stinet_f16_split_colsum": (I, [P, I64, I64, I64, P, P, P, P, I64, P, P, P, SZ, P]),
    "stinet_colsum_planes": (I, [P, P, I64, P, I64, I64, P, P, SZ, P]),
    "stinet_csr_dq_factor": (I, [P, P, P, I64, P, P]),
    "stinet_edge_message_fwd_planes": (I, [P, I64, P, I64, P, P, I64, I64, P, P, P, I64, P, P, P]),
    "stinet_edge_message_bwd_planes": (I, [P, I64, P, P, P, P, P, P, P, I64, I64, P, P, I64, P, P, P, SZ, P]),
    "stinet_edge_message_bwd_workspace_bytes": (SZ, [I64, I64]),
}

REDUCE = {"add": 0, "sum": 0, "mean": 1, "max": 2}
PREC = {"fp32": 0, "bf16": 1, "fp32_simt": 2, "tf32": 3, "bf16x3": 4, "fp32_tf32x3": 0}
# arithmetic modes that run on fp16 operand planes (stinet_linear_*_f16): name -> tcgen05 passes per product
PLANE_PASSES = {"fp32": 3, "f16": 1}
ACT_NONE, ACT_ELU = 0, 1


class StinetError(RuntimeError):
    pass


_lib = None


def load() -> ctypes.CDLL:
    """dlopen the library and type every entry point.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StinetError(
            f"{LIB_PATH} not found: build it with `make -C surface-texture-inpainting-net_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU/PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype, fn.argtypes = res, args
    if lib.stinet_abi_version() != 1:
        raise StinetError(f"ABI version mismatch: library {lib.stinet_abi_version()} != binding 1")
    _lib = lib
    return lib


class KernelProfiler:
    """Optional per-entry-point device timing with CUDA events on the launching stream (bench.py's roofline leg).
    `cost` = (algorithmic bytes, flops, tag) as defined in DESIGN.md / SURVEY 8d; never active on the timed path."""

    def __init__(self):
        self.events = []          # (key, ev0, ev1, bytes, flops)

    def __enter__(self):
        global _profiler
        _profiler = self
        return self

    def __exit__(self, *exc):
        global _profiler
        _profiler = None

    def summary(self):
        import torch
        torch.cuda.synchronize()
        out = {}
        for key, e0, e1, nbytes, flops in self.events:
            r = out.setdefault(key, {"calls": 0, "ms": 0.0, "bytes": 0, "flops": 0})
            r["calls"] += 1
            r["ms"] += e0.elapsed_time(e1)
            r["bytes"] += nbytes
            r["flops"] += flops
        return out


_profiler = None


def call(name: str, *args, cost=None) -> None:
    """Invoke an int-returning entry point; non-zero becomes StinetError with the library's message."""
    lib = load()
    if _profiler is None:
        rc = getattr(lib, name)(*args)
    else:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        nbytes, flops, tag = cost if cost is not None else (0, 0, "")
        _profiler.events.append((f"{name[7:]}[{tag}]" if tag else name[7:], e0, e1, int(nbytes), int(flops)))
    if rc != 0:
        raise StinetError(f"{name} failed ({rc}): {lib.stinet_last_error().decode()}")


def query(name: str, *args) -> int:
    return int(getattr(load(), name)(*args))
