"""ctypes binding of libmsm_we_b200.so (the C ABI declared in include/msm_we_b200.h).

There is no CPU fallback: if the shared library is missing or a symbol is absent the import fails
loudly.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` (or ``make -C
msm_we_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmsm_we_b200.so")

ABI_VERSION = 1
ERR_SLOTS = 4
ERR_OUT_OF_BINSPACE, ERR_NO_CENTERS, ERR_LABEL_RANGE, ERR_INTERNAL = 0, 1, 2, 3
FLAG_BASIS, FLAG_TARGET = 1, 2
MAPPER_RECTILINEAR, MAPPER_VORONOI, MAPPER_PRECOMPUTED = 0, 1, 2
ASSIGN_FP64, ASSIGN_TF32X3, ASSIGN_AUTO = 0, 1, 2
ASSIGN_REUSE_BUCKETS = 0x100

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_int = C.c_int
_sz = C.c_size_t
_f64 = C.c_double

# name -> (restype, argtypes); kept in the order of include/msm_we_b200.h
SIGNATURES = {
    "mwe_abi_version": (_int, []),
    "mwe_last_error": (C.c_char_p, []),
    "mwe_device_sm_count": (_int, []),
    "mwe_project_f64": (_int, [_p, _i64, _int, _i64, _p, _p, _int, _p, _i64, _p]),
    "mwe_device_malloc": (_int, [_sz, _p]),
    "mwe_device_free": (_int, [_p]),
    "mwe_ipc_export": (_int, [_p, _p]),
    "mwe_ipc_open": (_int, [_p, _p]),
    "mwe_ipc_close": (_int, [_p]),
    "mwe_flux_peer_allreduce_f64": (_int, [_p, _p, _p, _int, _int, _i64, C.c_double, C.c_uint32, _p, _p, _p]),
    "mwe_host_register": (_int, [_p, _sz]),
    "mwe_host_unregister": (_int, [_p]),
    "mwe_set_timing_events": (_int, [_p, _p]),
    "mwe_bin_flags_f64": (_int, [_p, _i64, _int, _int, _p, _p, _i32, _p, _p, _p, _p, _p, _p, _p, _p]),
    "mwe_rows_with_nan_f64": (_int, [_p, _i64, _int, _i64, _p, _p]),
    "mwe_assign_workspace_bytes": (_sz, [_i64, _i32]),
    "mwe_assign_workspace_bytes_ex": (_sz, [_i64, _i32, _int, _i32, _int]),
    "mwe_centers_sqnorm_f64": (_int, [_p, _i64, _int, _p, _p]),
    "mwe_assign_stratified_f64": (_int, [_p, _i64, _int, _i64, _p, _p, _p, _p, _p, _i32, _i32, _int, _p, _p, _p, _p, _sz, _p, _p]),
    "mwe_centroid_workspace_bytes": (_sz, [_i64, _i64]),
    "mwe_centroid_accumulate_f64": (_int, [_p, _i64, _int, _i64, _p, _p, _i64, _p, _p, _p, _sz, _p]),
    "mwe_minibatch_update_f64": (_int, [_p, _i64, _int, _i64, _p, _p, _i64, _p, _p, _p, _sz, _p]),
    "mwe_lloyd_finalize_f64": (_int, [_p, _p, _i64, _int, _p, _p]),
    "mwe_minibatch_finalize_f64": (_int, [_p, _p, _i64, _int, _p, _p, _p]),
    "mwe_point_center_dist2_f64": (_int, [_p, _i64, _int, _p, _i64, _p, _p, _p, _p]),
    "mwe_group_by_label": (_int, [_p, _i64, _i64, _p, _p, _p, _sz, _p]),
    "mwe_label_stats_f64": (_int, [_p, _i64, _p, _p, _i64, _p, _p, _p, _p, _p]),
    "mwe_segment_topk_f64": (_int, [_p, _p, _p, _p, _i32, _int, _p, _p, _p]),
    "mwe_flux_workspace_bytes": (_sz, [_i64]),
    "mwe_flux_accumulate_f64": (_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _i64, _int, _p, _i64, _p, _p, _p, _p, _p, _p, _sz, _p, _p]),
    "mwe_lineage_colour": (_int, [_p, _p, _i64, _p, _i64, _p, _i64, _p, _p]),
    "mwe_lineage_leaves": (_int, [_p, _p, _i64, _p, _i64, _p]),
    "mwe_lineage_records": (_int, [_p, _p, _i64, _p, _p, _p, _p, _i64, _p, _p, _p, _p, _p, _p]),
    "mwe_divide_f64": (_int, [_p, _i64, _f64, _p]),
    "mwe_hotpath_workspace_bytes": (_sz, [_i64, _i32]),
    "mwe_hotpath_workspace_bytes_ex": (_sz, [_i64, _i32, _int, _i32, _int]),
    "mwe_hotpath_step_f64": (_int, [_p, _i64, _int, _p, _int, _p, _i64, _p, _i64, _int, _p, _p, _i32, _p, _p, _p, _p, _p, _p,
                                    _i32, _int, _i64, _f64, _p, _p, _p, _sz, _p, _p]),
    "mwe_debug_set_tc_scores": (_int, [_p]),
    "mwe_debug_tc_columns": (_int, [_i32]),
    "mwe_debug_set_k1_profile": (_int, [_p]),
    "mwe_debug_set_tc_profile": (_int, [_p]),
    "mwe_sort_workspace_bytes": (_sz, [_i64]),
    "mwe_sort_pairs_u64_u32": (_int, [_p, _p, _i64, _int, _p, _sz, _p]),
}


class MweError(RuntimeError):
    """A C-ABI call returned a negative status."""


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built. There is no CPU fallback; run "
            "`python -c \"import __graft_entry__ as g; g.build()\"` or `make -C msm_we_b200/csrc`."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: intended
        fn.restype = res
        fn.argtypes = args
    got = lib.mwe_abi_version()
    if got != ABI_VERSION:
        raise ImportError(f"libmsm_we_b200.so ABI version {got} != expected {ABI_VERSION}; rebuild")
    return lib


lib = _load()


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.mwe_last_error()
        raise MweError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")


def set_timing_events(start, stop) -> None:
    """Ask the library to record these torch.cuda.Event objects around its dominant kernel (K1) in the
    calls that follow on this thread; ``None, None`` switches it off."""
    def handle(ev):
        if ev is None:
            return None
        if not ev.cuda_event:
            ev.record()          # torch creates the CUDA event lazily; the library re-records it
        return ev.cuda_event
    check(lib.mwe_set_timing_events(handle(start), handle(stop)), "mwe_set_timing_events")
