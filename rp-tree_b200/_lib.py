"""ctypes loader for librpforest.so -- the C ABI declared in include/rpforest.h.

The product path has no CPU fallback: if the shared library is missing or no CUDA device is usable,
creating an engine raises.  (Loading the library itself and calling the host-only entry points --
rpf_sample_hyperplanes, rpf_topology_plan, rpf_topology_plan_chunked, rpf_rptree_cfg -- works without a GPU.)
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "librpforest.so")

i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
f64p = C.POINTER(C.c_double)
H = C.c_void_p

# name -> (restype, argtypes); mirrors include/rpforest.h one to one
SIGNATURES = {
    "rpf_create": (C.c_int, [C.POINTER(H), C.c_int]),
    "rpf_create_multi": (C.c_int, [C.POINTER(H), i32p, C.c_int]),
    "rpf_comm_unique_id": (C.c_int, [C.c_void_p]),
    "rpf_comm_init_rank": (C.c_int, [H, C.c_int32, C.c_int32, C.c_void_p]),
    "rpf_num_gpus": (C.c_int, [H]),
    "rpf_destroy": (None, [H]),
    "rpf_last_error": (C.c_char_p, [H]),
    "rpf_abi_version": (C.c_int, []),
    "rpf_set_points": (C.c_int, [H, f64p, C.c_int64, C.c_int32]),
    "rpf_set_points_device": (C.c_int, [H, C.c_void_p, C.c_int64, C.c_int32]),
    "rpf_set_points_sparse": (C.c_int, [H, C.c_int64, C.c_int32, i64p, i32p, f64p]),
    "rpf_points_are_sparse": (C.c_int, [H]),
    "rpf_densify_rows": (C.c_int, [C.c_int64, C.c_int32, i64p, i32p, f64p, f64p, i32p]),
    "rpf_knn_s": (C.c_int, [H, f64p, i32p, C.c_int64, C.c_int32, C.c_int32, f64p, u32p, i32p]),
    "rpf_recall_s": (C.c_int, [H, f64p, i32p, C.c_int64, C.c_int32, f64p]),
    "rpf_brute_knn_s": (C.c_int, [H, f64p, i32p, C.c_int64, C.c_int32, f64p, u32p]),
    "rpf_set_hyperplanes": (C.c_int, [H, C.c_int32, C.c_int32, i64p, i32p, f64p]),
    "rpf_gen_hyperplanes": (C.c_int, [H, C.c_uint64, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int32, C.c_int32]),
    "rpf_hyperplane_nnz": (C.c_int64, [H]),
    "rpf_sample_hyperplanes": (C.c_int64, [C.c_uint64, C.c_int32, C.c_int32, C.c_double, C.c_int32, i64p, i32p, f64p]),
    "rpf_rptree_cfg": (None, [C.c_int64, C.c_int64, C.c_int64, i64p, i64p, f64p]),
    "rpf_get_hyperplanes": (C.c_int, [H, i64p, i32p, f64p]),
    "rpf_build": (C.c_int, [H, C.c_int32, C.c_int32]),
    "rpf_build_from_host": (C.c_int, [H, f64p, C.c_int64, C.c_int32, C.c_int32, C.c_int32]),
    "rpf_build_chunked": (C.c_int, [H, C.c_int32, C.c_int32, C.c_int64]),
    "rpf_insert_begin": (C.c_int, [H, C.c_int32, C.c_int32, C.c_int32]),
    "rpf_insert_chunk": (C.c_int, [H, f64p, C.c_int64]),
    "rpf_insert_end": (C.c_int, [H]),
    "rpf_num_nodes": (C.c_int64, [H]),
    "rpf_num_trees": (C.c_int32, [H]),
    "rpf_hyperplane_depth": (C.c_int32, [H]),
    "rpf_points_shape": (C.c_int, [H, i64p, i32p]),
    "rpf_topology": (C.c_int, [H, i64p, i32p, i64p, i64p]),
    "rpf_topology_plan": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32, i64p, i32p, i64p, i64p]),
    "rpf_topology_plan_chunked": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32, C.c_int64, i64p, i32p, i64p, i64p, i64p]),
    "rpf_points_lost": (C.c_int64, [H]),
    "rpf_leaf_order_exact": (C.c_int, [H]),
    "rpf_tree_export": (C.c_int, [H, C.c_int32, f64p, f64p, f64p, u32p]),
    "rpf_forest_export": (C.c_int, [H, f64p, f64p, f64p, u32p]),
    "rpf_set_export_sink": (C.c_int, [H, f64p, f64p, f64p, u32p]),
    "rpf_forest_save": (C.c_int, [H, C.c_char_p, C.c_int32]),
    "rpf_forest_load": (C.c_int, [H, C.c_char_p]),
    "rpf_candidates_count": (C.c_int, [H, f64p, C.c_int64, C.c_int32, i64p]),
    "rpf_candidates": (C.c_int, [H, f64p, C.c_int64, C.c_int32, i64p, u32p]),
    "rpf_knn": (C.c_int, [H, f64p, C.c_int64, C.c_int32, C.c_int32, f64p, u32p, i32p]),
    "rpf_knn_h_capacity": (C.c_int64, [H, C.c_int32]),
    "rpf_knn_h": (C.c_int, [H, f64p, i32p, C.c_int64, C.c_int32, C.c_int64, f64p, u32p, i32p]),
    "rpf_recall": (C.c_int, [H, f64p, C.c_int64, C.c_int32, f64p]),
    "rpf_brute_knn": (C.c_int, [H, f64p, C.c_int64, C.c_int32, f64p, u32p]),
    "rpf_merge_topk": (C.c_int, [H, C.c_int32, C.c_int64, C.c_int32, C.c_int32, f64p, u32p, i32p, f64p, u32p, i32p]),
    "rpf_knn_dev": (C.c_int, [H, f64p, i32p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rpf_merge_topk_dev": (C.c_int, [H, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, f64p, u32p, i32p]),
    "rpf_last_device_ms": (C.c_double, [H]),
    "rpf_set_profiling": (C.c_int, [H, C.c_int]),
    "rpf_get_profile": (C.c_int, [H, f64p, i64p, C.c_int]),
    "rpf_phase_name": (C.c_char_p, [C.c_int]),
    "rpf_launch_count": (C.c_int64, [H]),
    "rpf_set_bottom_cap": (C.c_int, [H, C.c_int32]),
    "rpf_set_option": (C.c_int, [H, C.c_char_p, C.c_int64]),
}

_LIB = None


class RPForestError(RuntimeError):
    pass


def lib():
    """Load librpforest.so (raises if it has not been built: there is no fallback implementation)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise RPForestError(
                "librpforest.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "-- the engine has no CPU fallback." % SO_PATH)
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)        # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB
