"""ctypes binding of the C ABI declared in include/gcp_abi.h.

The library is the product: if it is missing or does not load, importing the ops
fails loudly.  There is no CPU fallback and no alternative backend.
"""
from __future__ import annotations

import ctypes
import os

from . import build as _build

_lib = None

GCP_OK = 0
ABI_VERSION = 2   # GCP_ABI_VERSION of include/gcp_abi.h this binding was written against
ERRORS = {
    -1: "GCP_ERR_INVALID_ARG",
    -2: "GCP_ERR_WORKSPACE",
    -3: "GCP_ERR_WATCHDOG",
    -4: "GCP_ERR_SEGMENTS",
}

# every symbol include/gcp_abi.h declares (tests check the .so exports all of them)
SYMBOLS = (
    "gcp_abi_version", "gcp_workspace_bytes", "gcp_workspace_init", "gcp_workspace_status", "gcp_workspace_attach_flag", "gcp_workspace_selftest_abort",
    "gcp_cumprod_fwd_f32", "gcp_cumsum_fwd_f32", "gcp_cumprod_bwd_f32", "gcp_validate_segments",
    "gcp_set_variant", "gcp_set_option", "gcp_num_variants", "gcp_variant_name", "gcp_last_launch_count",
    "gcp_splat_expand", "gcp_splat_sort_bytes", "gcp_splat_sort", "gcp_splat_prepare_bytes", "gcp_splat_prepare", "gcp_splat_pack", "gcp_splat_alpha", "gcp_splat_color",
    "gcp_splat_bwd_w", "gcp_splat_bwd_grads", "gcp_splat_bwd_elem", "gcp_splat_bwd_reduce", "gcp_splat_bwd_reduce_bytes",
    "gcp_splat_place_bytes", "gcp_splat_place", "gcp_splat_set_fill_blocks", "gcp_splat_set_long_list_threshold", "gcp_splat_seg_shift",
    "gcp_splat_num_cells", "gcp_splat_long_lists", "gcp_splat_bwd_elem_cells", "gcp_splat_batch_table_ints",
    "gcp_tile_width", "gcp_tile_height", "gcp_tile_num_tiles", "gcp_tile_set_piece_pairs", "gcp_tile_piece_pairs",
    "gcp_view_plan_bytes", "gcp_view_pair_bytes", "gcp_view_layout", "gcp_view_plan", "gcp_view_render", "gcp_view_forward",
    "gcp_view_backward", "gcp_view_backward_scatter", "gcp_view_last_launch_count",
    "gcp_views_ctx_create", "gcp_views_ctx_destroy", "gcp_views_step", "gcp_views_step_split",
    "gcp_host_boundary_bits", "gcp_ids_from_bits_bytes", "gcp_ids_from_bits",
)


class ViewDesc(ctypes.Structure):
    """gcp_view_desc of include/gcp_abi.h: one view of a gcp_views_step batch."""
    _fields_ = [("sp", ctypes.c_void_p), ("ep", ctypes.c_void_p), ("mean", ctypes.c_void_p), ("lam", ctypes.c_void_p),
                ("opac", ctypes.c_void_p), ("l_d", ctypes.c_void_p), ("index", ctypes.c_void_p),
                ("target", ctypes.c_void_p), ("grad_image", ctypes.c_void_p), ("image", ctypes.c_void_p),
                ("n", ctypes.c_int64)]


class ViewsSplit(ctypes.Structure):
    """gcp_views_split of include/gcp_abi.h: the tail of a batch and the event in front of it."""
    _fields_ = [("first_tail_view", ctypes.c_int), ("g_mean", ctypes.c_void_p), ("g_lam", ctypes.c_void_p),
                ("g_opac", ctypes.c_void_p), ("g_l", ctypes.c_void_p), ("event", ctypes.c_void_p)]


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    # incremental: recompiles only translation units older than their sources, a no-op when everything is current and
    # where there is no nvcc (the GPU box uses the library that travelled with the tree).  Never a stale binary
    # behind fresh ctypes signatures.
    try:
        _build.build_lib()
    except Exception as e:  # noqa: BLE001
        raise ImportError(
            f"libgcp_b200.so is not built ({e}); run `python -c 'import __graft_entry__ as g; g.build()'`"
        ) from e
    # the library's only OpenMP region is the host-side key packing (gcp_host.cu): its idle worker threads should
    # sleep, not spin, between the chunks of a step (they share the cores with the other ranks of the node)
    os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
    L = ctypes.CDLL(path)
    vp, i64, sz, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_size_t, ctypes.c_int
    missing = [n for n in SYMBOLS if not hasattr(L, n)]
    if missing:
        raise ImportError(f"{path} does not export {missing}: stale build, run __graft_entry__.build()")
    L.gcp_abi_version.restype = ci
    if L.gcp_abi_version() != ABI_VERSION:
        raise ImportError(f"{path} has ABI version {L.gcp_abi_version()}, this binding expects {ABI_VERSION}: "
                          "stale build, run __graft_entry__.build()")
    L.gcp_workspace_attach_flag.argtypes = [vp, sz, vp, vp]
    L.gcp_workspace_attach_flag.restype = ci
    L.gcp_workspace_selftest_abort.argtypes = [vp, sz, vp]
    L.gcp_workspace_selftest_abort.restype = ci
    L.gcp_workspace_bytes.argtypes = [i64]
    L.gcp_workspace_bytes.restype = sz
    L.gcp_workspace_init.argtypes = [vp, sz, vp]
    L.gcp_workspace_status.argtypes = [vp, vp, ctypes.POINTER(ci)]
    L.gcp_cumprod_fwd_f32.argtypes = [vp, vp, vp, i64, vp, sz, vp]
    L.gcp_cumsum_fwd_f32.argtypes = [vp, vp, vp, i64, vp, sz, vp]
    L.gcp_cumprod_bwd_f32.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, vp, sz, vp]
    L.gcp_validate_segments.argtypes = [vp, vp, i64, i64, vp, sz, vp, ctypes.POINTER(i64)]
    L.gcp_set_variant.argtypes = [ci, ci]
    L.gcp_set_option.argtypes = [ci, ci]
    L.gcp_set_option.restype = ci
    L.gcp_num_variants.argtypes = [ci]
    L.gcp_variant_name.argtypes = [ci, ci]
    L.gcp_variant_name.restype = ctypes.c_char_p
    L.gcp_last_launch_count.restype = ci
    L.gcp_splat_expand.argtypes = [vp, vp, vp, i64, i64, vp, vp, vp]
    L.gcp_splat_sort_bytes.argtypes = [i64]
    L.gcp_splat_sort_bytes.restype = sz
    L.gcp_splat_sort.argtypes = [vp, vp, vp, vp, i64, ci, vp, sz, vp]
    L.gcp_splat_prepare_bytes.argtypes = [i64]
    L.gcp_splat_prepare_bytes.restype = sz
    L.gcp_splat_prepare.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp, sz, vp]
    L.gcp_splat_prepare.restype = ci
    L.gcp_splat_pack.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp]
    L.gcp_splat_pack.restype = ci
    L.gcp_splat_alpha.argtypes = [vp, vp, vp, i64, vp, vp]
    L.gcp_splat_color.argtypes = [vp, vp, vp, vp, vp, i64, ci, vp, vp]
    L.gcp_splat_bwd_w.argtypes = [vp, vp, vp, vp, vp, vp, i64, ci, vp, vp]
    L.gcp_splat_bwd_grads.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, ci, vp, vp, vp, vp, vp]
    L.gcp_splat_place_bytes.argtypes = [i64, ci, ci]
    L.gcp_splat_place_bytes.restype = sz
    L.gcp_splat_place.argtypes = [vp, vp, vp, i64, i64, ci, ci, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    L.gcp_splat_batch_table_ints.argtypes = [i64, ci, ci]
    L.gcp_splat_batch_table_ints.restype = i64
    L.gcp_splat_num_cells.argtypes = [ci, ci]
    L.gcp_splat_num_cells.restype = ci
    L.gcp_splat_long_lists.argtypes = [i64, ci, ci]
    L.gcp_splat_long_lists.restype = ci
    L.gcp_splat_bwd_elem_cells.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, ci, ci, vp, vp]
    L.gcp_splat_bwd_elem_cells.restype = ci
    L.gcp_splat_place.restype = ci
    L.gcp_splat_set_fill_blocks.argtypes = [ci]
    L.gcp_splat_set_fill_blocks.restype = ci
    L.gcp_splat_set_long_list_threshold.argtypes = [ci]
    L.gcp_splat_set_long_list_threshold.restype = ci
    L.gcp_splat_seg_shift.restype = ci
    L.gcp_splat_bwd_elem.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, ci, vp, vp]
    L.gcp_splat_bwd_reduce.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, vp, vp, vp, vp, vp, sz, vp]
    L.gcp_splat_bwd_reduce_bytes.argtypes = [i64, i64]
    L.gcp_splat_bwd_reduce_bytes.restype = sz
    L.gcp_host_boundary_bits.argtypes = [vp, i64, vp, ci]
    L.gcp_host_boundary_bits.restype = ci
    L.gcp_ids_from_bits_bytes.argtypes = [i64]
    L.gcp_ids_from_bits_bytes.restype = sz
    L.gcp_ids_from_bits.argtypes = [vp, i64, vp, vp, sz, vp]
    L.gcp_ids_from_bits.restype = ci
    L.gcp_tile_width.restype = ci
    L.gcp_tile_height.restype = ci
    L.gcp_tile_num_tiles.argtypes = [ci, ci]
    L.gcp_tile_num_tiles.restype = ci
    L.gcp_tile_set_piece_pairs.argtypes = [ci]
    L.gcp_tile_set_piece_pairs.restype = ci
    L.gcp_tile_piece_pairs.restype = ci
    L.gcp_view_plan_bytes.argtypes = [i64, ci, ci]
    L.gcp_view_plan_bytes.restype = sz
    L.gcp_view_pair_bytes.argtypes = [i64, ci, ci]
    L.gcp_view_pair_bytes.restype = sz
    L.gcp_view_layout.argtypes = [i64, ci, ci, i64, ctypes.POINTER(i64)]
    L.gcp_view_plan.argtypes = [vp, vp, i64, ci, ci, vp, sz, vp, vp]
    L.gcp_view_render.argtypes = [vp, vp, vp, vp, vp, vp, i64, ci, ci, vp, sz, vp, sz, i64, ci, vp, vp]
    L.gcp_view_forward.argtypes = [vp, vp, vp, vp, vp, vp, i64, ci, ci, vp, sz, vp, sz, i64, ci, vp, vp, vp]
    L.gcp_view_backward.argtypes = [vp, sz, vp, sz, i64, vp, i64, ci, ci, vp, vp, vp, vp, vp]
    L.gcp_view_backward_scatter.argtypes = [vp, sz, vp, sz, i64, vp, i64, ci, ci, vp, vp, vp, vp, vp, vp]
    L.gcp_views_ctx_create.argtypes = [ci, ctypes.POINTER(vp)]
    L.gcp_views_ctx_destroy.argtypes = [vp]
    L.gcp_views_ctx_destroy.restype = None
    L.gcp_views_step.argtypes = [vp, ctypes.POINTER(ViewDesc), ci, ci, ci, ctypes.POINTER(vp), sz, ctypes.POINTER(vp), sz,
                                 i64, vp, vp, vp, vp, vp, vp, vp]
    L.gcp_views_step_split.argtypes = [vp, ctypes.POINTER(ViewDesc), ci, ci, ci, ctypes.POINTER(vp), sz, ctypes.POINTER(vp),
                                       sz, i64, vp, vp, vp, vp, vp, vp, ctypes.POINTER(ViewsSplit), vp]
    for name in ("gcp_view_layout", "gcp_view_plan", "gcp_view_render", "gcp_view_forward", "gcp_view_backward",
                 "gcp_view_backward_scatter", "gcp_views_ctx_create", "gcp_views_step", "gcp_views_step_split",
                 "gcp_view_last_launch_count"):
        getattr(L, name).restype = ci
    for name in ("gcp_splat_expand", "gcp_splat_sort", "gcp_splat_alpha", "gcp_splat_color", "gcp_splat_bwd_w",
                 "gcp_splat_bwd_grads", "gcp_splat_bwd_elem", "gcp_splat_bwd_reduce"):
        getattr(L, name).restype = ci
    for name in ("gcp_workspace_init", "gcp_workspace_status", "gcp_cumprod_fwd_f32", "gcp_cumsum_fwd_f32",
                 "gcp_cumprod_bwd_f32", "gcp_validate_segments", "gcp_set_variant", "gcp_num_variants"):
        getattr(L, name).restype = ci
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc == GCP_OK:
        return
    if rc < 0:
        raise RuntimeError(f"{what}: {ERRORS.get(rc, rc)}")
    raise RuntimeError(f"{what}: CUDA error {rc}")
