"""CPU: the C-ABI library loads without a GPU and exports every symbol include/gcp_abi.h declares;
argument validation that needs no device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "gcp_abi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gcp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from simplegaussiansplat_tk71_b200 import _lib

    L = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/gcp_abi.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared
    assert L.gcp_abi_version() == _lib.ABI_VERSION == int(re.search(r"#define GCP_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "gcp_abi.h")).read()).group(1))


def test_workspace_bytes_is_monotone_and_small():
    from simplegaussiansplat_tk71_b200 import _lib

    L = _lib.lib()
    prev = 0
    for n in (0, 1, 1023, 1024, 1025, 1 << 20, 1 << 27, (1 << 31) - 1):
        b = L.gcp_workspace_bytes(n)
        assert b >= prev and b >= 256
        prev = b
    assert L.gcp_workspace_bytes((1 << 31) - 1) < 96 << 20  # 40 B per 1024 elements


def test_argument_validation_without_device():
    from simplegaussiansplat_tk71_b200 import _lib

    L = _lib.lib()
    assert L.gcp_cumprod_fwd_f32(None, None, None, -1, None, 0, None) == -1
    assert L.gcp_cumprod_fwd_f32(None, None, None, 0, None, 0, None) == 0     # n == 0 is a no-op
    assert L.gcp_cumsum_fwd_f32(None, None, None, 5, None, 0, None) == -1    # null pointers
    assert L.gcp_cumprod_bwd_f32(None, None, None, None, None, None, 0, 0, None, 0, None) == 0
    assert L.gcp_cumprod_bwd_f32(None, None, None, None, None, None, 4, 1, None, 0, None) == -1
    assert L.gcp_set_variant(0, 999) == -1 and L.gcp_set_variant(7, 0) == -1
    assert L.gcp_set_variant(0, -1) == 0
    assert L.gcp_num_variants(0) >= 2 and L.gcp_num_variants(1) >= 2
    assert L.gcp_variant_name(0, 0).decode().startswith("ldg")
    ws = (ctypes.c_char * 16)()
    # workspace too small for n
    fake = ctypes.cast(ws, ctypes.c_void_p)
    assert L.gcp_cumprod_fwd_f32(fake, fake, fake, 100, fake, 16, None) == -2


def test_ops_refuse_cpu_tensors_and_wrong_dtypes():
    import torch

    import grouped_cumprod as gc

    x = torch.ones(4)
    k = torch.zeros(4, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="no CPU path"):
        gc.grouped_cumprod_forward(x, k, torch.empty(4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        gc.grouped_cumsum_forward(x, k, torch.empty(4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        gc.grouped_cumprod_backward(x, x, x, k, torch.empty(4), torch.tensor([4], dtype=torch.int32))
    assert sorted(n for n in dir(gc) if n.startswith("grouped_")) == [
        "grouped_cumprod_backward", "grouped_cumprod_forward", "grouped_cumsum_forward"]


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "simplegaussiansplat_tk71_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "workloads.py" and False, f"{f} mentions the oracle"
    assert "oracle" not in open(os.path.join(ROOT, "grouped_cumprod.py")).read()


def test_splat_size_helpers_are_pure_host_arithmetic():
    """The sizing helpers of the compositor section run without a GPU (no compute): cells of one row x 2^S pixels,
    long-list selection by pairs per pixel, batch-table and reduce-scratch sizes."""
    from simplegaussiansplat_tk71_b200 import _lib

    L = _lib.lib()
    S = L.gcp_splat_seg_shift()
    assert S == 6
    W, H = 1920, 1080
    nseg = (W + (1 << S)) >> S
    assert L.gcp_splat_num_cells(W, H) == (H + 1) * nseg
    npix = (H + 1) * (W + 1)
    assert L.gcp_splat_long_lists(8 * npix, W, H) == 1 and L.gcp_splat_long_lists(8 * npix - 1, W, H) == 0
    P = 6_000_000
    assert L.gcp_splat_batch_table_ints(P, W, H) == ((P >> 5) + (H + 1) * nseg + 1) * (2 + (1 << S))
    assert L.gcp_splat_long_lists(-1, W, H) == 0 and L.gcp_splat_num_cells(-1, H) == 0
    # reduce scratch: header + two id arrays + one 32-byte partial per piece; grows with N
    a, b = L.gcp_splat_bwd_reduce_bytes(10_000_000, 100_000), L.gcp_splat_bwd_reduce_bytes(20_000_000, 100_000)
    assert 0 < a < b < 20_000_000 * 8
    try:
        assert L.gcp_splat_set_long_list_threshold(-1) == -1      # GCP_ERR_INVALID_ARG
    finally:
        L.gcp_splat_set_long_list_threshold(8)


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 63, 64, 65, 100_000, 1_048_577])
@pytest.mark.parametrize("threads", [1, 3])
def test_host_boundary_bits_match_numpy(n, threads):
    """The host half of the host-buffer entry point (no GPU involved): one run-start bit per element."""
    import numpy as np

    from simplegaussiansplat_tk71_b200 import _lib

    L = _lib.lib()
    rng = np.random.default_rng(n + threads)
    key = (np.cumsum(rng.uniform(size=n) < 0.1).astype(np.int32) * 3 + 5) if n else np.zeros(0, np.int32)
    key = np.concatenate([np.zeros(1, np.int32), key])[1:]          # an unaligned view of the data
    words = (n + 31) // 32
    bits = np.full(words + 1, 0xDEADBEEF, np.uint32)
    assert L.gcp_host_boundary_bits(key.ctypes.data, n, bits.ctypes.data, threads) == 0
    flags = np.ones(n, bool)
    flags[1:] = key[1:] != key[:-1]
    want = np.zeros(words, np.uint32)
    idx = np.flatnonzero(flags)
    np.bitwise_or.at(want, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
    assert np.array_equal(bits[:words], want)
    assert bits[words] == 0xDEADBEEF


def test_tile_size_helpers_are_pure_host_arithmetic():
    """Sizing helpers of the fused compositor route (csrc/gcp_tile.cu), no GPU involved: 8x4-pixel tiles over the
    inclusive pixel grid [0,W] x [0,H]; work units (pieces) of at most gcp_tile_piece_pairs() pairs."""
    from simplegaussiansplat_tk71_b200 import _lib

    L = _lib.lib()
    tw, th = L.gcp_tile_width(), L.gcp_tile_height()
    assert (tw, th) == (8, 4) and tw * th == 32
    assert L.gcp_tile_num_tiles(1920, 1080) == ((1920 + tw) // tw) * ((1080 + th) // th) == 241 * 271
    assert L.gcp_tile_num_tiles(7, 3) == 1 and L.gcp_tile_num_tiles(8, 4) == 4      # W+1 = 9, H+1 = 5 pixels
    assert L.gcp_tile_num_tiles(-1, 10) == 0 and L.gcp_tile_num_tiles(40000, 10) == 0
    piece = L.gcp_tile_piece_pairs()
    assert piece == 128
    P, W, H, n = 3_579_735, 1920, 1080, 905_932
    nt = L.gcp_tile_num_tiles(W, H)
    # plan arena: ~100 B per Gaussian (64 B record + counts, offsets, the big-Gaussian list) + ~24 B per tile
    pb = L.gcp_view_plan_bytes(n, W, H)
    assert 76 * n + 20 * nt < pb < 96 * n + 32 * nt + (1 << 16)
    # pair arena: 4 (pair list) + 16 (T checkpoints) + 32 (gradient partials) + 12 (piece state) bytes per pair
    # + one checkpoint row (128 B) per tile
    qb = L.gcp_view_pair_bytes(P, W, H)
    assert 52 * P + 128 * nt < qb < 68 * P + 128 * nt + (1 << 16)
    assert L.gcp_view_plan_bytes(-1, W, H) == 0 and L.gcp_view_pair_bytes(-1, W, H) == 0
    assert L.gcp_view_plan_bytes(0, 8, 4) > 0 and L.gcp_view_pair_bytes(0, 8, 4) > 0
    import ctypes

    out = (ctypes.c_int64 * 16)()
    assert L.gcp_view_layout(n, W, H, P, out) == 0
    assert all(out[i] % 256 == 0 for i in range(8)) and out[8] == 2 * (P // piece) + 2
    try:
        for bad in (0, 31, 33, 100, 1 << 21):
            assert L.gcp_tile_set_piece_pairs(bad) == -1    # GCP_ERR_INVALID_ARG: a multiple of 32 in [32, 2**20]
        assert L.gcp_tile_set_piece_pairs(64) == 0 and L.gcp_view_layout(n, W, H, P, out) == 0
        assert out[8] == 2 * (P // 64) + 2
    finally:
        L.gcp_tile_set_piece_pairs(piece)


def test_view_entry_points_validate_their_arguments_without_device():
    """Every check of the view / batch entry points that precedes the first CUDA call: bad sizes and pointers give
    GCP_ERR_INVALID_ARG / GCP_ERR_WORKSPACE, never a crash (the reference's raw kernel launches are unchecked,
    grouped_cumprod_backward.cu:56-64)."""
    from simplegaussiansplat_tk71_b200 import _lib

    L = _lib.lib()
    INVALID, WORKSPACE = -1, -2
    assert (_lib.ERRORS[INVALID], _lib.ERRORS[WORKSPACE]) == ("GCP_ERR_INVALID_ARG", "GCP_ERR_WORKSPACE")
    buf = ctypes.create_string_buffer(1 << 16)
    base = (ctypes.addressof(buf) + 255) & ~255          # a 256-byte aligned fake arena (never dereferenced)
    W, H, n = 64, 32, 10
    need = L.gcp_view_plan_bytes(n, W, H)
    assert 0 < need < (1 << 16) - 256
    totals = (ctypes.c_int64 * 2)()
    # plan: negative n, image too large, no arena, misaligned arena, arena too small, misaligned boxes
    assert L.gcp_view_plan(base, base, -1, W, H, base, need, totals, None) == INVALID
    assert L.gcp_view_plan(base, base, n, 40000, H, base, need, totals, None) == INVALID
    assert L.gcp_view_plan(base, base, n, W, H, None, need, totals, None) == INVALID
    assert L.gcp_view_plan(base, base, n, W, H, base + 8, need, totals, None) == WORKSPACE
    assert L.gcp_view_plan(base, base, n, W, H, base, need - 1, totals, None) == WORKSPACE
    assert L.gcp_view_plan(base + 4, base, n, W, H, base, need, totals, None) == INVALID
    # render / backward: capacity out of range, missing tables, arenas too small
    pair_need = L.gcp_view_pair_bytes(1000, W, H)
    args = (base, base, base, base, base, base, n, W, H, base, need, base, pair_need, 1000, 1, base, None)
    assert L.gcp_view_render(*args[:13], 2 ** 31, *args[14:]) == INVALID
    assert L.gcp_view_render(*args[:2], None, *args[3:]) == INVALID
    assert L.gcp_view_render(*args[:12], pair_need - 1, *args[13:]) == WORKSPACE
    bw = (base, need, base, pair_need, 1000, base, n, W, H, base, base, base, base, None)
    assert L.gcp_view_backward(*bw[:5], None, *bw[6:]) == INVALID                     # no grad_image
    assert L.gcp_view_backward(*bw[:9], None, *bw[10:]) == INVALID                    # no output
    assert L.gcp_view_backward(*bw[:9], base + 4, *bw[10:]) == INVALID                # g_mean must be 8-byte aligned
    assert L.gcp_view_backward(*bw[:3], pair_need - 1, *bw[4:]) == WORKSPACE
    assert L.gcp_view_backward(*bw[:6], 0, *bw[7:]) == 0                              # an empty view: nothing to do
    # batch: lanes out of range, no context
    ctx = ctypes.c_void_p()
    assert L.gcp_views_ctx_create(0, ctypes.byref(ctx)) == INVALID
    assert L.gcp_views_ctx_create(5, ctypes.byref(ctx)) == INVALID
    assert L.gcp_views_ctx_create(1, None) == INVALID
    plans = (ctypes.c_void_p * 1)(base)
    assert L.gcp_views_step(None, None, 0, W, H, plans, need, plans, pair_need, 1000, base, base, base, base, None,
                            totals, None) == INVALID
    L.gcp_views_ctx_destroy(None)                                                       # a no-op


def test_native_view_batch_refuses_cpu_tensors_and_bad_arguments():
    import torch

    from simplegaussiansplat_tk71_b200 import workloads as wl
    from simplegaussiansplat_tk71_b200.views import NativeViewBatch

    v = wl.splat_view(64, 48, 200, seed=1, device="cpu")
    g = torch.zeros(49, 65, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        NativeViewBatch([v], 64, 48, grad_images=[g])
    with pytest.raises(ValueError, match="either targets"):
        NativeViewBatch([v], 64, 48)
    with pytest.raises(ValueError, match="either targets"):
        NativeViewBatch([v], 64, 48, targets=[g], grad_images=[g])
