"""-m gpu: parity of the CUDA path (through the drop-in `grouped_cumprod` module -> C ABI) against the
fp64 oracle, the golden vectors, and the reference's own CUDA ops (oracle/_ref) on identical inputs.

Tolerance (BASELINE.json north_star): |got - fp64| <= 1e-6 + 1e-5*|fp64|; integers bit-exact.
Every kernel variant is exercised, including the look-back across tiles, unaligned base pointers,
ragged tails, empty inputs and exact zeros.
"""
import numpy as np
import pytest
import torch

from gpu_util import assert_close, dev, reference_ops, run_bwd, run_fwd, seg_arrays

pytestmark = pytest.mark.gpu


def _variants(op):
    from simplegaussiansplat_tk71_b200 import ops

    return list(range(len(ops.variants(op))))


@pytest.fixture(params=[(1, 1, 0), (0, 0, 0), (1, 0, 0), (1, 2, 2), (0, 1, 1), (1, 1, 1)],
                ids=["default", "tickets-fixup", "tickets-halo", "chain-by-hint", "chain-nohalo", "chain-fwd-too"],
                autouse=True)
def carry_mode(request):
    """Every test runs in five carry modes (GCP_OPT_HALO, GCP_OPT_CHAIN): the default (every CTA of the blocked
    backward walks one contiguous tile range and hands the carries on in registers; halo window on the first tile
    of a range); tickets with the halo switched off, so that EVERY tile goes through the descriptor walk + fix-up;
    tickets with the halo; the mode picked from the workspace hint; chained ranges without any halo; chained
    ranges in the forward kernel as well (GCP_OPT_CHAIN_FWD, off by default)."""
    from simplegaussiansplat_tk71_b200 import ops

    halo, chain, chain_fwd = request.param
    ops.set_option(0, halo)
    ops.set_option(1, chain)
    ops.set_option(2, chain_fwd)
    yield request.param
    ops.set_option(0, 1)
    ops.set_option(1, 1)
    ops.set_option(2, 0)


FWD_VARIANTS = [0, 1]   # 0 = plain-load kernel pair (the fallback), 1 = persistent blocked kernel (the default)
BWD_VARIANTS = [0, 1]


def _pow2_values(n, rng, span=12):
    """x in {0.5, 1, 2} with the running exponent reflected inside +-span: the product of every CONTIGUOUS range
    is an exact power of two within 2**(+-2*span), so any association order gives bit-identical fp32 results.
    (Products of non-adjacent ranges are not bounded: a scan that commutes operands can overflow on this data —
    it did, in the descriptor walk, until its warp reduction was made order-preserving.)"""
    steps = rng.integers(-1, 2, n)
    e = 0
    out = np.empty(n, np.float32)
    for i in range(n):
        s_ = int(steps[i])
        if abs(e + s_) > span:
            s_ = -s_
        e += s_
        out[i] = 2.0 ** s_
    return out


def _values(n, rng, zeros=0):
    a = 1.0 / (1.0 + np.exp(-rng.normal(1.735, 1.432, n)))
    x = (1.0 - a * np.exp(-4.5 * rng.uniform(size=n))).astype(np.float32)
    if zeros:
        x[rng.integers(0, n, zeros)] = 0.0
    return x


def test_kat1_through_dropin_module(golden, oracle):
    import grouped_cumprod as gc

    k = golden["kat1"]
    param = torch.tensor(k["x"], device="cuda", dtype=torch.float32)
    grad = torch.clone(param)
    index = torch.tensor(k["key"], device="cuda", dtype=torch.int32)
    param_cumprod = torch.zeros_like(param)
    assert gc.grouped_cumprod_forward(param, index, param_cumprod) is None   # cuda_test.py:23
    out = torch.zeros_like(param)
    index_len = torch.tensor(k["seg_end"], device="cuda", dtype=torch.int32)
    assert gc.grouped_cumprod_backward(param, param_cumprod, grad, index, out, index_len) is None  # :29
    assert_close(param_cumprod.cpu().numpy(), k["y"], "KAT1 fwd")
    assert_close(out.cpu().numpy(), k["grad_in"], "KAT1 bwd")


def test_kat2_through_dropin_module(golden):
    import grouped_cumprod as gc

    k = golden["kat2"]
    A = torch.tensor(k["A"], dtype=torch.float32, device="cuda")
    G = torch.tensor(k["G"], device="cuda")
    key = (G[:, 0] * (int(G.max()) + 1) + G[:, 1]).to(torch.int32)
    skey, index = torch.sort(key, stable=True)
    out = torch.zeros_like(A)
    gc.grouped_cumprod_forward(A[index], skey, out)
    y = out[torch.argsort(index)]
    assert y.cpu().tolist() == [float(v) for v in k["expected"]]


@pytest.mark.parametrize("variant", FWD_VARIANTS)
@pytest.mark.parametrize("op", ["mul", "add"])
def test_forward_sizes_and_layouts(oracle, variant, op):
    rng = np.random.default_rng(100 + variant)
    ofn = oracle.cumprod_fwd if op == "mul" else oracle.cumsum_fwd
    sizes = [1, 2, 3, 4, 5, 31, 127, 128, 129, 255, 256, 257, 1023, 1024, 1025, 2047, 2049, 4095, 4096, 4097, 8191, 8193,
             12289, 40000, 100003]
    for n in sizes:
        x = _values(n, rng) if op == "mul" else rng.uniform(0.0, 2.0, n).astype(np.float32)
        for layout in ("random", "one", "singletons", "runs"):
            if layout == "random":
                L = np.maximum(1, np.rint(rng.lognormal(np.log(9), 1.0, n))).astype(np.int64)
                key = np.repeat(np.arange(len(L)), L)[:n]
                key = (key * 7 + 3).astype(np.int32)
            elif layout == "one":
                key = np.full(n, 42, np.int32)
            elif layout == "singletons":
                key = np.arange(n, dtype=np.int32)
            else:  # unsorted keys with repeated runs: adjacent-run semantics
                key = np.repeat(rng.integers(0, 3, n), rng.integers(1, 6, n))[:n].astype(np.int32)
            got = run_fwd(op, x, key, variant)
            assert_close(got, ofn(x, key), f"fwd {op} v{variant} n={n} {layout}")


@pytest.mark.parametrize("variant", FWD_VARIANTS)
def test_forward_boundaries_on_tile_edges_and_long_segments(oracle, variant):
    rng = np.random.default_rng(7)
    # segments that start/end exactly on 128 / 1024 / 2048 / 4096 / 8192 element edges, and around them
    for seglen in (128, 256, 255, 257, 512, 1024, 2048, 4096, 8192, 4095, 4097, 2049):
        n = seglen * 9 + 5
        key = (np.arange(n) // seglen).astype(np.int32)
        x = (1.0 - 1e-3 * rng.uniform(size=n)).astype(np.float32)
        assert_close(run_fwd("mul", x, key, variant), oracle.cumprod_fwd(x, key), f"edge {seglen} v{variant}")
        xs = rng.uniform(0, 1, n).astype(np.float32)
        assert_close(run_fwd("add", xs, key, variant), oracle.cumsum_fwd(xs, key), f"edge add {seglen}")
    # one giant segment spanning > 64 tiles (multi-round descriptor walk) with a short prologue and epilogue.
    # (a) exactly representable data: the carry logic must be BIT-exact over 170 tiles
    n = 700_001
    key = np.zeros(n, np.int32)
    key[:300] = -5
    key[-7777:] = 9
    x = _pow2_values(n, rng)
    assert np.array_equal(run_fwd("mul", x, key, variant), oracle.cumprod_fwd(x, key).astype(np.float32))
    xi = rng.integers(0, 3, n).astype(np.float32)      # partial sums < 2^24: exact in fp32
    assert np.array_equal(run_fwd("add", xi, key, variant), oracle.cumsum_fwd(xi, key).astype(np.float32))
    # (b) realistic data: a 700k-long fp32 product carries ~1e-4 relative rounding error in ANY order
    #     (the sequential fp32 oracle shows it too), so the bound adds the sequential-fp32 error
    x = (1.0 - 2e-5 * rng.uniform(size=n)).astype(np.float32)
    ref = oracle.cumprod_fwd(x, key)
    seq = np.abs(oracle.cumprod_fwd(x, key, np.float32) - ref).max()
    got = run_fwd("mul", x, key, variant)
    assert np.all(np.abs(got - ref) <= 1e-6 + 1e-5 * np.abs(ref) + 8 * seq), np.abs(got - ref).max()


@pytest.mark.parametrize("variant", [0, 1])
def test_forward_unaligned_slices_and_zeros(oracle, variant):
    rng = np.random.default_rng(11)
    n = 50_001
    L = np.maximum(1, np.rint(rng.lognormal(np.log(16), 1.0, n))).astype(np.int64)
    key = np.repeat(np.arange(len(L)), L)[:n].astype(np.int32)
    x = _values(n, rng, zeros=25)
    ref = oracle.cumprod_fwd(x, key)
    for off in [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 2, 3), (1, 1, 1), (2, 2, 2), (3, 3, 3), (3, 3, 0), (4, 4, 4)]:
        assert_close(run_fwd("mul", x, key, variant, off), ref, f"unaligned {off} v{variant}")
    # exact zeros propagate exactly
    got = run_fwd("mul", x, key, variant)
    assert np.array_equal(got == 0.0, ref == 0.0)


def test_empty_inputs_are_noops():
    import grouped_cumprod as gc

    e = torch.empty(0, device="cuda")
    ei = torch.empty(0, device="cuda", dtype=torch.int32)
    gc.grouped_cumprod_forward(e, ei, torch.empty(0, device="cuda"))
    gc.grouped_cumsum_forward(e, ei, torch.empty(0, device="cuda"))
    gc.grouped_cumprod_backward(e, e, e, ei, torch.empty(0, device="cuda"), ei)
    torch.cuda.synchronize()


def test_dtype_and_shape_errors_match_reference_behaviour():
    import grouped_cumprod as gc

    x = torch.ones(8, device="cuda")
    k = torch.zeros(8, device="cuda", dtype=torch.int32)
    with pytest.raises(RuntimeError):  # reference: data_ptr<int>() on an int64 tensor raises
        gc.grouped_cumprod_forward(x, k.long(), torch.empty_like(x))
    with pytest.raises(RuntimeError):
        gc.grouped_cumprod_forward(x.double(), k, torch.empty_like(x))
    with pytest.raises(RuntimeError):
        gc.grouped_cumprod_forward(x, k[:4], torch.empty_like(x))


@pytest.mark.parametrize("variant", BWD_VARIANTS)
def test_backward_sizes_and_layouts(oracle, variant):
    rng = np.random.default_rng(200 + variant)
    sizes = [1, 2, 3, 5, 127, 128, 129, 255, 256, 257, 1023, 1024, 1025, 2047, 2049, 4095, 4096, 4097, 8193, 12289, 40000,
             100003]
    for n in sizes:
        for layout in ("random", "one", "singletons"):
            if layout == "random":
                L = np.maximum(1, np.rint(rng.lognormal(np.log(9), 1.0, n))).astype(np.int64)
                L = L[: np.searchsorted(np.cumsum(L), n) + 1]
                L[-1] -= L.sum() - n
            elif layout == "one":
                L = np.array([n])
            else:
                L = np.ones(n, np.int64)
            inv, seg_end = seg_arrays(L)
            x = _values(n, rng)
            g = rng.uniform(0.0, 1.0, n).astype(np.float32)
            y = oracle.cumprod_fwd(x, inv, np.float32)
            got = run_bwd(x, y, g, inv, seg_end, variant)
            assert_close(got, oracle.cumprod_bwd_exact(x, g, inv), f"bwd v{variant} n={n} {layout}")


@pytest.mark.parametrize("variant", BWD_VARIANTS)
def test_backward_tile_edges_long_segments_signed_grads(oracle, variant):
    rng = np.random.default_rng(13)
    for seglen in (128, 256, 255, 257, 1024, 2048, 4096, 8192, 4095, 4097):
        n = seglen * 7 + 3
        L = [seglen] * 7 + [3]
        inv, seg_end = seg_arrays(L)
        x = (1.0 - 1e-3 * rng.uniform(size=n)).astype(np.float32)
        g = rng.normal(size=n).astype(np.float32)
        y = oracle.cumprod_fwd(x, inv, np.float32)
        got = run_bwd(x, y, g, inv, seg_end, variant)
        # signed g: the honest bound scales with the condition of the sum, i.e. the gradient for |g|
        scale = oracle.cumprod_bwd_exact(x, np.abs(g), inv)
        assert_close(got, oracle.cumprod_bwd_exact(x, g, inv), f"bwd edge {seglen} v{variant}", scale=scale)
    # giant segment (descriptor walk over > 64 tiles) between a short prologue and epilogue;
    # x = 1 (exact products) and small-integer g: S_i are exact integers < 2^24 -> bit-exact
    L = [300, 700_001 - 300 - 7777, 7777]
    n = sum(L)
    inv, seg_end = seg_arrays(L)
    x = np.ones(n, np.float32)
    g = rng.integers(0, 3, n).astype(np.float32)
    y = oracle.cumprod_fwd(x, inv, np.float32)
    assert np.array_equal(run_bwd(x, y, g, inv, seg_end, variant),
                          oracle.cumprod_bwd_exact(x, g, inv).astype(np.float32))
    # decaying products: S_i sums ~1e4 significant terms; tolerance widened to the fp32 reach of that sum
    x = (1.0 - 2e-4 * rng.uniform(size=n)).astype(np.float32)
    g = rng.uniform(0, 1, n).astype(np.float32)
    y = oracle.cumprod_fwd(x, inv, np.float32)
    assert_close(run_bwd(x, y, g, inv, seg_end, variant), oracle.cumprod_bwd_exact(x, g, inv),
                 f"bwd giant v{variant}", rtol=2e-4)


@pytest.mark.parametrize("variant", [0, 1])
def test_backward_exact_at_zeros_and_unaligned(oracle, variant):
    rng = np.random.default_rng(17)
    n = 60_001
    L = np.maximum(1, np.rint(rng.lognormal(np.log(16), 1.0, n))).astype(np.int64)
    L = L[: np.searchsorted(np.cumsum(L), n) + 1]
    L[-1] -= L.sum() - n
    inv, seg_end = seg_arrays(L)
    x = _values(n, rng, zeros=40)
    g = rng.uniform(0.1, 1.0, n).astype(np.float32)
    y = oracle.cumprod_fwd(x, inv, np.float32)
    exact = oracle.cumprod_bwd_exact(x, g, inv)
    for off in (0, 1, 2, 3):
        assert_close(run_bwd(x, y, g, inv, seg_end, variant, off), exact, f"bwd zeros off={off} v{variant}")
    # SURVEY.md §3.6-4 example: reference formula gives [1,0,0], the true gradient is [1,.75,0]
    got = run_bwd(np.float32([.5, 0, .5]), np.float32([.5, 0, 0]), np.float32([1, 1, 1]), np.int32([0, 0, 0]),
                  np.int32([3]), variant)
    assert np.allclose(got, [1.0, 0.75, 0.0])


def test_inf_nan_do_not_leak_across_segments(oracle):
    n = 9000
    inv, seg_end = seg_arrays([3000, 3000, 3000])
    x = np.full(n, 0.999, np.float32)
    g = np.ones(n, np.float32)
    g[4000] = np.inf   # only segment 1 may be affected
    y = oracle.cumprod_fwd(x, inv, np.float32)
    got = run_bwd(x, y, g, inv, seg_end)
    assert np.all(np.isfinite(got[:3000])) and np.all(np.isfinite(got[6000:]))
    xf = x.copy()
    xf[4000] = np.nan
    gotf = run_fwd("mul", xf, inv)
    assert np.all(np.isfinite(gotf[:3000])) and np.all(np.isfinite(gotf[6000:]))


def test_against_reference_cuda_ops_on_identical_inputs(oracle):
    ref = reference_ops()
    if ref is None:
        pytest.skip("oracle/_ref/grouped_cumprod_ref.so not built")
    import grouped_cumprod as gc

    rng = np.random.default_rng(23)
    n = 300_000
    L = np.maximum(1, np.rint(rng.lognormal(np.log(8), 1.0, n))).astype(np.int64)
    L = L[: np.searchsorted(np.cumsum(L), n) + 1]
    L[-1] -= L.sum() - n
    inv_np, seg_end_np = seg_arrays(L)
    x = dev(_values(n, rng))                 # zero-free: the two backward formulas agree there
    g = dev(rng.uniform(0, 1, n).astype(np.float32))
    inv, seg_end = dev(inv_np), dev(seg_end_np)
    y_ref, y_our = torch.zeros_like(x), torch.zeros_like(x)
    ref.grouped_cumprod_forward(x, inv, y_ref)
    gc.grouped_cumprod_forward(x, inv, y_our)
    s_ref, s_our = torch.zeros_like(x), torch.zeros_like(x)
    ref.grouped_cumsum_forward(g, inv, s_ref)
    gc.grouped_cumsum_forward(g, inv, s_our)
    b_ref, b_our = torch.zeros_like(x), torch.zeros_like(x)
    ref.grouped_cumprod_backward(x, y_ref, g, inv, b_ref, seg_end)
    gc.grouped_cumprod_backward(x, y_our, g, inv, b_our, seg_end)
    torch.cuda.synchronize()
    assert_close(y_our.cpu().numpy(), y_ref.cpu().numpy(), "fwd ours vs reference op")
    assert_close(s_our.cpu().numpy(), s_ref.cpu().numpy(), "cumsum ours vs reference op")
    assert_close(b_our.cpu().numpy(), b_ref.cpu().numpy(), "bwd ours vs reference op", rtol=1e-4, atol=1e-5)
    # and the reference ops themselves against the oracle (pins the oracle to the real reference)
    xn, gn = x.cpu().numpy(), g.cpu().numpy()
    assert_close(y_ref.cpu().numpy(), oracle.cumprod_fwd(xn, inv_np), "reference fwd vs oracle")
    assert_close(s_ref.cpu().numpy(), oracle.cumsum_fwd(gn, inv_np), "reference cumsum vs oracle")
    assert_close(b_ref.cpu().numpy(), oracle.cumprod_bwd_ref(xn, y_ref.cpu().numpy(), gn, inv_np, seg_end_np),
                 "reference bwd vs oracle(ref formula)", rtol=1e-4, atol=1e-5)


def test_against_committed_reference_ops_fixture():
    """The reference's own CUDA ops' outputs on a B200 (tests/golden/ref_ops_fixture.npz)."""
    import os

    f = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_ops_fixture.npz"))
    x, g, key, inv, se = f["x"], f["g"], f["key"], f["inv"], f["seg_end"]
    for v in FWD_VARIANTS:
        assert_close(run_fwd("mul", x, key, v), f["y"], f"fixture fwd v{v}")
        assert_close(run_fwd("add", g, key, v), f["cumsum"], f"fixture cumsum v{v}")
    for v in BWD_VARIANTS:
        assert_close(run_bwd(x, f["y"], g, inv, se, v), f["grad_in"], f"fixture bwd v{v}", rtol=2e-5, atol=2e-6)


def test_validate_segments_bit_exact_contract():
    from simplegaussiansplat_tk71_b200 import ops

    inv_np, seg_end_np = seg_arrays([3, 1, 5000, 2, 9000])
    inv, seg_end = dev(inv_np), dev(seg_end_np)
    assert ops.validate_segments(inv, seg_end) == 0
    bad = seg_end.clone()
    bad[2] += 1
    assert ops.validate_segments(inv, bad) > 0
    inv2 = inv.clone()
    inv2[10] = 4
    assert ops.validate_segments(inv2, seg_end) > 0


def test_autograd_function_matches_torch_autograd():
    from simplegaussiansplat_tk71_b200 import grouped_cumprod as gcp_fn

    rng = np.random.default_rng(29)
    L = rng.integers(1, 50, size=400)
    inv_np, seg_end_np = seg_arrays(L)
    n = len(inv_np)
    x = dev(_values(n, rng)).requires_grad_(True)
    w = dev(rng.uniform(0, 1, n).astype(np.float32))
    y = gcp_fn(x, dev(inv_np), dev(seg_end_np))
    (y * w).sum().backward()
    # plain torch fp64 reference: per-segment cumprod with autograd
    xr = x.detach().double().cpu().requires_grad_(True)
    parts, s = [], 0
    for l in L:
        parts.append(torch.cumprod(xr[s:s + l], 0))
        s += l
    yr = torch.cat(parts)
    (yr * w.double().cpu()).sum().backward()
    assert_close(y.detach().cpu().numpy(), yr.detach().numpy(), "autograd fwd")
    assert_close(x.grad.cpu().numpy(), xr.grad.numpy(), "autograd bwd")


def test_host_streamer_matches_resident_ops(oracle):
    """Host-buffer entry point: chunks cut at segment boundaries, three streams, the keys crossing the link as one
    run-start bit per element and rebuilt on the device as dense segment ids."""
    from simplegaussiansplat_tk71_b200.host import HostStreamer, _cut_points

    rng = np.random.default_rng(31)
    n = 900_000
    L = np.maximum(1, np.rint(rng.lognormal(np.log(20), 1.0, n // 8))).astype(np.int64)
    L[1000] = 70_000                                        # one list longer than a chunk
    L = L[: np.searchsorted(np.cumsum(L), n) + 1]
    L[-1] -= L.sum() - n
    inv, _ = seg_arrays(L)
    key = (inv.astype(np.int64) * 3 + 7).astype(np.int32)
    x = _values(n, rng)
    g = rng.uniform(0, 1, n).astype(np.float32)
    hx, hk, hg = (torch.from_numpy(a).pin_memory() for a in (x, key, g))
    hy = torch.empty(n).pin_memory()
    hgin = torch.empty(n).pin_memory()
    st = HostStreamer("cuda", chunk_elems=100_000, depth=3)
    cuts = _cut_points(hk, 100_000)
    assert cuts[0] == 0 and cuts[-1] == n and len(cuts) > 5
    assert all(key[c] != key[c - 1] for c in cuts[1:-1])    # every cut is a segment boundary
    up, down = st.fwd_bwd(hx, hk, hg, hy, hgin)
    torch.cuda.synchronize()
    assert down == 8 * n and 8 * n + n // 8 <= up <= 8 * n + n // 8 + 4 * len(cuts)
    # twice more through the same buffers (the pinned bit buffers are reused round-robin)
    hy.zero_()
    hgin.zero_()
    st.fwd_bwd(hx, hk, hg, hy, hgin)
    torch.cuda.synchronize()
    assert_close(hy.numpy(), oracle.cumprod_fwd(x, key), "streamer fwd")
    assert_close(hgin.numpy(), oracle.cumprod_bwd_exact(x, g, inv), "streamer bwd")


@pytest.mark.parametrize("n", [1, 31, 32, 33, 4097, 1_000_003])
def test_ids_from_run_start_bits(n):
    """gcp_host_boundary_bits (host) + gcp_ids_from_bits (device) rebuild dense segment ids bit-exactly."""
    from simplegaussiansplat_tk71_b200 import _lib

    L = _lib.lib()
    rng = np.random.default_rng(n)
    key = np.cumsum(rng.uniform(size=n) < 0.07).astype(np.int32) * 5 - 11     # runs of equal keys, arbitrary labels
    bits = np.zeros((n + 31) // 32, np.uint32)
    assert L.gcp_host_boundary_bits(key.ctypes.data, n, bits.ctypes.data, 4) == 0
    dbits = torch.from_numpy(bits.view(np.int32)).cuda()
    ids = torch.full((n + 8,), -7, dtype=torch.int32, device="cuda")
    tmp = torch.empty(int(L.gcp_ids_from_bits_bytes(n)), dtype=torch.uint8, device="cuda")
    rc = L.gcp_ids_from_bits(dbits.data_ptr(), n, ids.data_ptr(), tmp.data_ptr(), tmp.numel(),
                             torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    want = np.concatenate([[0], np.cumsum(key[1:] != key[:-1])]).astype(np.int32)
    got = ids.cpu().numpy()
    assert np.array_equal(got[:n], want)
    assert (got[n:] == -7).all()            # nothing written past the end


@pytest.mark.parametrize("phase", [1, 2, 3])
def test_alignment_peel_of_the_blocked_kernels(oracle, phase):
    """Sliced tensors (all arrays in the same 16-byte phase, the reference's [cutting_number:] case) take the
    persistent TMA kernel with 1-3 phantom elements in front of element 0: same results as the aligned call up to
    fp32 association order (the tile grid is shifted by the phantom elements), nothing written in front of the outputs (run_fwd / run_bwd check the guard words), also when the first segment
    is longer than the first tile (tile 0 then goes through the backward fix-up with the phantom elements in it)."""
    from simplegaussiansplat_tk71_b200 import ops

    rng = np.random.default_rng(900 + phase)
    for first in (1, 7, 300, 5000, 20_000):
        n = 70_003
        L = np.maximum(1, np.rint(rng.lognormal(np.log(16), 1.0, n))).astype(np.int64)
        L[0] = first
        L = L[: np.searchsorted(np.cumsum(L), n) + 1]
        L[-1] -= L.sum() - n
        inv, seg_end = seg_arrays(L)
        x = _values(n, rng, zeros=5)
        # the long first segment holds powers of two: every contiguous product is exact in fp32 whatever the
        # association order, so a carry that skips or repeats ONE element shows bit for bit (a 5000-element product
        # of values within 1e-4 of 1 drifts by 2e-5 in any tree order: (1-a)(1-b) always rounds the +ab term away)
        x[:first] = _pow2_values(first, rng, span=6)
        g = rng.normal(0, 1, n).astype(np.float32)
        y64 = oracle.cumprod_fwd(x, inv)
        got = run_fwd("mul", x, inv, 1, (phase, phase, phase))
        assert ops.last_launch_count() == 1, "the sliced call did not take the persistent kernel"
        assert np.array_equal(got[:first].astype(np.float64), y64[:first]), f"peel fwd phase {phase} first {first}: carry"
        assert_close(got, y64, f"peel fwd phase {phase} first {first}")
        assert_close(run_fwd("add", g, inv, 1, (phase, phase, phase)), oracle.cumsum_fwd(g, inv),
                     f"peel cumsum phase {phase} first {first}", scale=oracle.cumsum_fwd(np.abs(g), inv))
        y = oracle.cumprod_fwd(x, inv, np.float32)
        scale = np.abs(oracle.cumprod_bwd_exact(x, np.abs(g), inv))
        gb = run_bwd(x, y, g, inv, seg_end, 1, phase)
        assert ops.last_launch_count() == 1
        assert_close(gb, oracle.cumprod_bwd_exact(x, g, inv), f"peel bwd phase {phase} first {first}", scale=scale)


def test_two_streams_run_the_persistent_kernels_concurrently(oracle):
    """Two persistent (cooperative) ops in flight on two streams at once, each with its own workspace: the grid
    barrier must not starve whatever way the CTAs of the two kernels interleave."""
    import grouped_cumprod as gc
    from simplegaussiansplat_tk71_b200 import ops

    rng = np.random.default_rng(77)
    n = 3_000_001
    L = np.maximum(1, np.rint(rng.lognormal(np.log(20), 1.0, n // 8))).astype(np.int64)
    L = L[: np.searchsorted(np.cumsum(L), n) + 1]
    L[-1] -= L.sum() - n
    inv_np, seg_end_np = seg_arrays(L)
    x_np = _values(n, rng)
    g_np = rng.uniform(0.1, 1.0, n).astype(np.float32)
    x, g, inv, se = dev(x_np), dev(g_np), dev(inv_np), dev(seg_end_np)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    ys = [torch.empty_like(x) for _ in streams]
    gs = [torch.empty_like(x) for _ in streams]
    torch.cuda.synchronize()
    for rep in range(20):
        for s_, y, gi in zip(streams, ys, gs):
            with torch.cuda.stream(s_):
                gc.grouped_cumprod_forward(x, inv, y)
                gc.grouped_cumprod_backward(x, y, g, inv, gi, se)
    torch.cuda.synchronize()
    for s_ in streams:
        with torch.cuda.stream(s_):
            assert ops.workspace_status() == 0, "watchdog fired with two streams in flight"
    yref = oracle.cumprod_fwd(x_np, inv_np)
    gref = oracle.cumprod_bwd_exact(x_np, g_np, inv_np)
    for y, gi in zip(ys, gs):
        assert_close(y.cpu().numpy(), yref, "two-stream fwd")
        assert_close(gi.cpu().numpy(), gref, "two-stream bwd")


def test_a_tripped_watchdog_fails_loudly_at_the_next_call():
    """The device-side abort signal reaches the pinned host word without a sync, the next op on that workspace
    raises GCP_ERR_WATCHDOG instead of launching, and the workspace is usable again afterwards."""
    import grouped_cumprod as gc
    from simplegaussiansplat_tk71_b200 import _lib, ops

    x = torch.full((10_000,), 0.5, device="cuda")
    k = torch.zeros(10_000, dtype=torch.int32, device="cuda")
    y = torch.empty_like(x)
    gc.grouped_cumprod_forward(x, k, y)
    dev_ = x.device
    ws, stream = ops._workspace(dev_, 10_000)
    _lib.check(_lib.lib().gcp_workspace_selftest_abort(ws.data_ptr(), ws.numel(), stream), "selftest")
    torch.cuda.synchronize()
    assert ws.flag_np[0] == 1, "the kernel's abort signal did not reach the pinned host word"
    with pytest.raises(RuntimeError, match="GCP_ERR_WATCHDOG"):
        gc.grouped_cumprod_forward(x, k, y)
    gc.grouped_cumprod_forward(x, k, y)          # re-initialised: works again
    torch.cuda.synchronize()
    assert ops.workspace_status() == 0
    assert float(y[3]) == 0.5 ** 4
