"""CPU: pins the oracle (oracle/gcp_oracle.c) to the reference's own known-answer vectors and to an
independent pure-Python statement of the same semantics.  The oracle is the checker for every GPU
parity test, so it is tested first."""
import numpy as np
import pytest


def test_kat1_forward_and_backward(oracle, golden):
    k = golden["kat1"]
    x, key, g = np.float32(k["x"]), np.int32(k["key"]), np.float32(k["grad_out"])
    seg_end = np.int32(k["seg_end"])
    y64 = oracle.cumprod_fwd(x, key)
    y32 = oracle.cumprod_fwd(x, key, np.float32)
    assert np.allclose(y64, k["y"], rtol=1e-6, atol=1e-7)
    assert np.allclose(y32, k["y"], rtol=1e-6, atol=1e-7)
    # the reference formula, literally (fp32 loop) and in fp64, and the exact division-free gradient
    gl = oracle.cumprod_bwd_refloop_f32(x, y32, g, key, seg_end)
    gr = oracle.cumprod_bwd_ref(x, y32, g, key, seg_end)
    ge = oracle.cumprod_bwd_exact(x, g, key)
    for got in (gl, gr, ge):
        assert np.allclose(got, k["grad_in"], rtol=1e-5, atol=1e-6), got


def test_kat2_unsorted_groups(oracle, golden):
    k = golden["kat2"]
    A = np.float32(k["A"])
    G = np.int64(k["G"])
    key = (G[:, 0] * (G.max() + 1) + G[:, 1]).astype(np.int32)
    order = np.argsort(key, kind="stable")  # the call site sorts first (gs_model.py:547)
    y_sorted = oracle.cumprod_fwd(A[order], key[order])
    y = np.empty_like(y_sorted)
    y[order] = y_sorted
    assert np.array_equal(y, np.float64(k["expected"]))


def test_adjacent_run_semantics_not_global_grouping(oracle):
    # thrust::inclusive_scan_by_key restarts on every adjacent change; a repeated key later is a NEW segment
    x = np.float32([2, 3, 5, 7, 11])
    key = np.int32([4, 4, 9, 4, 4])
    assert np.array_equal(oracle.cumprod_fwd(x, key), [2, 6, 5, 7, 77])
    assert np.array_equal(oracle.cumsum_fwd(x, key), [2, 5, 5, 7, 18])


@pytest.mark.parametrize("seed", range(6))
def test_c_oracle_matches_pure_python(oracle, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 200))
    L = rng.integers(1, 12, size=n)
    key = np.repeat(rng.integers(0, 5, size=n), L)[:n].astype(np.int32)
    x = rng.uniform(0.1, 1.5, n).astype(np.float32)
    if seed % 2:
        x[rng.integers(0, n, 3)] = 0.0
    g = rng.normal(size=n).astype(np.float32)
    assert np.allclose(oracle.cumprod_fwd(x, key), oracle.py_scan(x, key, "mul"), rtol=1e-12, atol=0)
    assert np.allclose(oracle.cumsum_fwd(x, key), oracle.py_scan(x, key, "add"), rtol=1e-12, atol=1e-12)
    inv = np.cumsum(np.r_[0, key[1:] != key[:-1]]).astype(np.int32)
    brute = oracle.py_bwd_bruteforce(x, g, inv)
    assert np.allclose(oracle.cumprod_bwd_exact(x, g, inv), brute, rtol=1e-10, atol=1e-12)


def test_reference_formula_equals_exact_gradient_without_zeros(oracle):
    rng = np.random.default_rng(7)
    L = rng.integers(1, 40, size=300)
    inv = np.repeat(np.arange(300), L).astype(np.int32)
    seg_end = np.cumsum(L).astype(np.int32)
    n = inv.size
    x = rng.uniform(0.3, 1.0, n).astype(np.float32)
    g = rng.normal(size=n).astype(np.float32)
    y = oracle.cumprod_fwd(x, inv, np.float32)
    ref = oracle.cumprod_bwd_ref(x, y, g, inv, seg_end)
    loop = oracle.cumprod_bwd_refloop_f32(x, y, g, inv, seg_end)
    exact = oracle.cumprod_bwd_exact(x, g, inv)
    assert oracle.allclose(ref, exact, rtol=1e-5, atol=1e-6)
    assert oracle.allclose(loop, exact, rtol=2e-4, atol=1e-5)  # fp32 sequential accumulation


def test_reference_returns_zero_at_zero_x_but_exact_does_not(oracle):
    # x=[.5,0,.5], g=1 -> exact [1,.75,0]; reference formula -> [1,0,0]
    x = np.float32([0.5, 0.0, 0.5])
    g = np.float32([1, 1, 1])
    inv = np.int32([0, 0, 0])
    y = oracle.cumprod_fwd(x, inv, np.float32)
    assert np.allclose(oracle.cumprod_bwd_exact(x, g, inv), [1.0, 0.75, 0.0])
    assert np.allclose(oracle.cumprod_bwd_ref(x, y, g, inv, np.int32([3])), [1.0, 0.0, 0.0])


def test_empty_and_single(oracle):
    e = np.float32([])
    assert oracle.cumprod_fwd(e, np.int32([])).size == 0
    assert oracle.cumprod_bwd_exact(e, e, np.int32([])).size == 0
    assert np.array_equal(oracle.cumprod_fwd(np.float32([3]), np.int32([5])), [3.0])
    assert np.array_equal(oracle.cumprod_bwd_exact(np.float32([3]), np.float32([2]), np.int32([0])), [2.0])


def test_omp_port_matches_sequential(oracle):
    rng = np.random.default_rng(3)
    L = rng.integers(1, 60, size=2000)
    key = np.repeat(np.arange(2000), L).astype(np.int32)
    n = key.size
    x = rng.uniform(0.2, 1.0, n).astype(np.float32)
    g = rng.normal(size=n).astype(np.float32)
    starts = oracle.segment_starts(key)
    assert starts[0] == 0 and starts[-1] == n and len(starts) == 2001
    y, gin = oracle.fwd_bwd_f32_omp(x, g, starts)
    assert oracle.allclose(y, oracle.cumprod_fwd(x, key))
    assert oracle.allclose(gin, oracle.cumprod_bwd_exact(x, g, key))


def test_oracle_matches_reference_ops_fixture(oracle):
    """tests/golden/ref_ops_fixture.npz = outputs of the reference's own CUDA ops (built unchanged by
    oracle/build_ref.sh) on a B200, generated by tests/golden/make_ref_fixture.py."""
    import os

    f = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_ops_fixture.npz"))
    x, g, key, inv, se = f["x"], f["g"], f["key"], f["inv"], f["seg_end"]
    assert oracle.allclose(f["y"], oracle.cumprod_fwd(x, key))
    assert oracle.allclose(f["cumsum"], oracle.cumsum_fwd(g, key))
    assert oracle.allclose(f["grad_in"], oracle.cumprod_bwd_ref(x, f["y"], g, inv, se))
    assert oracle.allclose(f["grad_in"], oracle.cumprod_bwd_exact(x, g, inv))   # zero-free input
    # integer side of the fixture is self-consistent, bit-exact
    starts = oracle.segment_starts(key)
    assert np.array_equal(starts[1:], se.astype(np.int64))
    assert np.array_equal(np.cumsum(np.r_[0, key[1:] != key[:-1]]).astype(np.int32), inv)
