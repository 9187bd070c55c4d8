"""Arrays longer than one call of the C ABI takes (GCP_MAX_ELEMENTS = 2**31 - 4, the reference's int indexing,
grouped_cumprod_backward.cu:52): ops.py cuts them at segment boundaries into independent calls — the counterpart of
the reference's chunk loop (gs_model.py:428, :675; row a6 of SURVEY.md §8) without its per-pixel carry.

CPU: the cut logic and the C ABI's refusal.  GPU (-m gpu): chunked calls against the oracle on a small limit, and
the real sizes — one call of exactly GCP_MAX_ELEMENTS elements, and an array above 2**31 — through size-independent
exact properties (positions inside the lists, element counts to the list end, exact powers of two)."""
import re

import numpy as np
import pytest
import torch


def _ids(lengths):
    return torch.from_numpy(np.repeat(np.arange(len(lengths), dtype=np.int32), np.asarray(lengths, np.int64)))


def test_segment_cuts_fall_on_boundaries_and_respect_the_limit():
    from simplegaussiansplat_tk71_b200 import ops

    rng = np.random.default_rng(3)
    L = np.maximum(1, np.rint(rng.lognormal(np.log(20), 1.0, 4000))).astype(np.int64)
    L[1234] = 900                                     # one list close to the limit
    ids = _ids(L)
    n = ids.numel()
    assert ops.segment_cuts(ids, n) == [0, n] and ops.segment_cuts(ids, 10 * n) == [0, n]
    for limit in (1000, 4097, 65536 + 7):
        cuts = ops.segment_cuts(ids, limit)
        assert cuts[0] == 0 and cuts[-1] == n and all(b > a for a, b in zip(cuts, cuts[1:]))
        ends = set(np.cumsum(L).tolist())
        for a, b in zip(cuts, cuts[1:]):
            assert b - a <= limit
            assert b in ends                          # every cut is the exclusive end of a list
        # greedy: a chunk could not have taken the next list as well
        starts = np.r_[0, np.cumsum(L)[:-1]]
        for a, b in zip(cuts[:-1], cuts[1:-1]):
            nxt = L[np.searchsorted(starts, b)]
            assert b - a + nxt > limit
    with pytest.raises(RuntimeError, match="cannot be cut"):
        ops.segment_cuts(ids, 899)                    # the 900-element list fits no chunk
    assert ops.segment_cuts(_ids([5]), 5) == [0, 5]
    assert ops.segment_cuts(torch.zeros(0, dtype=torch.int32), 5) == [0, 0]


def test_c_abi_refuses_more_than_max_elements_per_call():
    import os

    from simplegaussiansplat_tk71_b200 import _lib, ops

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    limit = int(re.search(r"#define GCP_MAX_ELEMENTS (\d+)LL", open(os.path.join(root, "include", "gcp_abi.h")).read()).group(1))
    assert limit == ops.MAX_ELEMENTS == 2 ** 31 - 4
    L = _lib.lib()
    assert L.gcp_cumprod_fwd_f32(None, None, None, limit + 1, None, 0, None) == -1     # GCP_ERR_INVALID_ARG
    assert L.gcp_cumsum_fwd_f32(None, None, None, 2 ** 31, None, 0, None) == -1
    assert L.gcp_cumprod_bwd_f32(None, None, None, None, None, None, limit + 1, 1, None, 0, None) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("limit", [5000, 70001])
def test_chunked_calls_match_the_oracle_gpu(oracle, monkeypatch, limit):
    """The chunked path on a small limit: against the fp64 oracle at the north star's tolerance, and bit for bit
    against the single call on exactly representable data."""
    import grouped_cumprod as gc
    from gpu_util import assert_close
    from simplegaussiansplat_tk71_b200 import ops

    rng = np.random.default_rng(11)
    L = np.maximum(1, np.rint(rng.lognormal(np.log(20), 1.0, 9000))).astype(np.int64)
    L[4000] = 4500                                     # spans a whole 4096-element tile of a chunk
    inv = np.repeat(np.arange(len(L), dtype=np.int32), L)
    n = inv.size
    x = (1.0 - 0.6 * rng.uniform(size=n) ** 4).astype(np.float32)
    g = rng.normal(size=n).astype(np.float32)
    p2 = (2.0 ** rng.integers(-1, 2, n)).astype(np.float32)
    seg_end = torch.from_numpy(np.cumsum(L).astype(np.int32)).cuda()
    dx, dg, dp, di = (torch.from_numpy(a).cuda() for a in (x, g, p2, inv))

    def run():
        y, s, gin, yp = (torch.full_like(dx, float("nan")) for _ in range(4))
        gc.grouped_cumprod_forward(dx, di, y)
        gc.grouped_cumsum_forward(dg, di, s)
        gc.grouped_cumprod_backward(dx, y, dg, di, gin, seg_end)
        gc.grouped_cumprod_forward(dp, di, yp)
        torch.cuda.synchronize()
        assert ops.workspace_status() == 0
        return y, s, gin, yp

    whole = run()
    monkeypatch.setattr(ops, "MAX_ELEMENTS", limit)
    assert len(ops.segment_cuts(di)) > 2
    y, s, gin, yp = run()
    assert_close(y.cpu().numpy(), oracle.cumprod_fwd(x, inv), "chunked fwd")
    assert_close(s.cpu().numpy(), oracle.cumsum_fwd(g, inv), "chunked cumsum", scale=oracle.cumsum_fwd(np.abs(g), inv))
    assert_close(gin.cpu().numpy(), oracle.cumprod_bwd_exact(x, g, inv), "chunked bwd",
                 scale=oracle.cumprod_bwd_exact(x, np.abs(g), inv))
    assert torch.equal(yp, whole[3])                   # powers of two: exact in any association order


@pytest.mark.gpu
def test_max_elements_in_one_call_and_more_than_2_31_chunked_gpu():
    """Maximum sizes.  (1) n = GCP_MAX_ELEMENTS = 2**31 - 4 in ONE call of each op (the last tile, the 64-bit
    element offsets and the tensor-map extents at the int32 edge); (2) n = 2**31 + 70 000, which no int-indexed
    op can take, through ops.py's cuts at segment boundaries.  Lists of 37 elements (the last one ragged); all
    checks are exact: cumsum of ones = position in the list, backward of ones = elements to the list end,
    cumprod of 2, 1/2, 2, ... = 2 or 1."""
    import grouped_cumprod as gc
    from simplegaussiansplat_tk71_b200 import ops

    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < 100 << 30:
        pytest.skip("needs 100 GB of device memory")
    n_big = 2 ** 31 + 70_000
    seg = 37
    idx = torch.arange(n_big, device="cuda", dtype=torch.int64)
    key_big = (idx // seg).to(torch.int32)
    pos_big = (idx % seg).to(torch.float32)
    del idx
    one_big = torch.ones(n_big, device="cuda")
    out_big = torch.empty(n_big, device="cuda")
    for n in (ops.MAX_ELEMENTS, n_big):
        key, pos, one, out = key_big[:n], pos_big[:n], one_big[:n], out_big[:n]
        length = torch.full((n,), float(seg), device="cuda")
        tail = n % seg
        if tail:
            length[n - tail:] = float(tail)
        out.fill_(float("nan"))
        gc.grouped_cumsum_forward(one, key, out)
        assert torch.equal(out, pos + 1)
        out.fill_(float("nan"))
        gc.grouped_cumprod_backward(one, one, one, key, out, torch.empty(0, dtype=torch.int32, device="cuda"))
        assert torch.equal(out, length - pos)
        del length
        even = torch.remainder(pos, 2.0) == 0
        xp = torch.where(even, 2.0, 0.5).to(torch.float32)
        out.fill_(float("nan"))
        gc.grouped_cumprod_forward(xp, key, out)
        del xp
        assert torch.equal(out, torch.where(even, 2.0, 1.0).to(torch.float32))
        del even
        torch.cuda.synchronize()
        assert ops.workspace_status() == 0
    assert len(ops.segment_cuts(key_big)) == 3


def test_segment_cuts_property_random_layouts():
    """Property test (hypothesis): for any segment layout and any limit that the longest segment fits, the cuts start
    at 0, end at n, increase strictly, fall on segment boundaries, respect the limit, and are greedy (the next segment
    would not have fitted) — so the number of calls is minimal for cuts at boundaries."""
    from hypothesis import given, settings, strategies as st

    from simplegaussiansplat_tk71_b200 import ops

    @settings(max_examples=150, deadline=None)
    @given(st.lists(st.integers(min_value=1, max_value=300), min_size=1, max_size=120), st.integers(0, 1000), st.booleans())
    def check(lengths, slack, sparse_ids):
        L = np.asarray(lengths, np.int64)
        ids = np.repeat(np.arange(len(L), dtype=np.int32) * (3 if sparse_ids else 1), L)   # ids need not be dense
        limit = int(L.max()) + slack
        cuts = ops.segment_cuts(torch.from_numpy(ids), limit)
        n = int(L.sum())
        ends = np.cumsum(L)
        assert cuts[0] == 0 and cuts[-1] == n
        assert all(b > a for a, b in zip(cuts, cuts[1:])) or n == 0
        starts = np.r_[0, ends[:-1]]
        for a, b in zip(cuts, cuts[1:]):
            assert b - a <= limit and b in set(ends.tolist())
            if b < n:
                assert b - a + int(L[np.searchsorted(starts, b)]) > limit

    check()
