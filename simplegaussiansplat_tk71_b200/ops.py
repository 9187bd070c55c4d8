"""Host-side mirror of the reference's `grouped_cumprod` extension ops.

Same names, positional signatures, in-place output convention and dtype rules as
/root/reference/cuda_kernel/cuda_kernel.cpp:5-22; the work is done by the hand-written
sm_100a kernels behind the C ABI (include/gcp_abi.h).  PyTorch is used only for device
memory, the current stream and error types.

Differences from the reference, all on the safe side:
  * inputs are checked (CUDA, dtype, 1-D contiguous, equal numel, same device) instead of
    being trusted; the reference only fails on dtype (data_ptr<T>()).
  * launches go to the *current* torch stream and never block the host (the reference's
    thrust calls run on the legacy stream and synchronise).
  * grouped_cumprod_backward is the exact division-free gradient: where an x is 0 the
    reference returns 0 (0/1e-8), this returns the true value.
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import os

import torch

from . import _lib

_workspaces: dict = {}
MAX_ELEMENTS = 2 ** 31 - 4   # GCP_MAX_ELEMENTS (include/gcp_abi.h): elements per call of the C ABI
_VALIDATE = os.environ.get("GCP_VALIDATE_SEGMENTS", "") not in ("", "0")


class _Workspace:
    """A self-resetting scan workspace (device) plus the pinned host word a tripped watchdog writes to."""
    __slots__ = ("buf", "flag", "flag_np", "stream")

    def __init__(self, device, nbytes: int, stream: int):
        L = _lib.lib()
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.flag = torch.zeros(1, dtype=torch.int32).pin_memory()   # cudaHostAlloc: device-visible at the same address
        self.flag_np = self.flag.numpy()                              # the per-call check is one host load
        self.stream = stream
        self.reset()

    def reset(self) -> None:
        L = _lib.lib()
        self.flag.zero_()
        _lib.check(L.gcp_workspace_init(self.buf.data_ptr(), self.buf.numel(), self.stream), "gcp_workspace_init")
        _lib.check(L.gcp_workspace_attach_flag(self.buf.data_ptr(), self.buf.numel(), self.flag.data_ptr(),
                                               self.stream), "gcp_workspace_attach_flag")

    def numel(self) -> int:
        return self.buf.numel()

    def data_ptr(self) -> int:
        return self.buf.data_ptr()


def _workspace(device: torch.device, n: int):
    """One cached, self-resetting workspace per (device, stream); grown geometrically.

    Fails loudly: if a kernel's bounded wait ever expired during an EARLIER op on this workspace, that op's results
    were invalid and the kernel said so in the attached pinned word (include/gcp_abi.h, gcp_workspace_attach_flag).
    The word is checked here, before every launch, with a plain host read: the workspace is re-initialised and
    GCP_ERR_WATCHDOG raised — never a silent wrong gradient, never a workspace that stays poisoned."""
    L = _lib.lib()
    stream = torch.cuda.current_stream(device)
    key = (device.index, stream.cuda_stream)
    need = int(L.gcp_workspace_bytes(n))
    ws = _workspaces.get(key)
    if ws is not None and ws.flag_np[0] != 0:
        torch.cuda.synchronize(device)
        ws.reset()
        raise RuntimeError("GCP_ERR_WATCHDOG: a bounded wait expired inside a scan kernel of an earlier call on this "
                           "stream; the results of that call are invalid (workspace re-initialised)")
    if ws is None or ws.numel() < need:
        if not _workspaces:
            _options_from_env()
        size = max(need, 1 << 20)
        if ws is not None:
            size = max(size, 2 * ws.numel())
        ws = _Workspace(device, size, stream.cuda_stream)
        _workspaces[key] = ws
    return ws, stream.cuda_stream


def _check(t: torch.Tensor, name: str, dtype: torch.dtype, n: int | None = None, device=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        # the reference raises here too: data_ptr<float>() / data_ptr<int>() (grouped_cumprod_forward.cu:8-10)
        raise RuntimeError(f"expected scalar type {dtype} for {name} but found {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if n is not None and t.numel() != n:
        raise RuntimeError(f"{name} has {t.numel()} elements, expected {n}")
    if device is not None and t.device != device:
        raise RuntimeError(f"{name} is on {t.device}, expected {device}")


def segment_cuts(ids: torch.Tensor, limit: int | None = None) -> list:
    """[0, c1, ..., n]: cut points of an array longer than one call of the C ABI takes, every inner one at a segment
    boundary (ids[c] != ids[c-1]) and at most `limit` elements after the one before.

    The counterpart of the reference's chunk loop (gs_model.py:428, :675, chunks of 2**29 elements made by
    uitility.py:478-488) for the int32-indexed ops: because no segment is split, the chunks are independent calls
    and nothing like the per-pixel carry of gs_model.py:582-594 is needed.  One small host sync per cut (the
    position of a boundary); arrays of up to `limit` elements return [0, n] without touching the device."""
    limit = MAX_ELEMENTS if limit is None else int(limit)
    n = ids.numel()
    cuts = [0]
    while n - cuts[-1] > limit:
        hi = cuts[-1] + limit            # the next cut is the last boundary at or below hi
        window, cut = 1 << 16, None
        while cut is None:
            lo = max(cuts[-1] + 1, hi - window)
            w = ids[lo - 1:hi + 1]
            at = (w[1:] != w[:-1]).nonzero()
            if at.numel():
                cut = lo + int(at[-1])
            elif lo == cuts[-1] + 1:
                raise RuntimeError(f"a segment of more than {limit} elements cannot be cut at a segment boundary")
            window *= 16
        cuts.append(cut)
    cuts.append(n)
    return cuts


def _scan(fn_name: str, x: torch.Tensor, key: torch.Tensor, y: torch.Tensor) -> None:
    _check(x, "x", torch.float32)
    n = x.numel()
    _check(key, "key", torch.int32, n, x.device)
    _check(y, "y", torch.float32, n, x.device)
    if n == 0:
        return
    with torch.cuda.device(x.device):
        cuts = segment_cuts(key) if n > MAX_ELEMENTS else [0, n]
        ws, stream = _workspace(x.device, max(b - a for a, b in zip(cuts, cuts[1:])))
        fn = getattr(_lib.lib(), fn_name)
        for a, b in zip(cuts, cuts[1:]):
            _lib.check(fn(x.data_ptr() + 4 * a, key.data_ptr() + 4 * a, y.data_ptr() + 4 * a, b - a, ws.data_ptr(),
                          ws.numel(), stream), fn_name)


def grouped_cumprod_forward(x: torch.Tensor, key: torch.Tensor, y: torch.Tensor) -> None:
    """y[i] = x[i] at a segment head, else y[i-1]*x[i]; written into `y`.  Returns None.

    Reference: cuda_kernel/grouped_cumprod_forward.cu:6-24, called at gs_model.py:551.
    """
    _scan("gcp_cumprod_fwd_f32", x, key, y)


def grouped_cumsum_forward(x: torch.Tensor, key: torch.Tensor, y: torch.Tensor) -> None:
    """Segmented inclusive sum, same contract.  Reference: grouped_cumsum_forward.cu:6-24 (gs_model.py:553)."""
    _scan("gcp_cumsum_fwd_f32", x, key, y)


def grouped_cumprod_backward(param: torch.Tensor, param_cumprod: torch.Tensor, grad_out: torch.Tensor,
                             inv: torch.Tensor, grad_in: torch.Tensor, inv_len: torch.Tensor) -> None:
    """grad_in[i] = dL/dparam[i] for L = sum grad_out*cumprod; written into `grad_in`.

    Reference: cuda_kernel/grouped_cumprod_backward.cu:43-65 (cuda_test.py:29).  `inv` are dense
    segment ids, `inv_len` the exclusive end offset of each segment (cuda_test.py:27).

    Contract on the segment layout.  The reference sums element i up to `inv_len[inv[i]]`; this op derives the
    segment tails from `inv` alone (a tail is where inv[i+1] != inv[i]) and reads neither `inv_len` nor its length,
    which is the same thing exactly when the two arguments are consistent: ids non-decreasing and dense,
    inv_len[s] = 1 + last index of s.  Inconsistent pairs (or ids that repeat non-adjacently) give a different
    result from the reference's without an error.  Set GCP_VALIDATE_SEGMENTS=1 in the environment to have every
    call with a non-empty `inv_len` checked first (`validate_segments`, one extra pass and a host sync) and
    raise GCP_ERR_SEGMENTS on a mismatch.  An empty `inv_len` is accepted: the compositor passes the sorted pixel
    keys as `inv`, for which no offsets exist.

    More than MAX_ELEMENTS = 2**31 - 4 elements (beyond what the reference's int-indexed kernel can address) are
    served as independent calls on chunks cut at segment boundaries (`segment_cuts`).
    """
    _check(param, "param", torch.float32)
    n = param.numel()
    dev = param.device
    _check(param_cumprod, "param_cumprod", torch.float32, n, dev)
    _check(grad_out, "grad_out", torch.float32, n, dev)
    _check(inv, "inv", torch.int32, n, dev)
    _check(grad_in, "grad_in", torch.float32, n, dev)
    _check(inv_len, "inv_len", torch.int32, None, dev)
    if n == 0:
        return
    if _VALIDATE and inv_len.numel() > 0:
        if n > MAX_ELEMENTS:
            raise RuntimeError("GCP_VALIDATE_SEGMENTS checks arrays of up to 2**31 - 4 elements")
        bad = validate_segments(inv, inv_len)
        if bad:
            raise RuntimeError(f"GCP_ERR_SEGMENTS: inv / inv_len are inconsistent at {bad} positions")
    with torch.cuda.device(dev):
        cuts = segment_cuts(inv) if n > MAX_ELEMENTS else [0, n]
        ws, stream = _workspace(dev, max(b - a for a, b in zip(cuts, cuts[1:])))
        L = _lib.lib()
        for a, b in zip(cuts, cuts[1:]):
            _lib.check(L.gcp_cumprod_bwd_f32(param.data_ptr() + 4 * a, param_cumprod.data_ptr() + 4 * a,
                                             grad_out.data_ptr() + 4 * a, inv.data_ptr() + 4 * a, inv_len.data_ptr(),
                                             grad_in.data_ptr() + 4 * a, b - a, inv_len.numel(), ws.data_ptr(),
                                             ws.numel(), stream), "gcp_cumprod_bwd_f32")


def validate_segments(inv: torch.Tensor, inv_len: torch.Tensor) -> int:
    """Number of violations of the (inv, inv_len) layout contract; 0 means consistent.  Synchronises."""
    import ctypes

    _check(inv, "inv", torch.int32)
    _check(inv_len, "inv_len", torch.int32, None, inv.device)
    with torch.cuda.device(inv.device):
        ws, stream = _workspace(inv.device, 0)
        out = ctypes.c_int64(0)
        _lib.check(_lib.lib().gcp_validate_segments(inv.data_ptr(), inv_len.data_ptr(), inv.numel(),
                                                    inv_len.numel(), ws.data_ptr(), ws.numel(), stream,
                                                    ctypes.byref(out)), "gcp_validate_segments")
    return int(out.value)


def workspace_status(device=None) -> int:
    """Synchronise the current stream and return the sticky watchdog status (0 = OK)."""
    import ctypes

    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(device):
        ws, stream = _workspace(device, 0)
        st = ctypes.c_int(0)
        _lib.check(_lib.lib().gcp_workspace_status(ws.data_ptr(), stream, ctypes.byref(st)), "gcp_workspace_status")
    return int(st.value)


def set_variant(op: str, variant: int) -> None:
    """Test hook: op in {'fwd','bwd'}; 0 = plain-load kernel pair, 1 = persistent blocked kernel, -1 = default."""
    _lib.check(_lib.lib().gcp_set_variant(0 if op == "fwd" else 1, int(variant)), "gcp_set_variant")


def set_option(option: int, value: int) -> None:
    """Tuning hook: option 0 = halo resolution of tile carries (1 on / 0 always look back); option 1 = chained
    tile ranges in the blocked backward (0 tickets / 1 always, the default / 2 from the workspace hint)."""
    _lib.check(_lib.lib().gcp_set_option(int(option), int(value)), "gcp_set_option")


def _options_from_env() -> None:
    """Experiments only: GCP_OPT_HALO / GCP_OPT_CHAIN in the environment override the defaults at first use."""
    for name, idx in (("GCP_OPT_HALO", 0), ("GCP_OPT_CHAIN", 1), ("GCP_OPT_CHAIN_FWD", 2)):
        if os.environ.get(name, "") != "":
            set_option(idx, int(os.environ[name]))


def variants(op: str) -> list:
    L = _lib.lib()
    o = 0 if op == "fwd" else 1
    return [L.gcp_variant_name(o, i).decode() for i in range(L.gcp_num_variants(o))]


def last_launch_count() -> int:
    return int(_lib.lib().gcp_last_launch_count())
