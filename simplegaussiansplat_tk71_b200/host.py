"""Host-buffer entry point: segmented cumprod forward + backward for arrays that live in HOST memory.

The reference never moves the element arrays over PCIe (they are produced on the device, gs_model.py:598-605),
but a caller that does hold them on the host gets the best the link allows from this streamer:

  * only x, grad_out and ONE BIT per element of `key` go up (8.125 B/element): the ops need to know where the
    runs of equal adjacent keys start and nothing else, so the host packs the run starts into bits
    (`gcp_host_boundary_bits`: OpenMP + AVX2, memory-bound, ~0.5 ms per 8 Mi-element chunk) and the device rebuilds
    dense segment ids with one scan (`gcp_ids_from_bits`); those ids serve as `key` of the forward and as `inv` of
    the backward op, so `inv` / `inv_len` never cross the bus either;
  * the list is cut into chunks at segment boundaries (pixel lists are independent) and the chunks flow through
    three CUDA streams — H2D copies, the two scan launches, D2H copies — with `depth` buffer sets, so the
    uploads, the kernels and the downloads of neighbouring chunks overlap and PCIe runs full duplex.

All host tensors must be pinned for the copies to be asynchronous.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _lib, ops


def _chunk_sizes(n: int, chunk: int) -> list:
    """Target chunk sizes: small first and last chunks (chunk/8, /4, /2, ... ramping up to `chunk` and down again).
    While the first chunk goes up nothing comes down, and while the last one comes down nothing goes up, so short
    end chunks cut the fill / drain time of the full-duplex pipeline (1.5 ms of a 14 ms step with equal chunks)."""
    ramp = [max(1, chunk >> s) for s in (3, 2, 1)]
    if n <= 2 * sum(ramp) + chunk:
        return [chunk] * max(1, -(-n // chunk))
    mid = n - 2 * sum(ramp)
    k = -(-mid // chunk)
    return ramp + [-(-mid // k)] * k + ramp[::-1]


def _cut_points(key: torch.Tensor, chunk: int, ramp: bool = False) -> list:
    """Chunk boundaries near the targets (multiples of `chunk`, or the ramped sizes of _chunk_sizes), each moved
    forward to the next segment boundary."""
    n = key.numel()
    cuts = [0]
    k = key.numpy()
    sizes = _chunk_sizes(n, chunk) if ramp else None
    pos = sizes[0] if ramp else chunk
    while pos < n:
        # first index >= pos where a new segment starts
        window = 1 << 16
        j = pos
        while j < n:
            w = k[j - 1:min(n, j + window)]
            d = np.flatnonzero(w[1:] != w[:-1])
            if d.size:
                j = j + int(d[0])
                break
            j += window
        if j >= n:
            break
        cuts.append(j)
        pos = j + (sizes[min(len(cuts) - 1, len(sizes) - 1)] if ramp else chunk)
    cuts.append(n)
    return cuts


class HostStreamer:
    """Reusable pipeline for one device; buffers are allocated once for `max_chunk` elements."""

    def __init__(self, device="cuda", chunk_elems: int = 8 << 20, depth: int = 3):
        self.device = torch.device(device)
        self.chunk = int(chunk_elems)
        self.depth = depth
        cap = self.chunk + (1 << 20)  # a chunk ends at the first segment boundary after `chunk`
        self.cap = cap
        mk = lambda dt: [torch.empty(cap, dtype=dt, device=self.device) for _ in range(depth)]  # noqa: E731
        self.dx, self.dg, self.dy, self.dgin = mk(torch.float32), mk(torch.float32), mk(torch.float32), mk(torch.float32)
        self.dk = mk(torch.int32)
        words = (cap + 31) // 32 + 1
        self.dbits = [torch.empty(words, dtype=torch.int32, device=self.device) for _ in range(depth)]
        self.hbits = [torch.empty(words, dtype=torch.int32).pin_memory() for _ in range(depth)]
        self.lib = _lib.lib()
        self.scan_tmp = torch.empty(int(self.lib.gcp_ids_from_bits_bytes(cap)), dtype=torch.uint8, device=self.device)
        # host threads of the key packing: the cores divided among the ranks of this node (torchrun sets
        # LOCAL_WORLD_SIZE) — with every rank taking all of them the packing threads of the ranks fight each other
        ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1))
        self.threads = max(1, min(16, (os.cpu_count() or 1) // ranks))
        self.s_up = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_down = torch.cuda.Stream(self.device)
        self.empty_len = torch.empty(0, dtype=torch.int32, device=self.device)

    def fwd_bwd(self, x, key, grad_out, y_out, grad_in_out, cuts=None, sync: bool = True):
        """y_out = segmented inclusive cumprod(x by key); grad_in_out = its gradient for upstream grad_out.
        All five are 1-D pinned host tensors (f32, i32, f32, f32, f32).  Returns (h2d_bytes, d2h_bytes).

        sync=True (default): returns when y_out and grad_in_out hold the results (the download stream is
        synchronised).  sync=False: returns as soon as everything is queued — the host buffers are then valid only
        after the caller synchronises the current stream (which has been made to wait for the downloads)."""
        n = x.numel()
        if cuts is None:
            cuts = _cut_points(key, self.chunk, ramp=True)
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_up, self.s_run, self.s_down):
            s.wait_stream(cur)
        up_done = [None] * self.depth
        run_done = [None] * self.depth
        down_done = [None] * self.depth
        for c in range(len(cuts) - 1):
            a, b = cuts[c], cuts[c + 1]
            m = b - a
            if m > self.cap:
                raise RuntimeError(f"segment longer than the streaming buffers ({m} > {self.cap} elements)")
            i = c % self.depth
            words = (m + 31) // 32
            if up_done[i] is not None:
                up_done[i].synchronize()                    # the pinned bit buffer i has been read by its upload
            # where the runs of equal keys start, one bit per element (a chunk begins at a segment boundary)
            _lib.check(self.lib.gcp_host_boundary_bits(key.data_ptr() + 4 * a, m, self.hbits[i].data_ptr(),
                                                       self.threads), "gcp_host_boundary_bits")
            with torch.cuda.stream(self.s_up):
                if down_done[i] is not None:
                    self.s_up.wait_event(down_done[i])      # buffer set i is free again
                self.dx[i][:m].copy_(x[a:b], non_blocking=True)
                self.dbits[i][:words].copy_(self.hbits[i][:words], non_blocking=True)
                self.dg[i][:m].copy_(grad_out[a:b], non_blocking=True)
                up_done[i] = self.s_up.record_event()
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(up_done[i])
                xs, ks, gs, ys, gi = self.dx[i][:m], self.dk[i][:m], self.dg[i][:m], self.dy[i][:m], self.dgin[i][:m]
                _lib.check(self.lib.gcp_ids_from_bits(self.dbits[i].data_ptr(), m, ks.data_ptr(),
                                                      self.scan_tmp.data_ptr(), self.scan_tmp.numel(),
                                                      self.s_run.cuda_stream), "gcp_ids_from_bits")
                ops.grouped_cumprod_forward(xs, ks, ys)
                ops.grouped_cumprod_backward(xs, ys, gs, ks, gi, self.empty_len)
                run_done[i] = self.s_run.record_event()
            with torch.cuda.stream(self.s_down):
                self.s_down.wait_event(run_done[i])
                y_out[a:b].copy_(self.dy[i][:m], non_blocking=True)
                grad_in_out[a:b].copy_(self.dgin[i][:m], non_blocking=True)
                down_done[i] = self.s_down.record_event()
        cur.wait_stream(self.s_down)
        if sync:
            self.s_down.synchronize()
        return 8 * n + 4 * sum((cuts[c + 1] - cuts[c] + 31) // 32 for c in range(len(cuts) - 1)), 8 * n
