"""TEST INFRASTRUCTURE — numpy restatement of the integer side of the fused compositor route (csrc/gcp_tile.cu).

The reference sorts the Gaussian x pixel elements by pixel key with a stable sort (gs_model.py:538-548), so inside a
pixel the Gaussians stay in depth (= index) order.  The tile route never builds that list; it bins (tile, Gaussian)
PAIRS instead.  These functions restate the binning so that tests can check, without a GPU, that walking the pairs
of a tile in order visits every pixel's Gaussians in exactly the order of the reference's sorted element list, and,
on the GPU, that the native pair list / tile offsets / piece plan are bit-identical to this restatement.
Only tests/ may import this module."""
from __future__ import annotations

import numpy as np

TW, TH = 8, 4   # tile = 8 x 4 pixels = one warp


def num_tiles(W: int, H: int, tw: int = TW, th: int = TH):
    """Tiles over the inclusive pixel grid [0, W] x [0, H] (gs_model.py:505: the image is (H+1, W+1, 3))."""
    ntx = (W + tw) // tw
    nty = (H + th) // th
    return ntx, nty


def tile_pairs(sp, ep, W: int, H: int, tw: int = TW, th: int = TH):
    """Gaussian-major (tile, Gaussian) pairs: one per tile the box (clipped to the image) touches, row-major over
    the tiles of the box.  Returns tiles i64[P], gids i32[P], toff i64[n+1] (exclusive pair offsets)."""
    ntx, _ = num_tiles(W, H, tw, th)
    tiles, gids, counts = [], [], []
    for g in range(len(sp)):
        sx, sy = max(int(sp[g][0]), 0), max(int(sp[g][1]), 0)
        ex, ey = min(int(ep[g][0]), W), min(int(ep[g][1]), H)
        c = 0
        if ex >= sx and ey >= sy:
            for ty in range(sy // th, ey // th + 1):
                for tx in range(sx // tw, ex // tw + 1):
                    tiles.append(ty * ntx + tx)
                    gids.append(g)
                    c += 1
        counts.append(c)
    toff = np.concatenate([[0], np.cumsum(np.asarray(counts, np.int64))]).astype(np.int64)
    return np.asarray(tiles, np.int64), np.asarray(gids, np.int32), toff


def tile_pairs_np(sp, ep, W: int, H: int, tw: int = TW, th: int = TH):
    """tile_pairs without the Python loops (views of 10^5..10^6 Gaussians); same three arrays, bit for bit."""
    sp, ep = np.asarray(sp, np.int64), np.asarray(ep, np.int64)
    ntx, _ = num_tiles(W, H, tw, th)
    sx, sy = np.maximum(sp[:, 0], 0), np.maximum(sp[:, 1], 0)
    ex, ey = np.minimum(ep[:, 0], W), np.minimum(ep[:, 1], H)
    tx0, ty0 = sx // tw, sy // th
    nx, ny = ex // tw - tx0 + 1, ey // th - ty0 + 1
    cnt = np.where((ex >= sx) & (ey >= sy), nx * ny, 0)
    toff = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    gids = np.repeat(np.arange(len(sp), dtype=np.int64), cnt)
    local = np.arange(int(toff[-1]), dtype=np.int64) - toff[:-1][gids]
    row = local // np.maximum(nx[gids], 1)
    tiles = (ty0[gids] + row) * ntx + tx0[gids] + (local - row * nx[gids])
    return tiles, gids.astype(np.int32), toff


def sort_by_tile(tiles, gids, ntiles: int):
    """Stable sort of the pairs by tile (Gaussian order kept inside a tile); tile_start i32[ntiles+1]."""
    order = np.argsort(tiles, kind="stable")
    start = np.searchsorted(tiles[order], np.arange(ntiles + 1), side="left").astype(np.int32)
    return gids[order], start, order


def piece_plan(tile_start, piece: int):
    """Work units of the walk kernels: a tile's list cut into pieces of at most `piece` pairs (an empty tile is one
    piece).  Returns piece_start i32[ntiles+1] (exclusive piece offsets per tile) and piece_tile i32[pieces]."""
    length = np.diff(tile_start.astype(np.int64))
    cnt = np.where(length <= piece, 1, -(-length // piece))
    pstart = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    ptile = np.repeat(np.arange(len(length), dtype=np.int32), cnt)
    return pstart, ptile


def pixel_lists_from_pairs(sp, ep, W: int, H: int, tw: int = TW, th: int = TH):
    """For every pixel key y*10000+x the Gaussians covering it, in the order a warp meets them when it walks the
    pixel's tile list: {key: [gid, ...]}."""
    ntx, nty = num_tiles(W, H, tw, th)
    tiles, gids, _ = tile_pairs(sp, ep, W, H, tw, th)
    gid_s, start, _ = sort_by_tile(tiles, gids, ntx * nty)
    out = {}
    for t in range(ntx * nty):
        ty, tx = divmod(t, ntx)
        for g in gid_s[start[t]:start[t + 1]]:
            sx, sy = max(int(sp[g][0]), 0), max(int(sp[g][1]), 0)
            ex, ey = min(int(ep[g][0]), W), min(int(ep[g][1]), H)
            for y in range(max(sy, ty * th), min(ey, ty * th + th - 1) + 1):
                for x in range(max(sx, tx * tw), min(ex, tx * tw + tw - 1) + 1):
                    out.setdefault(y * 10000 + x, []).append(int(g))
    return out
