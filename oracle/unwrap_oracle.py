"""TEST INFRASTRUCTURE (never imported by the product path): CPU restatement of 2-D phase unwrapping by reliability
sorting (Herraez, Burton, Lalor, Gdeisat, Appl. Opt. 41, 7437 (2002)) -- the algorithm behind
``skimage.restoration.unwrap_phase`` for 2-D input, which ``utils/functions.py:44-59`` of the reference calls per image.

PARITY UNPINNED against scikit-image itself: the package is not installed in this image, not vendored by the reference
(requirements.txt has no pin for it) and the reference ships no unwrapped golden vectors.  This file restates the published
steps; `tests/test_oracle.py` pins it through properties every correct unwrapper has (exact recovery of residue-free surfaces
up to one global multiple of 2 pi; output - input is an integer multiple of 2 pi), and `tests/test_gpu_unwrap.py` compares the
CUDA kernel with it.

Steps (numbered as in csrc/unwrap.cuh): reliability from wrapped second differences (borders: a large constant), edges with
the summed reliability and the jump count of their two pixels, stable ascending sort, sequential group merging (single pixel
joins its neighbour's group, else the smaller group joins the larger, ties: the first pixel's group joins).
"""
from __future__ import annotations

import numpy as np

PI = np.float32(np.pi)
TWO_PI = np.float32(2 * np.pi)


def _wrap(x):
    x = x.astype(np.float32)
    return np.where(x > PI, x - TWO_PI, np.where(x < -PI, x + TWO_PI, x)).astype(np.float32)


def reliability(ph: np.ndarray) -> np.ndarray:
    ph = ph.astype(np.float32)
    h, w = ph.shape
    rel = np.full((h, w), 9999999.0, dtype=np.float32)
    c = ph[1:-1, 1:-1]
    H = _wrap(ph[1:-1, :-2] - c) - _wrap(c - ph[1:-1, 2:])
    V = _wrap(ph[:-2, 1:-1] - c) - _wrap(c - ph[2:, 1:-1])
    D1 = _wrap(ph[:-2, :-2] - c) - _wrap(c - ph[2:, 2:])
    D2 = _wrap(ph[:-2, 2:] - c) - _wrap(c - ph[2:, :-2])
    rel[1:-1, 1:-1] = (H * H + V * V) + (D1 * D1 + D2 * D2)
    return rel


def unwrap2d(ph: np.ndarray) -> np.ndarray:
    """One image [H, W] -> unwrapped float32 [H, W]."""
    ph = np.ascontiguousarray(ph, dtype=np.float32)
    h, w = ph.shape
    rel = reliability(ph).reshape(-1)
    flat = ph.reshape(-1)
    idx = np.arange(h * w).reshape(h, w)
    p1 = np.concatenate([idx[:, :-1].reshape(-1), idx[:-1, :].reshape(-1)])
    p2 = np.concatenate([idx[:, 1:].reshape(-1), idx[1:, :].reshape(-1)])
    key = (rel[p1] + rel[p2]).astype(np.float32)
    order = np.argsort(key, kind="stable")
    parent = np.arange(h * w)
    off = np.zeros(h * w, dtype=np.int64)
    size = np.ones(h * w, dtype=np.int64)
    base = np.zeros(h * w, dtype=np.int64)

    def find(x):
        acc = 0
        while parent[x] != x:
            px = parent[x]
            if parent[px] != px:
                off[x] += off[px]
                parent[x] = parent[px]
            acc += off[x]
            x = parent[x]
        return x, acc

    for k in order:
        a, b = int(p1[k]), int(p2[k])
        r1, o1 = find(a)
        r2, o2 = find(b)
        if r1 == r2:
            continue
        d = np.float32(flat[a] - flat[b])
        e = -1 if d > PI else (1 if d < -PI else 0)
        inc1, inc2 = base[r1] + o1, base[r2] + o2
        group2_joins = True if size[r2] == 1 else (False if size[r1] == 1 else bool(size[r1] > size[r2]))
        if group2_joins:
            off[r2] = base[r2] + (inc1 - e - inc2) - base[r1]
            parent[r2] = r1
            size[r1] += size[r2]
        else:
            off[r1] = base[r1] + (inc2 + e - inc1) - base[r2]
            parent[r1] = r2
            size[r2] += size[r1]
    inc = np.zeros(h * w, dtype=np.int64)
    for x in range(h * w):
        r, o = find(x)
        inc[x] = base[r] + o
    return (flat + TWO_PI * inc.astype(np.float32)).reshape(h, w).astype(np.float32)


def unwrap(x: np.ndarray) -> np.ndarray:
    """[B, 1, H, W] or [B, H, W] -> [B, 1, H, W] (the shape utils/functions.py:58 returns)."""
    x = np.asarray(x)
    imgs = x.reshape((-1,) + x.shape[-2:])
    return np.stack([unwrap2d(im) for im in imgs])[:, None]
