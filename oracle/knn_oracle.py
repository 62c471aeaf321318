"""numpy restatement of pc_processor/postproc/knn.py:55-143 (KNN.forward).  TEST INFRASTRUCTURE.

Pinned against the reference's own torch implementation by tests/test_oracle_pinning.py (live, when
/root/reference is present) and tests/golden/knn_*.npz (everywhere).

Tie rule.  The reference selects neighbours with ``topk(k, largest=False, sorted=False)`` whose tie order is
unspecified; the restatement (and the CUDA kernel) break distance ties by the LOWER window index
(row-major over the S x S window).  On inputs with no exact distance ties at the k-th place the result is
the reference's result; ``tie_free_mask`` reports which points are in that situation so tests can compare
bit-exactly there.
"""
import math

import numpy as np


def gaussian_kernel(kernel_size, sigma):
    """knn.py:12-34 — fp32, same operation order as the torch code."""
    import torch  # fp32 exp must match torch's; used only to build this S*S table

    x_coord = torch.arange(kernel_size)
    x_grid = x_coord.repeat(kernel_size).view(kernel_size, kernel_size)
    y_grid = x_grid.t()
    xy_grid = torch.stack([x_grid, y_grid], dim=-1).float()
    mean = (kernel_size - 1) / 2.0
    variance = sigma ** 2.0
    g = (1.0 / (2.0 * math.pi * variance)) * torch.exp(-torch.sum((xy_grid - mean) ** 2.0, dim=-1) / (2 * variance))
    g = g / torch.sum(g)
    return g.view(kernel_size, kernel_size).numpy()


def inv_gauss_table(search, sigma):
    """(1 - gaussian) flattened row-major, fp32 (knn.py:103-105)."""
    return (1 - gaussian_kernel(search, sigma)).astype(np.float32).reshape(-1)


def knn_vote(proj_range, unproj_range, proj_argmax, px, py, knn, search, sigma, cutoff, nclasses,
             return_aux=False):
    """proj_range (H,W) f32 with <0 at empty pixels; unproj_range (P,) f32; proj_argmax (H,W) int;
    px (P,) column index, py (P,) row index.  Returns int64 (P,) labels in 1..nclasses-1."""
    if search % 2 == 0:
        raise ValueError("Nearest neighbor kernel must be odd number")  # knn.py:73-74
    H, W = proj_range.shape
    P = unproj_range.shape[0]
    pad = (search - 1) // 2
    S2 = search * search
    center = (S2 - 1) // 2
    rng = np.zeros((H + 2 * pad, W + 2 * pad), np.float32)  # F.unfold zero-pads (knn.py:80-82)
    rng[pad:pad + H, pad:pad + W] = proj_range
    lab = np.zeros((H + 2 * pad, W + 2 * pad), np.int64)  # and the label image too (:115-117)
    lab[pad:pad + H, pad:pad + W] = proj_argmax
    px = np.asarray(px, np.int64)
    py = np.asarray(py, np.int64)
    r = np.asarray(unproj_range, np.float32)
    nr = np.empty((S2, P), np.float32)
    nl = np.empty((S2, P), np.int64)
    for i in range(search):
        for j in range(search):
            nr[i * search + j] = rng[py + i, px + j]
            nl[i * search + j] = lab[py + i, px + j]
    nr[nr < 0] = np.inf  # :91
    nr[center] = r  # :94-95
    with np.errstate(invalid="ignore"):
        d = np.abs(nr - r[None, :]) * inv_gauss_table(search, sigma)[:, None]  # :98-108 (fp32)
    d = d.astype(np.float32)
    order = np.argsort(d, axis=0, kind="stable")[:knn]  # k smallest, ties -> lower window index
    sel_d = np.take_along_axis(d, order, axis=0)
    sel_l = np.take_along_axis(nl, order, axis=0)
    if cutoff > 0:
        sel_l = np.where(sel_d > np.float32(cutoff), nclasses, sel_l)  # :125-128
    votes = np.zeros((nclasses + 1, P), np.int32)
    for k in range(knn):
        np.add.at(votes, (sel_l[k], np.arange(P)), 1)  # :132-135
    out = votes[1:-1].argmax(axis=0).astype(np.int64) + 1  # :138 (first max wins)
    if return_aux:
        # A tie at the k-th place can only change the vote if the tied candidates carry different
        # (post-cutoff) labels; everything else is uniquely determined whatever order topk returns.
        ds = np.sort(d, axis=0, kind="stable")
        if knn < S2:
            eff = np.where(d > np.float32(cutoff), nclasses, nl) if cutoff > 0 else nl
            tied = d == ds[knn - 1][None, :]
            lo = np.where(tied, eff, np.iinfo(np.int64).max).min(axis=0)
            hi = np.where(tied, eff, -1).max(axis=0)
            tie_free = ~((ds[knn - 1] == ds[knn]) & (lo != hi))
        else:
            tie_free = np.ones(P, bool)
        return out, tie_free
    return out
