"""``make_batches`` (reference ``src/ptyrad/reconstruction.py:479-587``) with the position grouping on the GPU.

Same signature and return value (a list of index arrays).  'random' is the reference's NumPy permutation.  'compact' clusters the
selected positions with k-means (k-means++ seeding + Lloyd iterations as tensor ops on the device; the reference calls scikit-learn's
MiniBatchKMeans, which is randomised, so the groups agree in kind, not index by index).  'sparse' runs the reference's greedy loop --
every remaining point joins the group whose nearest member is farthest from it -- in ONE kernel (``ptyb200_sparse_groups``,
csrc/grouping.cuh): given the same compact groups it returns exactly the reference's groups, in well under a second where the
reference documents "more than 10 min on a CPU" for a 256x256 scan (reconstruction.py:491).
"""
from __future__ import annotations

from time import time

import numpy as np
import torch

from . import _lib, engine
from ._lib import ptr


def kmeans_labels(pos_s: torch.Tensor, k: int, iters: int = 20, seed: int | None = None) -> torch.Tensor:
    """k-means++ seeding and Lloyd iterations on (n,2) device positions; returns (n,) int64 labels.  Empty clusters are re-seeded
    with the point farthest from its centre."""
    g = torch.Generator(device=pos_s.device)
    if seed is not None:
        g.manual_seed(int(seed))
    n = pos_s.shape[0]
    x = pos_s.to(torch.float64)
    centres = torch.empty((k, 2), dtype=torch.float64, device=x.device)
    centres[0] = x[torch.randint(n, (1,), generator=g, device=x.device)]
    d2 = ((x - centres[0]) ** 2).sum(1)
    for c in range(1, k):
        pick = torch.multinomial(d2 / d2.sum(), 1, generator=g)
        centres[c] = x[pick]
        d2 = torch.minimum(d2, ((x - centres[c]) ** 2).sum(1))
    labels = None
    for _ in range(iters):
        dist = torch.cdist(x, centres)
        new = dist.argmin(1)
        if labels is not None and torch.equal(new, labels):
            break
        labels = new
        cnt = torch.bincount(labels, minlength=k)
        sums = torch.zeros_like(centres).index_add_(0, labels, x)
        centres = torch.where(cnt[:, None] > 0, sums / cnt.clamp(min=1)[:, None], centres)
        empty = (cnt == 0).nonzero().flatten()
        if empty.numel():
            far = dist.gather(1, labels[:, None]).flatten().topk(empty.numel()).indices
            centres[empty] = x[far]
    return labels


def sparse_groups_from_compact(indices, pos, compact_batches, device="cuda"):
    """The 'sparse' branch of the reference given its compact groups (reconstruction.py:540-587): seeds = the selected point closest
    to each compact centroid, then the greedy assignment on the device.  Returns a list of index arrays (seed first, then the points
    in the order they joined, like the reference's lists)."""
    indices = np.asarray(indices)
    pos = np.asarray(pos, dtype=np.float64)
    pos_s = pos[indices]
    G = len(compact_batches)
    centroids = np.array([np.mean(pos[np.asarray(cb)], axis=0) for cb in compact_batches])
    used = [int(np.argmin(np.linalg.norm(pos_s - centroids[g], axis=1))) for g in range(G)]
    if len(set(used)) != G:
        raise ValueError("two compact groups share their closest point: cannot seed the sparse groups")
    rest = np.delete(np.arange(len(indices)), used)            # positions inside `indices`, reference order
    order = np.concatenate([np.asarray(used, dtype=np.int64), rest])
    ordered = torch.as_tensor(np.ascontiguousarray(pos_s[order]), dtype=torch.float64, device=device)
    labels = torch.empty(len(order), dtype=torch.int32, device=device)
    _lib.check(_lib.lib().ptyb200_sparse_groups(ptr(ordered), len(order), G, ptr(labels), engine._stream()))
    labels = labels.cpu().numpy()
    members = indices[order]
    by_group = np.argsort(labels, kind="stable")               # stable: the order of joining inside every group is kept
    bounds = np.searchsorted(labels[by_group], np.arange(G + 1))
    return [members[by_group[bounds[g]:bounds[g + 1]]] for g in range(G)]


def make_batches(indices, pos, batch_size, mode="random", verbose=True, device="cuda", seed=None):
    """Drop-in for the reference's make_batches (reconstruction.py:479-587); `device` / `seed` are additions."""
    indices = np.asarray(indices)
    if len(indices) > len(pos):
        raise ValueError(f"len(indices) = '{len(indices)}' is larger than total number of probe positions ({len(pos)}), check your indices generation params")
    if indices.max() > len(pos):
        raise ValueError(f"Maximum index '{indices.max()}' is larger than total number of probe positions ({len(pos)}), check your indices generation params")
    num_batch = len(indices) // batch_size
    t0 = time()
    if mode == "random":
        rng = np.random.default_rng(seed)
        batches = np.array_split(rng.permutation(indices), num_batch)
    elif mode in ("compact", "sparse"):
        pos_d = torch.as_tensor(np.asarray(pos, dtype=np.float64)[indices], device=device)
        labels = kmeans_labels(pos_d, num_batch, seed=seed).cpu().numpy()
        batches = [indices[np.where(labels == b)[0]] for b in range(num_batch)]
        if mode == "sparse":
            batches = sparse_groups_from_compact(indices, pos, batches, device=device)
            flat = np.sort(np.concatenate(batches))
            assert np.array_equal(flat, np.sort(indices)), "sparse grouping lost or duplicated an index"
    else:
        raise ValueError(f"Batch grouping mode '{mode}' not implemented, please use 'random', 'compact' or 'sparse'")
    if verbose:
        print(f"Generated {num_batch} '{mode}' groups of ~{batch_size} scan positions in {time() - t0:.3f} sec")
    return batches
