"""make_batches (reference src/ptyrad/reconstruction.py:479-587): 'random' on the CPU, 'compact' / 'sparse' grouping on the GPU."""
import numpy as np
import pytest


def _scan(ny, nx, seed=0):
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    return np.stack([yy.ravel(), xx.ravel()], 1) * 7.3 + rng.normal(0, 0.4, (ny * nx, 2))


def reference_sparse_greedy(indices, pos, compact_batches):
    """NumPy restatement of the reference's 'sparse' branch given the compact groups (reconstruction.py:546-586)."""
    from scipy.spatial.distance import cdist
    pos_s = pos[indices]
    sparse_indices = indices.copy()
    centroids = np.array([np.mean(pos[cb], axis=0) for cb in compact_batches])
    pairwise = cdist(pos, pos)
    sparse_batches, used = [], []
    for g in range(len(compact_batches)):
        closest_s = np.argmin(np.linalg.norm(pos_s - centroids[g], axis=1))
        sparse_batches.append([indices[closest_s]])
        used.append(closest_s)
    sparse_indices = np.delete(sparse_indices, used)
    for idx in sparse_indices:
        mins = [np.min(pairwise[sparse_batches[g], idx]) for g in range(len(compact_batches))]
        sparse_batches[int(np.argmax(mins))].append(idx)
    return [np.array(b) for b in sparse_batches]


def test_make_batches_random_is_a_partition():
    from ptyrad_b200.batching import make_batches
    idx = np.arange(5, 405)
    pos = _scan(21, 21)
    b = make_batches(idx, pos, 32, mode="random", verbose=False, seed=3)
    assert len(b) == 400 // 32
    assert np.array_equal(np.sort(np.concatenate(b)), idx)
    with pytest.raises(ValueError):
        make_batches(np.arange(500), pos, 32, verbose=False)
    with pytest.raises(ValueError):
        make_batches(idx, pos, 32, mode="spiral", verbose=False)


@pytest.mark.gpu
@pytest.mark.parametrize("subset", [False, True])
def test_sparse_groups_equal_the_reference_greedy_loop(subset):
    """Same compact groups in, same sparse groups out (members AND order of joining), for the full scan and for a sub-selection."""
    import torch
    from ptyrad_b200.batching import kmeans_labels, sparse_groups_from_compact
    pos = _scan(24, 20, seed=1)
    indices = np.arange(len(pos))
    if subset:
        indices = np.sort(np.random.default_rng(2).permutation(len(pos))[:333])
    G = len(indices) // 40
    labels = kmeans_labels(torch.as_tensor(pos[indices], device="cuda"), G, seed=5).cpu().numpy()
    compact = [indices[np.where(labels == g)[0]] for g in range(G)]
    assert all(len(c) for c in compact)
    ref = reference_sparse_greedy(indices, pos, compact)
    ours = sparse_groups_from_compact(indices, pos, compact)
    assert len(ours) == len(ref)
    for a, b in zip(ours, ref):
        assert np.array_equal(a, b)


@pytest.mark.gpu
def test_make_batches_compact_and_sparse_properties():
    """compact groups are spatially tight, sparse groups spread over the field of view (what the two modes are for); both are
    partitions of the indices; a 128x128 scan is grouped in seconds."""
    import time
    from ptyrad_b200.batching import make_batches
    pos = _scan(32, 32, seed=4)
    idx = np.arange(len(pos))
    spread = lambda batches: float(np.mean([pos[b].std(0).mean() for b in batches]))
    out = {}
    for mode in ("random", "compact", "sparse"):
        b = make_batches(idx, pos, 64, mode=mode, verbose=False, seed=7)
        assert np.array_equal(np.sort(np.concatenate(b)), idx)
        assert len(b) == 16
        out[mode] = spread(b)
    assert out["compact"] < 0.5 * out["random"] and out["sparse"] > 0.9 * out["random"]
    # nearest-neighbour distance inside a sparse group is far larger than inside a random one
    def nn(b):
        from scipy.spatial.distance import cdist
        d = cdist(pos[b], pos[b]) + np.eye(len(b)) * 1e9
        return d.min(1).mean()
    bs = make_batches(idx, pos, 64, mode="sparse", verbose=False, seed=7)
    br = make_batches(idx, pos, 64, mode="random", verbose=False, seed=7)
    assert np.mean([nn(b) for b in bs]) > 1.5 * np.mean([nn(b) for b in br])
    big = _scan(128, 128, seed=5)
    t0 = time.time()
    bb = make_batches(np.arange(len(big)), big, 256, mode="sparse", verbose=False, seed=1)
    assert time.time() - t0 < 60 and np.array_equal(np.sort(np.concatenate(bb)), np.arange(len(big)))
