"""Batch construction of an HNSW-shaped index on the GPU ("next" row).

The reference builds its graph by serial insertion under a write lock (idx/trees/hnsw/mod.rs:230-394); that
is not a data-parallel algorithm, so instead of porting it this builder uses the brute-force KNN engine:
every layer l holds the elements whose level (floor(-ln U * ml), ml = 1/ln m -- the reference's level law,
hnsw/mod.rs:263-266) is >= l, and each element's neighbours in layer l are its exact m (m0 on layer 0)
nearest elements of that layer, found with the tcgen05 screen + exact re-rank.  The result is a valid input
for the layer-walk kernel (same CSR format the reference's Hn records decode to); it is NOT the graph the
reference would have built, so parity claims apply to the walk on a given graph, and quality is measured as
recall against exact brute force.
"""
import math

import numpy as np

from .engine import VectorColumn


def assign_levels(n, m, seed):
    rng = np.random.default_rng(seed)
    ml = 1.0 / math.log(m)
    u = rng.random(n)
    u[u == 0.0] = 0.5
    return np.floor(-np.log(u) * ml).astype(np.int64)


def build_layers(ctx, vectors_dev, n, dim, metric="EUCLIDEAN", m=16, m0=32, seed=1, batch=4096, progress=None):
    """vectors_dev: torch CUDA float32 tensor (n, dim).  -> (layers, entry_point, levels) with
    layers = [(row_ptr u64[n+1], col_idx u32[e]), ...] (layer 0 first), element id = row index."""
    import torch
    levels = assign_levels(n, m, seed)
    top = int(levels.max()) if n else 0
    layers = []
    for l in range(top + 1):
        members = np.nonzero(levels >= l)[0].astype(np.int64)
        k_nb = m0 if l == 0 else m
        row_ptr = np.zeros(n + 1, np.uint64)
        if members.size <= 1:
            layers.append((row_ptr, np.zeros(0, np.uint32)))
            continue
        midx = torch.from_numpy(members).to(vectors_dev.device)
        sub = vectors_dev if members.size == n else vectors_dev.index_select(0, midx).contiguous()
        col = VectorColumn(ctx, dim, metric, "F32", capacity=members.size)
        col.append_device(sub.data_ptr(), members.size)
        col.finalize()
        k = min(k_nb + 1, members.size)  # +1: the element itself comes back at distance 0
        nbrs = np.zeros((members.size, k_nb), np.int64)
        counts = np.zeros(members.size, np.int64)
        o_r = torch.zeros((batch, k), dtype=torch.int64, device=vectors_dev.device)
        o_d = torch.zeros((batch, k), dtype=torch.float64, device=vectors_dev.device)
        o_c = torch.zeros((batch,), dtype=torch.int32, device=vectors_dev.device)
        for b0 in range(0, members.size, batch):
            b1 = min(members.size, b0 + batch)
            q = sub[b0:b1].to(torch.float64).contiguous()
            col.knn_device(q.data_ptr(), b1 - b0, k, 0, o_r.data_ptr(), o_d.data_ptr(), o_c.data_ptr())
            r = o_r[: b1 - b0].cpu().numpy()
            cnt = o_c[: b1 - b0].cpu().numpy().astype(np.int64)
            valid = np.arange(k)[None, :] < cnt[:, None]
            keep = valid & (r != (b0 + np.arange(b1 - b0))[:, None])  # drop self, keep nearest-first order
            order = np.argsort(~keep, axis=1, kind="stable")[:, :k_nb]
            nbrs[b0:b1, : order.shape[1]] = np.take_along_axis(r, order, axis=1)
            counts[b0:b1] = np.minimum(keep.sum(1), k_nb)
            if progress:
                progress(l, b1, members.size)
        col.close()
        deg = np.zeros(n, np.int64)
        deg[members] = counts
        row_ptr[1:] = np.cumsum(deg)
        col_idx = np.zeros(int(row_ptr[-1]), np.uint32)
        flat = members[nbrs]  # local -> global element ids
        mask = np.arange(k_nb)[None, :] < counts[:, None]
        col_idx[:] = flat[mask].astype(np.uint32)
        layers.append((row_ptr, col_idx))
    entry = int(np.argmax(levels)) if n else -1
    return layers, entry, levels
