"""Batch construction of an HNSW-shaped index on the GPU ("next" row).

The reference builds its graph by serial insertion under a write lock (idx/trees/hnsw/mod.rs:230-394); that is not a
data-parallel algorithm, so instead of porting it this builder works layer by layer on whole batches:

  * levels follow the reference's law floor(-ln U * ml), ml = 1/ln m (hnsw/mod.rs:263-266); layer l holds the elements
    of level >= l;
  * candidates of an element = its efc (150) nearest elements of the layer, from the brute-force engine in approximate
    mode -- the reference selects among the efc results of its insertion search (layer.rs:352-358); with prefix=True the
    candidates come from the id prefix [0, 2^ceil(log2 i)) only, which emulates insertion order and keeps the long-range
    links early elements get in the incremental algorithm;
  * Heuristic::select (heuristic.rs:61-81,201-216) on the GPU (sdb_hnsw_select_neighbors) picks <= m_max of them;
  * every selected edge is mirrored (graph.rs:52-64) and a node that ends up with more than m_max edges is re-selected
    among its own picks plus up to (rev_factor-1)*m_max reverse edges ordered by distance (layer.rs:362-378).

The result is a valid input for the layer-walk kernel (the same CSR the reference's Hn records decode to).  It is NOT
the graph the reference would have built, so parity claims apply to the walk on a given graph; quality is measured as
recall against exact brute force next to a reference-style (oracle) graph: tests/dev/hnsw_quality.py, DESIGN.md section 6.
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


def _merge_reverse(fwd, cnt, m_max, cap=None):
    """bidirectional linking (graph.rs:52-64): every selected edge u->v also adds v->u; a node keeps its own
    selection first (nearest first) and fills up with reverse edges until m_max."""
    n = fwd.shape[0]
    valid = np.arange(fwd.shape[1])[None, :] < cnt[:, None]
    src = np.repeat(np.arange(n, dtype=np.int64), cnt)
    dst = fwd[valid].astype(np.int64)
    order = np.argsort(dst, kind="stable")
    dst_s, src_s = dst[order], src[order]
    start = np.searchsorted(dst_s, np.arange(n + 1))
    rank = np.arange(dst_s.size) - start[dst_s]
    cap = cap or 2 * m_max
    rev_cols = max(cap - m_max, m_max)  # reverse edges kept per node before the re-selection
    keep = rank < rev_cols
    rev = np.full((n, rev_cols), -1, np.int64)
    rev[dst_s[keep], rank[keep]] = src_s[keep]
    f = np.where(valid, fwd.astype(np.int64), -2)
    for j in range(rev_cols):  # drop reverse edges that duplicate a forward edge
        col = rev[:, j]
        dup = (f == col[:, None]).any(1)
        rev[dup, j] = -1
    allc = np.concatenate([np.where(valid, fwd.astype(np.int64), -1), rev], axis=1)
    ok = allc >= 0
    order2 = np.argsort(~ok, axis=1, kind="stable")[:, :cap]
    out = np.take_along_axis(allc, order2, axis=1)
    out_cnt = np.minimum(ok.sum(1), cap)
    return out, out_cnt


def build_layers(ctx, vectors_dev, n, dim, metric="EUCLIDEAN", m=16, m0=32, seed=1, batch=4096, progress=None,
                 heuristic=True, prefix=False, efc=150, rev_factor=4, levels=None):
    """vectors_dev: torch CUDA float32 tensor (n, dim).  -> (layers, entry_point, levels) with
    layers = [(row_ptr u64[n+1], col_idx u32[e]), ...] (layer 0 first), element id = row index.
    heuristic=True: candidates = 2*m_max nearest, pruned by Heuristic::select on the GPU
    (sdb_hnsw_select_neighbors), then bidirectional linking and re-selection of over-full nodes; False: plain exact
    m_max-NN lists.  prefix=True additionally restricts element i's candidates to the id prefix [0, 2^ceil(log2 i)),
    emulating insertion order (better cross-cluster links on strongly clustered data, worse on diffuse data)."""
    import ctypes as C
    import torch
    from . import _lib as L
    levels = assign_levels(n, m, seed) if levels is None else np.asarray(levels, np.int64)
    top = int(levels.max()) if n else 0
    layers = []
    dev = vectors_dev.device
    for l in range(top + 1):
        members = np.nonzero(levels >= l)[0].astype(np.int64)
        k_nb = m0 if l == 0 else m
        row_ptr = np.zeros(n + 1, np.uint64)
        if members.size <= 1:
            layers.append((row_ptr, np.zeros(0, np.uint32)))
            continue
        midx = torch.from_numpy(members).to(dev)
        sub = vectors_dev if members.size == n else vectors_dev.index_select(0, midx).contiguous()
        # candidates per element: efc nearest (the reference selects among the efc results of its insertion search,
        # layer.rs:352-358), +1 because the element itself comes back too
        kc_full = (max(2 * k_nb, min(efc, 255)) if heuristic else k_nb) + 1
        nbrs = np.zeros((members.size, k_nb), np.int64)
        counts = np.zeros(members.size, np.int64)
        o_r = torch.zeros((batch, kc_full), dtype=torch.int64, device=dev)
        o_d = torch.zeros((batch, kc_full), dtype=torch.float64, device=dev)
        o_c = torch.zeros((batch,), dtype=torch.int32, device=dev)
        s_o = torch.zeros((batch, k_nb), dtype=torch.int32, device=dev)
        s_c = torch.zeros((batch,), dtype=torch.int32, device=dev)
        # "insertion order" emulation: element i (ids are assumed shuffled) draws its candidates from the prefix
        # [0, 2^ceil(log2 i)) only, like an element inserted when the index held that many points -- early elements
        # therefore keep long-range links, which is what makes the incremental HNSW navigable across clusters.
        seg_lo = 0
        seg_hi = min(members.size, 4096) if (heuristic and prefix) else members.size
        while seg_lo < members.size:
            torch.cuda.current_stream().synchronize()
            col = VectorColumn(ctx, dim, metric, "F32", capacity=seg_hi)
            col.append_device(sub.data_ptr(), seg_hi)
            col.finalize()
            col.set_exact(False)  # candidate generation does not need the exactness proof
            kc = min(kc_full, seg_hi)
            for b0 in range(seg_lo, seg_hi, batch):
                b1 = min(seg_hi, b0 + batch)
                q = sub[b0:b1].to(torch.float64).contiguous()
                torch.cuda.current_stream().synchronize()  # torch wrote q on ITS stream; the library runs on its own
                o_rv = o_r.view(-1)[: batch * kc].view(batch, kc)
                o_dv = o_d.view(-1)[: batch * kc].view(batch, kc)
                col.knn_device(q.data_ptr(), b1 - b0, kc, 0, o_rv.data_ptr(), o_dv.data_ptr(), o_c.data_ptr())
                if heuristic:
                    L.check(L.lib().sdb_hnsw_select_neighbors(ctx.h, C.c_void_p(sub.data_ptr()), dim, L.METRIC[metric.upper()],
                                                              b0, b1 - b0, C.c_void_p(o_rv.data_ptr()), C.c_void_p(o_c.data_ptr()),
                                                              kc, k_nb, 1, C.c_void_p(s_o.data_ptr()), C.c_void_p(s_c.data_ptr())))
                    nbrs[b0:b1] = s_o[: b1 - b0].cpu().numpy()
                    counts[b0:b1] = s_c[: b1 - b0].cpu().numpy()
                else:
                    r = o_rv[: b1 - b0].cpu().numpy()
                    cnt = o_c[: b1 - b0].cpu().numpy().astype(np.int64)
                    valid = np.arange(kc)[None, :] < cnt[:, None]
                    keep = valid & (r != (b0 + np.arange(b1 - b0))[:, None])  # drop self, keep nearest-first order
                    order = np.argsort(~keep, axis=1, kind="stable")[:, :k_nb]
                    nbrs[b0:b1, : order.shape[1]] = np.take_along_axis(r, order, axis=1)
                    counts[b0:b1] = np.minimum(keep.sum(1), k_nb)
                if progress:
                    progress(l, b1, members.size)
            col.close()
            seg_lo, seg_hi = seg_hi, min(members.size, 2 * seg_hi)
        if heuristic:
            # bidirectional edges, then re-select every over-full node among (own selection + reverse edges) ordered by
            # distance -- layer.rs:362-378
            if progress:
                progress(l, -1, float(counts.mean()))
            union, ucnt = _merge_reverse(nbrs, counts, k_nb, cap=rev_factor * k_nb)
            if progress:
                progress(l, -2, float(ucnt.mean()))
            u_dev = torch.from_numpy(np.maximum(union, 0).astype(np.int64)).to(dev)
            c_dev = torch.from_numpy(ucnt.astype(np.int32)).to(dev)
            r_o = torch.zeros((members.size, k_nb), dtype=torch.int32, device=dev)
            r_c = torch.zeros((members.size,), dtype=torch.int32, device=dev)
            torch.cuda.current_stream().synchronize()
            L.check(L.lib().sdb_hnsw_select_neighbors(ctx.h, C.c_void_p(sub.data_ptr()), dim, L.METRIC[metric.upper()], 0,
                                                      members.size, C.c_void_p(u_dev.data_ptr()), C.c_void_p(c_dev.data_ptr()),
                                                      rev_factor * k_nb, k_nb, 0, C.c_void_p(r_o.data_ptr()), C.c_void_p(r_c.data_ptr())))
            nbrs = r_o.cpu().numpy().astype(np.int64)
            counts = r_c.cpu().numpy().astype(np.int64)
        deg = np.zeros(n, np.int64)
        deg[members] = counts
        row_ptr[1:] = np.cumsum(deg)
        col_idx = np.zeros(int(row_ptr[-1]), np.uint32)
        flat = members[np.maximum(nbrs, 0)]  # local -> global element ids
        mask = np.arange(k_nb)[None, :] < counts[:, None]
        col_idx[:] = flat[mask].astype(np.uint32)
        layers.append((row_ptr, col_idx))
    entry = int(np.argmax(levels)) if n else -1
    return layers, entry, levels


def build_incremental(ctx, x, metric="COSINE", m=16, m0=32, efc=150, seed=1, growth=0.25, boot_min=65536,
                      search_chunk=1 << 16, rev_extra=32, progress=None, settle=True):
    """Batched TRUE insertion: the reference inserts one element at a time -- search the current graph
    with efc, select <= m_max neighbours with the heuristic, link both ways, re-select over-full neighbours
    (hnsw/mod.rs:297-377, hnsw/layer.rs:342-387).  Here the same four steps run for a whole BATCH of new elements
    against the graph built so far, with the layer-walk kernel itself as the insertion search
    (sdb_hnsw_load_device + sdb_hnsw_search_device), Heuristic::select on the GPU (sdb_hnsw_select_neighbors[_ids]) and
    the linking as a handful of device-side scatter operations.  Batches grow geometrically (`growth` x the current
    size), so every element is inserted into a graph at least 1 / (1 + growth) of the size it would have seen in the
    serial algorithm; elements of one batch do not see each other.

    Elements are first re-ordered by level (highest first, random inside a level -- ids are assumed exchangeable), so
    the upper layers and a bootstrap prefix are complete before the bulk of layer 0 arrives; that prefix (every element
    of level >= 1, at least `boot_min`) is built by the kNN batch builder above.

    x: torch CUDA float32 (n, dim).  Returns a dict: x (re-ordered copy, device), order (new -> original index, numpy),
    layers_dev [(row_ptr int64 (n+1), col_idx int32)] layer 0 first (device tensors, CSR over NEW ids), entry (new id),
    levels (new order)."""
    import ctypes as C
    import torch
    from . import _lib as L
    n, dim = x.shape
    dev = x.device
    levels = assign_levels(n, m, seed)
    order = np.argsort(-levels, kind="stable")
    levels = levels[order]
    x = x.index_select(0, torch.from_numpy(order).to(dev)).contiguous()
    n_up = int((levels >= 1).sum())
    n_boot = min(n, max(n_up, boot_min))
    torch.cuda.synchronize()
    boot_layers, entry, _ = build_layers(ctx, x[:n_boot], n_boot, dim, metric, m=m, m0=m0, seed=seed, prefix=True, efc=efc,
                                         levels=levels[:n_boot], progress=progress)
    n_layers = len(boot_layers)
    # upper layers are final: CSR over all n ids (rows >= n_boot are empty)
    upper = []
    for l in range(1, n_layers):
        rp, ci = boot_layers[l]
        rp_full = np.full(n + 1, rp[-1], np.int64)
        rp_full[: n_boot + 1] = rp.astype(np.int64)
        upper.append((torch.from_numpy(rp_full).to(dev), torch.from_numpy(ci.astype(np.int32) if ci.size else np.zeros(1, np.int32)).to(dev)))
    # layer 0 as fixed-width adjacency while it grows
    adj0 = torch.full((n, m0), -1, dtype=torch.int32, device=dev)
    deg0 = torch.zeros(n, dtype=torch.int32, device=dev)
    rp0, ci0 = boot_layers[0]
    d0 = np.diff(rp0.astype(np.int64))
    rows = np.repeat(np.arange(n_boot), d0)
    cols = np.arange(ci0.size) - np.repeat(rp0[:-1].astype(np.int64), d0)
    adj0[torch.from_numpy(rows).to(dev), torch.from_numpy(cols).to(dev)] = torch.from_numpy(ci0.astype(np.int32)).to(dev)
    deg0[:n_boot] = torch.from_numpy(d0.astype(np.int32)).to(dev)
    mcode = L.METRIC[metric.upper()]
    col_ids = torch.arange(m0, device=dev, dtype=torch.int32)[None, :]

    def csr0():
        rp = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        rp[1:] = torch.cumsum(deg0.to(torch.int64), 0)
        ci = adj0[col_ids < deg0[:, None]]
        if ci.numel() == 0:
            ci = torch.zeros(1, dtype=torch.int32, device=dev)
        return rp, ci.contiguous()

    def search_select(lo, hi, drop_self):
        """insertion search of elements [lo, hi) on the current graph + Heuristic::select -> (sel (b, m0) int32, count)"""
        b = hi - lo
        rp, ci = csr0()
        lay = [(rp, ci)] + upper
        RP = (C.c_void_p * n_layers)(*[t[0].data_ptr() for t in lay])
        CI = (C.c_void_p * n_layers)(*[t[1].data_ptr() for t in lay])
        h = C.c_void_p()
        torch.cuda.synchronize()
        L.check(L.lib().sdb_hnsw_load_device(ctx.h, dim, mcode, n, C.c_void_p(x.data_ptr()), n_layers, RP, CI, int(entry), C.byref(h)))
        sel = torch.empty((b, m0), dtype=torch.int32, device=dev)
        scnt = torch.empty((b,), dtype=torch.int32, device=dev)
        try:
            for c0 in range(0, b, search_chunk):
                c1 = min(b, c0 + search_chunk)
                nqc = c1 - c0
                cand = torch.empty((nqc, efc), dtype=torch.int64, device=dev)
                cdist = torch.empty((nqc, efc), dtype=torch.float64, device=dev)
                ccnt = torch.empty((nqc,), dtype=torch.int32, device=dev)
                torch.cuda.synchronize()
                L.check(L.lib().sdb_hnsw_search_device(h, C.c_void_p(x[lo + c0].data_ptr()), nqc, efc, efc,
                                                       C.c_void_p(cand.data_ptr()), C.c_void_p(cdist.data_ptr()),
                                                       C.c_void_p(ccnt.data_ptr())))
                if drop_self:  # the element is part of the graph by now: take it out of its own candidate list (order kept)
                    me = torch.arange(lo + c0, lo + c1, device=dev, dtype=torch.int64)[:, None]
                    live = torch.arange(efc, device=dev)[None, :] < ccnt[:, None]
                    is_me = (cand == me) & live
                    key = torch.arange(efc, device=dev)[None, :] + is_me.to(torch.int64) * efc  # self goes last
                    cand = torch.gather(cand, 1, torch.argsort(key, dim=1, stable=True)).contiguous()
                    ccnt = (ccnt - is_me.sum(1).to(torch.int32)).contiguous()
                    torch.cuda.synchronize()
                L.check(L.lib().sdb_hnsw_select_neighbors(ctx.h, C.c_void_p(x.data_ptr()), dim, mcode, lo + c0, nqc,
                                                          C.c_void_p(cand.data_ptr()), C.c_void_p(ccnt.data_ptr()), efc, m0, 1,
                                                          C.c_void_p(sel[c0:c1].data_ptr()), C.c_void_p(scnt[c0:c1].data_ptr())))
                del cand, cdist, ccnt
        finally:
            L.lib().sdb_hnsw_destroy(h)
        return sel, scnt

    def link(lo, hi, sel, scnt, dedup):
        """forward edges of [lo, hi) := sel; reverse edges into the targets, re-selecting the nodes that overflow m0"""
        b = hi - lo
        valid = col_ids < scnt[:, None]
        adj0[lo:hi] = torch.where(valid, sel, torch.full_like(sel, -1))
        deg0[lo:hi] = scnt
        src = torch.arange(lo, hi, device=dev, dtype=torch.int32)[:, None].expand(b, m0)[valid]
        dst = sel[valid].to(torch.int64)
        if dedup and dst.numel():  # the target may hold this edge already (second pass over the same elements)
            keep = torch.ones(dst.numel(), dtype=torch.bool, device=dev)
            for e0 in range(0, dst.numel(), 1 << 22):
                e1 = min(dst.numel(), e0 + (1 << 22))
                keep[e0:e1] = ~(adj0[dst[e0:e1]] == src[e0:e1, None]).any(1)
            src, dst = src[keep], dst[keep]
        if not dst.numel():
            return
        dst_s, perm = torch.sort(dst, stable=True)
        src_s = src[perm]
        uniq, counts = torch.unique_consecutive(dst_s, return_counts=True)
        start = torch.cumsum(counts, 0) - counts
        rank = torch.arange(dst_s.numel(), device=dev) - torch.repeat_interleave(start, counts)
        pos = deg0[dst_s].to(torch.int64) + rank
        ok = pos < m0
        adj0[dst_s[ok], pos[ok]] = src_s[ok]
        new_deg = deg0[uniq].to(torch.int64) + counts
        deg0[uniq] = torch.clamp(new_deg, max=m0).to(torch.int32)
        over = new_deg > m0
        ov_nodes = uniq[over]
        n_ov = int(ov_nodes.numel())
        if n_ov:
            # re-select the over-full nodes among their m0 current edges + the reverse edges that did not fit
            kc = m0 + rev_extra
            union = torch.full((n_ov, kc), 0, dtype=torch.int64, device=dev)
            union[:, :m0] = adj0[ov_nodes].to(torch.int64)
            node_slot = torch.full((n,), -1, dtype=torch.int64, device=dev)
            node_slot[ov_nodes] = torch.arange(n_ov, device=dev)
            ex_dst, ex_src, ex_rank = dst_s[~ok], src_s[~ok], (pos[~ok] - m0)
            keep = ex_rank < rev_extra
            union[node_slot[ex_dst[keep]], m0 + ex_rank[keep]] = ex_src[keep].to(torch.int64)
            ucnt = (m0 + torch.clamp(new_deg[over] - m0, max=rev_extra)).to(torch.int32)
            out = torch.empty((n_ov, m0), dtype=torch.int32, device=dev)
            ocnt = torch.empty((n_ov,), dtype=torch.int32, device=dev)
            ids32 = ov_nodes.to(torch.int32).contiguous()
            torch.cuda.synchronize()
            L.check(L.lib().sdb_hnsw_select_neighbors_ids(ctx.h, C.c_void_p(x.data_ptr()), dim, mcode, C.c_void_p(ids32.data_ptr()),
                                                          n_ov, C.c_void_p(union.data_ptr()), C.c_void_p(ucnt.data_ptr()), kc, m0, 0,
                                                          C.c_void_p(out.data_ptr()), C.c_void_p(ocnt.data_ptr())))
            v2 = col_ids < ocnt[:, None]
            adj0[ov_nodes] = torch.where(v2, out, torch.full_like(out, -1))
            deg0[ov_nodes] = ocnt

    n_cur = n_boot
    while n_cur < n:
        b = int(min(n - n_cur, max(4096, int(n_cur * growth))))
        sel, scnt = search_select(n_cur, n_cur + b, False)
        link(n_cur, n_cur + b, sel, scnt, False)
        if settle:
            # Elements of one batch did not see each other.  Second pass: the same insertion search on the graph that now
            # holds the whole batch, so close neighbours that arrived together get linked (serial insertion would have
            # linked the later one to the earlier one).
            sel, scnt = search_select(n_cur, n_cur + b, True)
            link(n_cur, n_cur + b, sel, scnt, True)
        if progress:
            progress(0, n_cur + b, n)
        n_cur += b
    rp, ci = csr0()
    return {"x": x, "order": order, "layers_dev": [(rp, ci)] + upper, "entry": int(entry), "levels": levels}
