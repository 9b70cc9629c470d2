"""GPU side of the sharded path on ONE device: two row shards searched separately, blocks packed as in the
all-gather buffer, merged by sdb_topk_merge_device -> identical to the unsharded search and to the oracle."""
import numpy as np
import pytest

from oracle import pyoracle as O

pytestmark = pytest.mark.gpu


def test_sharded_search_plus_merge_equals_unsharded():
    import torch
    from surrealdb_b200 import Context, VectorColumn
    from surrealdb_b200.engine import shard_block_layout, topk_merge_device
    from surrealdb_b200.sharding import shard_range
    from surrealdb_b200.synthetic import gen_f32
    rows, dim, nq, k, world = 20000, 64, 40, 10, 3
    ctx = Context(0)
    dev = torch.device("cuda", 0)
    corpus = gen_f32(21, 0, rows * dim).reshape(rows, dim)
    corpus[5000:5004] = corpus[100:104]  # cross-shard exact ties must resolve by global row
    queries = gen_f32(22, 0, nq * dim).reshape(nq, dim).astype(np.float64)
    qd = torch.from_numpy(queries).to(dev)
    torch.cuda.synchronize()  # torch streams and the library stream are not ordered with each other
    off_rows, off_dist, off_cnt, blk = shard_block_layout(nq, k)
    gathered = torch.zeros(world * blk, dtype=torch.uint8, device=dev)
    for r in range(world):
        base, n_local = shard_range(rows, world, r)
        col = VectorColumn(ctx, dim, "COSINE", "F32", capacity=n_local)
        col.append(corpus[base:base + n_local])
        col.finalize()
        p = gathered.data_ptr() + r * blk
        col.knn_device(qd.data_ptr(), nq, k, base, p + off_rows, p + off_dist, p + off_cnt)
    f_rows = torch.zeros((nq, k), dtype=torch.int64, device=dev)
    f_dist = torch.zeros((nq, k), dtype=torch.float64, device=dev)
    f_cnt = torch.zeros((nq,), dtype=torch.int32, device=dev)
    gp = gathered.data_ptr()
    topk_merge_device(ctx, world, nq, k, gp + off_rows, gp + off_dist, gp + off_cnt, f_rows.data_ptr(),
                      f_dist.data_ptr(), f_cnt.data_ptr(), stride_rows=blk // 8, stride_dist=blk // 8,
                      stride_counts=blk // 4)
    torch.cuda.synchronize()
    fr, fd = f_rows.cpu().numpy(), f_dist.cpu().numpy()
    for q in range(nq):
        r, d = O.knn_topk(corpus, queries[q], "cosine", k)
        assert list(fr[q]) == list(r) and fd[q].tobytes() == d.tobytes()
