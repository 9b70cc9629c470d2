"""Counter-based synthetic data generator (numpy mirror of csrc/gen.cu): value(seed, i) in [-1, 1) on a
2^-23 grid.  Used by bench.py / smoke() to produce host-side queries that are bit-identical to what
sdb_corpus_append_synthetic generates in HBM."""
import numpy as np

_M1, _M2 = np.uint64(0xBF58476D1CE4E5B9), np.uint64(0x94D049BB133111EB)


def gen_f32(seed, first, n):
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) + np.uint64(first)
        z = np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + idx + np.uint64(0x632BE59BD9B4E019)
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    m = (z >> np.uint64(40)).astype(np.float32)
    return m * np.float32(1.0 / 8388608.0) - np.float32(1.0)
