"""Pins the oracle's Vec<Number> / typed-vector arithmetic against every known-answer vector the
reference's own tests hold for this path."""
import math

import numpy as np
import pytest

from oracle import pyoracle as O

A, B = [1.0, 2.0, 3.0], [2.0, 3.0, 4.0]

# core/idx/trees/vector.rs:723-772 -- `test_distance(dist, [1,2,3], [2,3,4], res)` asserts BOTH the
# generic Distance::compute (Vec<Number>) and the typed F64 Distance::calculate equal `res`.
VECTOR_RS_KATS = [
    ("chebyshev", 1.0),
    ("cosine_distance", 0.007416666029069652),
    ("euclidean", 1.7320508075688772),
    ("hamming", 3.0),
    ("jaccard", 0.5),
    ("manhattan", 3.0),
    ("minkowski", 1.4422495703074083),
    ("pearson", 1.0),
]


@pytest.mark.parametrize("name,res", VECTOR_RS_KATS)
def test_vector_rs_kats_generic(name, res):
    st, v = O.num_metric(name, A, B, p=3.0)
    assert st == 0
    assert O.num_cmp(v, res) == 0, (name, v, res)  # assert_eq!(dist.compute(..), res.into())


def test_vector_rs_kats_typed_f64():
    assert O.vec_distance_f64("cosine", A, B) == 0.007416666029069652
    assert O.vec_distance_f64("euclidean", A, B) == 1.7320508075688772
    # typed F32 path agrees to f32 precision on the same literals
    assert abs(O.vec_distance_f32("cosine", A, B) - 0.007416666029069652) < 1e-6
    assert abs(O.vec_distance_f32("euclidean", A, B) - 1.7320508075688772) < 1e-6
    assert O.vec_distance_f32("manhattan", A, B) == 3.0
    assert O.vec_distance_f32("chebyshev", A, B) == 1.0
    assert O.vec_distance_f32("hamming", A, B) == 3.0


NAN = float("nan")

# surrealdb/core/tests/function.rs:3304-3640 (SurrealQL literals: bare ints are Number::Int)
FUNCTION_RS = [
    ("cosine_similarity", [1, 2, 3], [1, 2, 3], 1.0),
    ("cosine_similarity", [1, 2, 3], [-1, -2, -3], -1.0),
    ("cosine_similarity", [NAN, 1, 2, 3], [NAN, 1, 2, 3], NAN),
    ("cosine_similarity", [10, 50, 200], [400, 100, 20], 0.15258215962441316),
    ("jaccard", [1, 2, 3], [3, 2, 1], 1.0),
    ("jaccard", [1, 2, 3], [-3, -2, -1], 0.0),
    ("jaccard", [1, -2, 3, -4], [4, 3, 2, 1], 0.3333333333333333),
    ("jaccard", [NAN, 1, 2, 3], [NAN, 2, 3, 4], 0.6),
    ("jaccard", [0, 1, 2, 5, 6], [0, 2, 3, 4, 5, 7, 9], 0.3333333333333333),
    ("pearson", [1, 2, 3, 4, 5], [1, 2.5, 3.5, 4.2, 5.1], 0.9894065340659606),
    ("pearson", [NAN, 1, 2, 3, 4, 5], [NAN, 1, 2.5, 3.5, 4.2, 5.1], NAN),
    ("pearson", [1, 2, 3], [1, 5, 7], 0.9819805060619659),
    ("euclidean", [1, 2, 3], [1, 2, 3], 0.0),
    ("euclidean", [NAN, 2, 3], [-1, NAN, -3], NAN),
    ("euclidean", [1, 2, 3], [-1, -2, -3], 7.483314773547883),
    ("euclidean", [10, 50, 200], [400, 100, 20], 432.43496620879307),
    ("euclidean", [10, 20, 15, 10, 5], [12, 24, 18, 8, 7], 6.082762530298219),
    ("manhattan", [1, 2, 3], [4, 5, 6], 9),
    ("manhattan", [1, 2, 3], [-4, -5, -6], 21),
    ("manhattan", [1.1, 2, 3.3], [4, 5.5, 6.6], 9.7),
    ("manhattan", [NAN, 1, 2, 3], [NAN, 4, 5, 6], NAN),
    ("manhattan", [10, 20, 15, 10, 5], [12, 24, 18, 8, 7], 13),
    ("hamming", [1, 2, 2], [1, 2, 3], 1),
    ("hamming", [-1, -2, -3], [-2, -2, -2], 2),
    ("hamming", [1.1, 2.2, -3.3], [1.1, 2, -3.3], 1),
    ("hamming", [NAN, 1, 2, 3], [NAN, 1, 2, 3], 0),
    ("hamming", [0, 0, 0, 0, 0, 1], [0, 0, 0, 0, 1, 0], 2),
    ("chebyshev", [1, 2, 3], [4, 5, 6], 3.0),
    ("chebyshev", [-1, -2, -3], [-4, -5, -6], 3.0),
    ("chebyshev", [1.1, 2.2, 3], [4, 5.5, 6.6], 3.5999999999999996),
    ("chebyshev", [NAN, 1, 2, 3], [NAN, 4, 5, 6], 3.0),
    ("chebyshev", [2, 4, 5, 3, 8, 2], [3, 1, 5, -3, 7, 2], 6.0),
    ("dot", [1, 2, 3], [1, 2, 3], 14),
]


@pytest.mark.parametrize("name,a,b,res", FUNCTION_RS)
def test_function_rs_kats(name, a, b, res):
    st, v = O.num_metric(name, a, b)
    assert st == 0
    if isinstance(res, float) and math.isnan(res):
        assert math.isnan(v)
    else:
        assert v == res and type(v) is type(res), (name, v, res)


MINKOWSKI = [
    ([1, 2, 3], [4, 5, 6], 3, 4.3267487109222245),
    ([-1, -2, -3], [-4, -5, -6], 3, 4.3267487109222245),
    ([1.1, 2.2, 3], [4, 5.5, 6.6], 3, 4.747193170917638),
    ([10, 20, 15, 10, 5], [12, 24, 18, 8, 7], 1, 13.0),
    ([10, 20, 15, 10, 5], [12, 24, 18, 8, 7], 2, 6.082762530298219),
]


@pytest.mark.parametrize("a,b,p,res", MINKOWSKI)
def test_function_rs_minkowski(a, b, p, res):
    st, v = O.num_metric("minkowski", a, b, p=float(p))
    assert st == 0 and v == res


def test_function_rs_magnitude():
    # function.rs: vector::magnitude([]) 0f, [1] 1f, [5] 5f, [1,2,3,3,3,4,5] 8.54400374531753
    assert O.num_magnitude([]) == 0.0
    assert O.num_magnitude([1]) == 1.0
    assert O.num_magnitude([5]) == 5.0
    assert O.num_magnitude([1, 2, 3, 3, 3, 4, 5]) == 8.54400374531753


@pytest.mark.parametrize("name", ["dot", "cosine_similarity", "cosine_distance", "euclidean", "manhattan",
                                  "hamming", "chebyshev", "pearson", "minkowski"])
def test_dimension_mismatch_is_error(name):
    # "The two vectors must be of the same dimension."  fnc/util/math/vector.rs:23-32
    assert O.num_metric(name, [1, 2, 3], [4, 5])[0] == 1
    assert O.num_metric(name, [1, 2], [4, 5, 5])[0] == 1


def test_number_ordering():
    # val/number.rs:620-680: -0.0 == 0.0; NaN (positive) sorts after +inf; Int vs Float by value
    assert O.num_cmp(-0.0, 0.0) == 0
    assert O.num_cmp(float("inf"), NAN) == -1
    assert O.num_cmp(1, 1.0) == 0
    assert O.num_cmp(1, 1.5) == -1
    assert O.num_cmp(2, 1.5) == 1
    assert O.num_cmp(10**18, float("inf")) == -1
    assert O.num_cmp(-5, float("-inf")) == 1


def test_fast_path_equals_number_path():
    rng = np.random.default_rng(7)
    for d in (1, 3, 8, 127, 768):
        a = rng.uniform(-20, 20, d)
        b = rng.uniform(-20, 20, d)
        st, v = O.num_metric("cosine_distance", list(a), list(b))
        assert v == O.f64_cosine_distance(a, b)
        st, v = O.num_metric("euclidean", list(a), list(b))
        assert v == O.f64_euclidean(a, b)


def test_f32_lane_order_documented_shape():
    # The 8-lane order is "parity unpinned" (ndarray not vendored).  What we can pin: for n < 8 it
    # degenerates to the sequential order, and the result is within the f32 rounding envelope.
    rng = np.random.default_rng(3)
    a = rng.uniform(-1, 1, 7).astype(np.float32)
    b = rng.uniform(-1, 1, 7).astype(np.float32)
    s = np.float32(0)
    for x, y in zip(a, b):
        s = np.float32(s + np.float32(x * y))
    import ctypes as C
    got = O.lib().orc_nd_dot_f32(a.ctypes.data_as(C.POINTER(C.c_float)), b.ctypes.data_as(C.POINTER(C.c_float)),
                                 C.c_size_t(7))
    assert np.float32(got) == s
    a = rng.uniform(-1, 1, 1536).astype(np.float32)
    b = rng.uniform(-1, 1, 1536).astype(np.float32)
    got = O.vec_distance_f32("cosine", a, b)
    ref = 1.0 - float(np.dot(a.astype(np.float64), b.astype(np.float64))) / (
        np.linalg.norm(a.astype(np.float64)) * np.linalg.norm(b.astype(np.float64)))
    assert abs(got - ref) <= 1e-5 * max(1.0, abs(ref))


@pytest.mark.parametrize("metric,num", [("manhattan", "manhattan"), ("chebyshev", "chebyshev"), ("hamming", "hamming"),
                                        ("pearson", "pearson"), ("cosine", "cosine_distance"), ("euclidean", "euclidean")])
def test_all_float_fast_metrics_equal_number_path(metric, num):
    rng = np.random.default_rng(17)
    for d in (1, 2, 9, 128):
        a = rng.uniform(-20, 20, d)
        b = rng.uniform(-20, 20, d)
        if metric == "hamming":
            b[::2] = a[::2]
        st, v = O.num_metric(num, list(a), list(b))
        got = O.f64_metric(metric, a, b)
        assert (math.isnan(got) and math.isnan(float(v))) or got == float(v), (metric, d, got, v)
    # the reference's own KATs through the fast path (function.rs)
    assert O.f64_metric("manhattan", [1.1, 2, 3.3], [4, 5.5, 6.6]) == 9.7
    assert O.f64_metric("chebyshev", [1.1, 2.2, 3], [4, 5.5, 6.6]) == 3.5999999999999996
    assert O.f64_metric("pearson", [1, 2, 3, 4, 5], [1, 2.5, 3.5, 4.2, 5.1]) == 0.9894065340659606
    assert O.f64_metric("hamming", [1.1, 2.2, -3.3], [1.1, 2, -3.3]) == 1.0
