"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/sdbgpu.h declares, and refuses to compute without a GPU (no silent CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    import __graft_entry__ as g
    so = os.path.join(ROOT, "surrealdb_b200", "csrc", "libsdbgpu.so")
    if not os.path.exists(so):
        g.build()
    return so


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "sdbgpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sdb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    so = _ensure_built()
    lib = C.CDLL(so)
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/sdbgpu.h but not exported"


def test_python_binding_lists_the_same_symbols():
    _ensure_built()
    from surrealdb_b200 import _lib
    assert sorted(_lib.ABI_SYMBOLS) == header_symbols()
    _lib.lib()  # argtypes resolve


def test_no_cpu_fallback_without_gpu():
    _ensure_built()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from surrealdb_b200 import Context, SdbError
    with pytest.raises(SdbError) as e:
        Context(0)
    assert "SDB_ECUDA" in str(e.value)


def test_pass_schedule_covers_every_tile_once():
    # host logic of the threshold-refinement schedule (csrc/api.cu build_passes / internal.cuh pass_tile)
    R, TILE = 8, 256

    def build(n_rows, cap):
        T = (n_rows + TILE - 1) // TILE
        if T == 0:
            return []
        max0 = cap // TILE
        stride = 1
        while (T + stride - 1) // stride > max0:
            stride *= R
        passes = [(stride, 0, (T + stride - 1) // stride)]
        s = stride // R
        while s >= 1:
            M = (T + s - 1) // s
            passes.append((s, 1, M - (M + R - 1) // R))
            if s == 1:
                break
            s //= R
        return passes

    def tile(p, w):
        stride, excl, _ = p
        i = (w // (R - 1)) * R + (w % (R - 1)) + 1 if excl else w
        return i * stride

    for n_rows in (1, 255, 256, 257, 4096 * 256, 4096 * 256 + 1, 1_000_000, 10_000_000):
        T = (n_rows + TILE - 1) // TILE
        seen = []
        for p in build(n_rows, 4096):
            seen += [tile(p, w) for w in range(p[2])]
        assert sorted(seen) == list(range(T)), n_rows
        assert build(n_rows, 4096)[0][2] * TILE <= 4096


def _build_c_driver(tmp_path):
    import subprocess
    so = _ensure_built()
    exe = str(tmp_path / "abi_driver")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "abi_driver.c"), "-o", exe, so,
                           "-Wl,-rpath," + os.path.dirname(so)])
    return exe


def test_header_is_plain_c_and_a_c_program_links(tmp_path):
    import subprocess
    import torch
    exe = _build_c_driver(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    if not torch.cuda.is_available():
        assert out.stdout.startswith("NO_GPU"), out.stdout  # refuses loudly, no CPU fallback


@pytest.mark.gpu
def test_c_driver_runs_the_three_paths_on_the_gpu(tmp_path):
    import subprocess
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    exe = _build_c_driver(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ABI_DRIVER_OK" in out.stdout, (out.stdout, out.stderr)


def test_stats_struct_layout_matches_the_header(tmp_path):
    # sdb_knn_last_stats writes through a caller-provided struct: the ctypes mirror must have the C layout exactly
    import ctypes as C
    import subprocess
    from surrealdb_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sdbgpu.h"\n'
                   'int main(){printf("%zu", sizeof(sdb_knn_stats));\n'
                   + "".join(f'printf(" %zu", offsetof(sdb_knn_stats, {name}));\n' for name, _ in _lib.KnnStats._fields_)
                   + 'return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    vals = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert vals[0] == C.sizeof(_lib.KnnStats)
    assert vals[1:] == [getattr(_lib.KnnStats, name).offset for name, _ in _lib.KnnStats._fields_]
