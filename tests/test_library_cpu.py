"""CPU-side checks of libldpc_cuda: the shared library builds/loads, exports every symbol that
include/ldpc_cuda.h declares, and fails loudly (no fallback) when no GPU is present."""
import ctypes as C
import os
import re

import pytest

from ldpc_erasure_codes_b200 import _lib
from ldpc_erasure_codes_b200.build import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build()
    return _lib.load()


def test_header_symbols_are_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "ldpc_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b((?:ldpc|rs)_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_abi_version(lib):
    assert lib.ldpc_cuda_abi_version() == 2


def test_argument_errors_do_not_need_a_gpu(lib):
    h = C.c_void_p()
    assert lib.ldpc_ctx_create(C.byref(h), None, 1, 60, 0, 16) == -1       # S not a multiple of 16
    assert b"multiple of 16" in lib.ldpc_last_error_string()
    assert lib.ldpc_ctx_create(C.byref(h), None, 7, 64, 0, 16) == -1       # unknown built-in code
    assert lib.ldpc_ctx_create(C.byref(h), b"/nonexistent.mat", 0, 64, 0, 16) == -2
    assert lib.ldpc_ctx_create(C.byref(h), __file__.encode(), 0, 64, 0, 16) == -3  # not a MAT file
    assert lib.ldpc_decode(None, None, None, None, None, 50, 0, 1, None) == -1


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.ldpc_ctx_create(C.byref(h), None, 1, 64, 0, 16)
    assert rc == -4 and not h.value, "context creation must fail without a CUDA device"
    from ldpc_erasure_codes_b200.codec import LdpcCodec
    with pytest.raises(RuntimeError):
        LdpcCodec(code=1)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "ldpc_erasure_codes_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                for pat in (r"import\s+oracle", r"from\s+oracle", r"oracle/", r"libldpc_oracle", r"orc_"):
                    assert not re.search(pat, txt), f"{f} reaches into the oracle ({pat})"


# ---------------------------------------------------------------------------- the C++ MAT-v5 loader, on CPU
def _scipy_csr(path):
    from oracle import oracle as orc
    return orc.Code.from_mat(path)


@pytest.mark.parametrize("name", ["n2000_k1000", "n2040_k1530", "n4000_k2000"])
def test_cpp_loader_reads_exported_codes(name):
    from ldpc_erasure_codes_b200.codec import read_h_file
    path = os.path.join(ROOT, "ldpc_erasure_codes_b200", "codes", name + ".mat")
    m, n, tri, rp, ci = read_h_file(path)
    ref = _scipy_csr(path)
    assert (m, n) == (ref.m, ref.n) and tri
    assert (rp == ref.row_ptr).all() and (ci == ref.col_idx).all()


@pytest.mark.skipif(not os.path.isdir("/root/reference/Matlab"), reason="reference not mounted")
@pytest.mark.parametrize("fname", ["n2000_k1000_no6cycles_triangleForm_OpenCL_H.mat",
                                   "n2040_k1530_irreg_H_no6cycles_triangleForm.mat",
                                   "n4000_k2000_no6cycles_triangleForm.mat"])
def test_cpp_loader_reads_the_reference_files_themselves(fname):
    """The loader is fed the reference's own committed .mat files (MATLAB-written, different zlib stream)."""
    from ldpc_erasure_codes_b200.codec import read_h_file
    path = os.path.join("/root/reference/Matlab", fname)
    m, n, tri, rp, ci = read_h_file(path)
    ref = _scipy_csr(path)
    assert (m, n) == (ref.m, ref.n) and tri
    assert (rp == ref.row_ptr).all() and (ci == ref.col_idx).all()


def test_cpp_loader_rejects_non_sparse_garbage(tmp_path):
    import scipy.io as sio
    import numpy as np
    from ldpc_erasure_codes_b200 import _lib as L
    p = tmp_path / "x.mat"
    sio.savemat(str(p), {"other": np.eye(3)})
    lib = L.load()
    dims = (C.c_int32 * 4)()
    assert lib.ldpc_read_h_file(str(p).encode(), C.byref(dims), None, None) == -3      # no H_sparse inside
    # a full (non-sparse) 0/1 matrix named H_sparse is accepted too
    H = np.zeros((2, 5)); H[0, [0, 1, 3]] = 1; H[1, [1, 2, 4]] = 1
    sio.savemat(str(p), {"H_sparse": H})
    assert lib.ldpc_read_h_file(str(p).encode(), C.byref(dims), None, None) == 0
    assert list(dims) == [2, 5, 6, 1]
