"""Code design front-end: girth-8 triangular-form H generator and short-cycle checker of
libldpc_cuda (csrc/hgen.cpp), the counterparts of Matlab/Hgen_irregularDegree_no6cycles_systematic_encoding.m and
Matlab/Cycle_Finder_length{4_fromroot,6}.m.  Host-only entry points; matrices come back as scipy CSR / are saved as the
MAT-v5 `H_sparse` files the loader (and the reference's MATLAB) reads."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _check(rc):
    if rc != 0:
        raise _lib.LdpcCudaError(rc, _lib.load().ldpc_h_last_error_string().decode())


def generate(deg_c_prof, deg_v_prof, seed=1, max_tries=100):
    """deg_*_prof: [(count, degree), ...] with degrees in descending order (the script's deg_c_prof / deg_v_prof).
    Returns (H as scipy.sparse.csr_matrix of float64 ones, tries used)."""
    import scipy.sparse as sp
    lib = _lib.load()
    dc = np.ascontiguousarray(deg_c_prof, dtype=np.int32).reshape(-1, 2)
    dv = np.ascontiguousarray(deg_v_prof, dtype=np.int32).reshape(-1, 2)
    m = int(dc[:, 0].sum())
    cap = int((dc[:, 0] * dc[:, 1]).sum()) + 2 * m
    dims = (C.c_int32 * 4)()
    row_ptr = np.zeros(m + 1, dtype=np.int32)
    col_idx = np.zeros(cap, dtype=np.int32)
    tries = C.c_int32(0)
    _check(lib.ldpc_h_generate(dc.ctypes.data, len(dc), dv.ctypes.data, len(dv), seed, max_tries, C.byref(dims),
                               row_ptr.ctypes.data, col_idx.ctypes.data, cap, C.byref(tries)))
    mm, n, nnz = dims[0], dims[1], dims[2]
    H = sp.csr_matrix((np.ones(nnz), col_idx[:nnz].copy(), row_ptr.copy()), shape=(mm, n))
    return H, tries.value


def count_short_cycles(H):
    """(variables on a 4-cycle, variables on a cycle of length <= 6) as the reference's rooted finders see them."""
    import scipy.sparse as sp
    H = sp.csr_matrix(H)
    H.sort_indices()
    row_ptr = np.ascontiguousarray(H.indptr, dtype=np.int32)
    col_idx = np.ascontiguousarray(H.indices, dtype=np.int32)
    n4, n6 = C.c_int64(0), C.c_int64(0)
    _check(_lib.load().ldpc_h_count_short_cycles(row_ptr.ctypes.data, col_idx.ctypes.data, H.shape[0], H.shape[1],
                                                 C.byref(n4), C.byref(n6)))
    return n4.value, n6.value


def save_mat(path, H):
    """MAT-v5 file with the sparse double `H_sparse`, as the reference's generator saves it (Hgen...m:228-229)."""
    import scipy.io as sio
    import scipy.sparse as sp
    sio.savemat(path, {"H_sparse": sp.csc_matrix(H, dtype=np.float64)}, do_compression=True)
