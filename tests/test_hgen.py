"""the girth-8 triangular-form H generator and the short-cycle checker (csrc/hgen.cpp) against
the plain-Python restatement of the reference's cycle finders (oracle/hgen_oracle.py), the reference's committed codes,
and the properties Matlab/Hgen_irregularDegree_no6cycles_systematic_encoding.m promises.  Host-only entry points: these
tests need no GPU (the generated code is pushed through encode / decode on the GPU in test_parity_gpu.py)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from ldpc_erasure_codes_b200 import hgen
from oracle import hgen_oracle as ho

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODES = os.path.join(ROOT, "ldpc_erasure_codes_b200", "codes")


def _lists(H):
    H = sp.csr_matrix(H)
    H.sort_indices()
    return ho.lists_from_csr(H.indptr, H.indices, H.shape[0], H.shape[1])


def test_cycle_finders_on_planted_cycles():
    # a 4-cycle: variables 0 and 1 share checks 0 and 1
    H4 = np.array([[1, 1, 0, 0], [1, 1, 1, 0], [0, 0, 1, 1]], dtype=np.uint8)
    # (variable 2 is not on it, but the length-6 finder rooted there walks c1 -> v0, v1 -> c0 twice and flags it: the
    #  finders are only meant to be run on graphs that had no short cycle before the last edge)
    assert hgen.count_short_cycles(H4) == (2, 3)
    assert ho.count_short_cycles(*_lists(H4)) == (2, 3)
    # a 6-cycle v0-c0-v1-c1-v2-c2-v0 and a pendant variable
    H6 = np.array([[1, 1, 0, 0], [0, 1, 1, 0], [1, 0, 1, 1]], dtype=np.uint8)
    assert hgen.count_short_cycles(H6) == (0, 3)
    assert ho.count_short_cycles(*_lists(H6)) == (0, 3)
    # a tree
    Ht = np.array([[1, 1, 0, 0], [0, 1, 1, 0], [0, 0, 1, 1]], dtype=np.uint8)
    assert hgen.count_short_cycles(Ht) == (0, 0)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_cycle_finders_match_restatement_on_random_graphs(seed):
    rng = np.random.default_rng(seed)
    m, n = 40, 90
    H = (rng.random((m, n)) < 0.06).astype(np.uint8)
    H[rng.integers(m, size=n), np.arange(n)] = 1          # no empty column
    assert hgen.count_short_cycles(H) == ho.count_short_cycles(*_lists(H))


@pytest.mark.parametrize("name,expect", [("n2000_k1000", (0, 0)), ("n4000_k2000", (0, 5)), ("n2040_k1530", (0, 41))])
def test_committed_codes(name, expect):
    """The reference's own matrices through its own finders: no 4-cycles anywhere; the (2000,1000) code has girth 8, the
    other two carry a few variables on 6-cycles (the script adds its last-row and staircase edges unchecked, :214-222)."""
    import scipy.io as sio
    H = sio.loadmat(os.path.join(CODES, name + ".mat"))["H_sparse"]
    got = hgen.count_short_cycles(H)
    assert got == expect
    if name == "n2000_k1000":                             # (the restatement in pure Python: one code is enough)
        assert ho.count_short_cycles(*_lists(H)) == expect


@pytest.mark.parametrize("c_prof,v_prof", [([(200, 4)], [(400, 2)]), ([(1000, 6)], [(2000, 3)])])
def test_generator_properties(c_prof, v_prof):
    H, tries = hgen.generate(c_prof, v_prof, seed=7, max_tries=50)
    m, n = H.shape
    k = n - m
    assert (m, n) == (sum(c for c, _ in c_prof), sum(c for c, _ in v_prof)) and tries >= 1
    Hd = sp.csr_matrix(H)
    Hd.sort_indices()
    # lower triangular right part with a unit diagonal: the last entry of row r is column k + r (Hgen...m:189-194)
    for r in range(m):
        assert Hd.indices[Hd.indptr[r + 1] - 1] == k + r
    T = Hd[:, k:].toarray()
    assert np.all(np.triu(T, 1) == 0) and np.all(np.diag(T) == 1)
    # degree profile: rows as asked except the last (diagonal + staircase only, :214-222) and the rows that received a
    # staircase edge; columns within one edge of the profile (:114-116)
    rw = np.diff(Hd.indptr)
    want = c_prof[0][1]
    assert np.all((rw[:-1] == want) | (rw[:-1] == want + 1)) and rw[-1] <= 2
    cw = np.asarray(Hd.sum(axis=0)).ravel()
    assert cw.max() <= v_prof[0][1] + 1 and cw.min() >= 1
    # girth >= 8, by the library and by the restatement
    assert hgen.count_short_cycles(H) == (0, 0)
    if n <= 400:
        assert ho.count_short_cycles(*_lists(H)) == (0, 0)
    # same seed, same matrix; another seed, another matrix
    H2, _ = hgen.generate(c_prof, v_prof, seed=7, max_tries=50)
    assert (H != H2).nnz == 0
    H3, _ = hgen.generate(c_prof, v_prof, seed=8, max_tries=50)
    assert (H != H3).nnz > 0


def test_generator_rejects_bad_profiles():
    from ldpc_erasure_codes_b200._lib import LdpcCudaError
    with pytest.raises(LdpcCudaError):
        hgen.generate([(10, 4)], [(20, 3)])               # 40 check edges vs 60 variable edges (Hgen...m:64-66)
    with pytest.raises(LdpcCudaError):
        hgen.generate([(30, 6)], [(60, 3)], max_tries=3)  # too small for girth 8 at these degrees: gives up, no hang


def test_generated_code_loads(tmp_path):
    """generator -> MAT-v5 file -> the library's own loader: triangular, same CSR."""
    import ctypes as C
    from ldpc_erasure_codes_b200 import _lib
    H, _ = hgen.generate([(200, 4)], [(400, 2)], seed=11)
    path = str(tmp_path / "gen.mat")
    hgen.save_mat(path, H)
    dims = (C.c_int32 * 4)()
    lib = _lib.load()
    assert lib.ldpc_read_h_file(path.encode(), C.byref(dims), None, None) == 0
    assert (dims[0], dims[1], dims[2], dims[3]) == (200, 400, H.nnz, 1)
