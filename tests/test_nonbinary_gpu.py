"""SURVEY 8(f) rank 3 -- the non-binary GF(256) LDPC code (csrc/nb_ldpc.cuh, ldpc_nb_*) against the oracle's restatement of
Matlab/ErasureCodes_NonBinaryLDPCSim.m:176-182 (encoder) and Matlab/My_LDPC_HybridML_NonBinary_Erasure_Decoder.m (sweeps +
Gauss-Jordan over GF(256)).  Bit-exact: payloads, failure flags, elimination counters."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


def _setup(ci, S, seed, max_batch=256):
    from ldpc_erasure_codes_b200.codec import LdpcCodec, NbLdpcCodec
    base = LdpcCodec(code=ci, symbol_bytes=S, device=0, max_batch=max_batch)
    nb = NbLdpcCodec(base, seed=seed)
    code = orc.Code.builtin(ci)
    coef = orc.nb_coefficients(code, seed)
    assert np.array_equal(nb.coefficients(), coef)          # the same draw on both sides
    return base, nb, code, coef


@pytest.mark.parametrize("ci,S", [(1, 64), (0, 32), (2, 16), (1, 128)])
def test_nb_encode_bit_exact(ci, S):
    base, nb, code, coef = _setup(ci, S, seed=11 + ci)
    B = 37
    info = torch.from_numpy(np.random.default_rng(ci).integers(0, 256, (B, code.k, S), dtype=np.uint8)).cuda()
    cw = nb.encode(info)
    ref = orc.nb_encode(code, coef, _np(info))
    assert np.array_equal(_np(cw), ref)
    # the codeword satisfies every check of H_nb (independent of the encoder's order of operations)
    mul = orc.gf256_tables()["mul"]
    for r in range(0, code.m, 37):
        acc = np.zeros(S, np.uint8)
        for j in range(code.row_ptr[r], code.row_ptr[r + 1]):
            acc ^= mul[coef[j]][ref[0, code.col_idx[j]]]
        assert not acc.any()
    nb.close(); base.close()


@pytest.mark.parametrize("ci,S,P,mode,it", [(1, 64, 10, "peel", 50), (1, 64, 13, "peel", 10), (1, 64, 13, "hybrid", 10),
                                            (1, 64, 14, "hybrid", 10), (0, 32, 26, "hybrid", 10), (0, 32, 24, "peel", 3),
                                            (2, 16, 27, "hybrid", 10), (1, 128, 12, "hybrid", 10), (1, 64, 17, "hybrid", 10)])
def test_nb_decode_bit_exact(ci, S, P, mode, it):
    base, nb, code, coef = _setup(ci, S, seed=5, max_batch=64)      # B > max_batch: chunked
    B = 100
    info = np.random.default_rng(P).integers(0, 256, (B, code.k, S), dtype=np.uint8)
    cw = orc.nb_encode(code, coef, info)
    flags = orc.gen_erasures_iid(code.n, 40 + P, B, P=P)
    rx = cw.copy()
    rx[flags == 1] = 0
    d_rx = torch.from_numpy(rx).cuda()
    mask = base.gen_erasures(B, 40 + P, P=P)
    base.reset_stats()
    out, fail = nb.decode(d_rx, mask, max_iter=it, mode=mode)
    ref = orc.nb_decode(code, coef, rx, flags, max_iter=it, mode=mode)
    assert np.array_equal(_np(fail), ref["fail_sys"])
    assert np.array_equal(_np(out), ref["out"])
    good = ref["fail_sys"] == 0
    assert np.array_equal(ref["out"][good], info[good])
    st = base.stats()
    assert st["frames"] == B
    if mode == "hybrid":
        assert st["ml_attempts"] == int((ref["status"] > 0).sum())
        assert st["ml_failures"] == int((ref["status"] == 2).sum())
        if P in (13, 14, 26, 27):
            assert (ref["status"] == 1).sum() > 5              # the elimination really ran and succeeded
        if P == 17:
            assert (ref["status"] == 2).sum() > 5              # and met rank-deficient sets
    nb.close(); base.close()


def test_nb_user_coefficients_and_generated_code(tmp_path):
    """Coefficients supplied by the caller, on a code from the library's own generator."""
    from ldpc_erasure_codes_b200 import hgen
    from ldpc_erasure_codes_b200.codec import LdpcCodec, NbLdpcCodec
    H, _ = hgen.generate([(300, 4)], [(600, 2)], seed=9)
    path = str(tmp_path / "g.mat")
    hgen.save_mat(path, H)
    S, B = 48, 50
    base = LdpcCodec(code=path, symbol_bytes=S, device=0, max_batch=64)
    code = orc.Code(H)
    coef = np.random.default_rng(1).integers(1, 256, len(code.col_idx), dtype=np.uint8)
    nb = NbLdpcCodec(base, coef=coef)
    info = np.random.default_rng(2).integers(0, 256, (B, code.k, S), dtype=np.uint8)
    cw = nb.encode(torch.from_numpy(info).cuda())
    ref_cw = orc.nb_encode(code, coef, info)
    assert np.array_equal(_np(cw), ref_cw)
    flags = orc.gen_erasures_iid(code.n, 3, B, P=24)
    rx = ref_cw.copy()
    rx[flags == 1] = 0
    out, fail = nb.decode(torch.from_numpy(rx).cuda(), base.gen_erasures(B, 3, P=24), max_iter=10, mode="hybrid")
    ref = orc.nb_decode(code, coef, rx, flags, max_iter=10, mode="hybrid")
    assert np.array_equal(_np(fail), ref["fail_sys"]) and np.array_equal(_np(out), ref["out"])
    with pytest.raises(Exception):
        bad = coef.copy(); bad[5] = 0
        NbLdpcCodec(base, coef=bad)
    nb.close(); base.close()
