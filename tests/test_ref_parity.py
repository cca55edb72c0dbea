"""Parity against the REFERENCE'S OWN code, not a restatement of it.

oracle/_ref/libldpc_ref.so is the reference's device sources (ldpc_erasure_decoder.cl, _old.pro, _perf_tests.cl,
ldpc_erasure_encoder.cl, data_in + Random123 threefry.h) compiled unmodified by gcc behind oracle/ref_shim/.
tests/golden/ref_vectors.json holds digests of its outputs on seeded cases (tools/make_ref_golden.py).

CPU tests: the restatement (oracle/ldpc_oracle.c) == golden digests, and == the compiled reference directly on
more seeded frames when the library is present (it is built where /root/reference exists and travels prebuilt).
GPU tests: the CUDA path, through the C ABI, == golden digests and == the compiled reference.
"""
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("make_ref_golden", os.path.join(ROOT, "tools", "make_ref_golden.py"))
gold = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gold)

with open(os.path.join(ROOT, "tests", "golden", "ref_vectors.json")) as _f:
    CASES = json.load(_f)["cases"]

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (no reference tree, no prebuilt library)")


def _case_id(c):
    return f"c{c['code']}-S{c['S']}-P{c['P']}-i{c['num_iter']}"


# ------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("case", CASES, ids=_case_id)
def test_oracle_matches_reference_golden(case):
    """encoder bytes, erasure flags, decoder output bytes and flags of the restatement == the reference's."""
    code = orc.Code.builtin(case["code"])
    info = gold.case_inputs(case, code.k)
    cw = orc.encode(code, info)
    assert gold.digest(cw) == case["cw"]
    flags = orc.gen_erasures_iid(code.n, case["seed"], case["B"], P=case["P"])
    assert gold.digest(flags) == case["flags"] and int(flags.sum()) == case["erased"]
    rx = cw.copy()
    rx[flags == 1] = 0
    # the canonical kernel has no early stop (decoder.cl:49); the restatement's early stop is output-neutral: both forms
    for early in (False, True):
        d = orc.decode(code, rx, flags, max_iter=case["num_iter"], early_stop=early)
        assert gold.digest(d["out"]) == case["out"]
        assert gold.digest(d["erased"][:, :code.k]) == case["out_flags"]
        assert [int(x) for x in d["fail_sys"]] == case["fail_sys"]


def test_golden_file_is_current():
    assert len(CASES) == len(gold.CASES)
    for c, g in zip(CASES, gold.CASES):
        assert all(c[key] == g[key] for key in g)


@needs_ref
def test_reference_tables_and_types():
    assert ref.code_params(0) == [2000, 1000, 0, 999, 250, 125] and ref.code_params(1) == [2040, 1530, 1000, 1509, 255, 192]
    assert ref.sizeof_symbol_type(top=True) == 1032 == ref.sizeof_symbol_type(top=False)   # main.cpp:42-47
    for ci in (0, 1):
        code = orc.Code.builtin(ci)
        rows = ref.vlist_rows(ci)
        assert all(list(code.col_idx[code.row_ptr[r]:code.row_ptr[r + 1]]) == rows[r] for r in range(code.m))


@needs_ref
@pytest.mark.parametrize("ci,P,seed", [(0, 19, 12345), (1, 13, 12345), (1, 9, -7), (0, 24, 999999), (1, 0, 1), (1, 64, 1)])
def test_reference_data_in_matches_oracle_generator(ci, P, seed):
    n = ref.code_params(ci)[0]
    assert np.array_equal(ref.data_in(ci, seed, P, 40), orc.gen_erasures_iid(n, seed, 40, P=P))


@needs_ref
@pytest.mark.parametrize("ci", [0, 1])
@pytest.mark.parametrize("S", [16, 64, 1024])
def test_reference_codec_matches_oracle(ci, S):
    code = orc.Code.builtin(ci)
    B = 10 if S == 1024 else 40
    info = gold.mix_bytes(B * code.k * S, 31 + ci + S).reshape(B, code.k, S)
    cw = ref.encode(ci, info, nthreads=2)
    assert np.array_equal(cw, orc.encode(code, info))
    if S == 1024:
        assert np.array_equal(cw, ref.encode(ci, info, top=True))
    for P in (9, 13, 19, 24):
        flags = ref.data_in(ci, 500 + P, P, B)
        rx = cw.copy()
        rx[flags == 1] = 0
        junk = gold.mix_bytes(rx.size, 77).reshape(rx.shape)      # NOT a codeword: only the exact serial schedule matches
        junk[flags == 1] = 0
        for it in (1, 2, 50):
            for inp in (rx, junk):
                r = ref.decode(ci, inp, flags, num_iter=it, variant="canon", top=(S == 1024), nthreads=3)
                o = orc.decode(code, inp, flags, max_iter=it)
                assert np.array_equal(r["out"], o["out"]) and np.array_equal(r["fail_sys"], o["fail_sys"])
                assert np.array_equal(r["out_flags"], o["erased"][:, :code.k])
            # the early-stop variant (_old.pro) reports through ERROR_STAT: same frame failures, RS blocks of (250,125)
            ro = ref.decode(ci, rx, flags, num_iter=it, variant="old", top=(S == 1024), nthreads=2)
            assert np.array_equal(ro["fail_sys"], o["fail_sys"])
            rs = np.array([orc.rs_mds_count(code.n - code.n % 250, 250, 125, f[:code.n - code.n % 250]) for f in flags])
            assert np.array_equal(ro["rs_errors"], rs)


@needs_ref
def test_reference_perf_variant_counts_rs_blocks_like_the_oracle():
    """_perf_tests.cl (the variant whose args the host sets): its RS-equivalent MDS count uses ldpc_params cols 4-5
    (:48-50,70-80) and must equal the oracle's; its frame-error count is the known-buggy one (SURVEY a-9) and only
    bounds the true count from above where peeling succeeds."""
    for ci, P in ((0, 24), (1, 12)):
        code = orc.Code.builtin(ci)
        B = 16
        flags = ref.data_in(ci, 4242, P, B)
        zero = np.zeros((B, code.n, 1024), dtype=np.uint8)
        rp = ref.decode(ci, zero, flags, num_iter=50, variant="perf", top=True, nthreads=4)
        prm = ref.code_params(ci)
        rs = np.array([orc.rs_mds_count(code.n, prm[4], prm[5], f) for f in flags])
        assert np.array_equal(rp["rs_errors"], rs)
        o = orc.decode(code, zero, flags, max_iter=50)
        assert (rp["fail_sys"] >= o["fail_sys"]).all()


# ------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def codecs():
    from ldpc_erasure_codes_b200.codec import LdpcCodec
    cache = {}

    def get(ci, S):
        if (ci, S) not in cache:
            cache[(ci, S)] = LdpcCodec(code=ci, symbol_bytes=S, device=0, max_batch=256)
        return cache[(ci, S)]
    yield get
    for c in cache.values():
        c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=_case_id)
def test_cuda_matches_reference_golden(codecs, case):
    import torch
    from ldpc_erasure_codes_b200.codec import unpack_mask
    codec = codecs(case["code"], case["S"])
    info = torch.from_numpy(gold.case_inputs(case, codec.k)).cuda()
    cw = codec.encode(info)
    assert gold.digest(cw.cpu().numpy()) == case["cw"]
    rx = cw.clone()
    mask = codec.gen_erasures(case["B"], case["seed"], P=case["P"], payload=rx)
    assert gold.digest(unpack_mask(mask, codec.n)) == case["flags"]
    out, fail = codec.decode(rx, mask, max_iter=case["num_iter"])
    assert gold.digest(out.cpu().numpy()) == case["out"]
    assert [int(x) for x in fail.cpu().numpy()] == case["fail_sys"]


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("ci", [0, 1])
@pytest.mark.parametrize("P", [9, 13, 19, 24])
def test_cuda_matches_compiled_reference(codecs, ci, P):
    """S = 1024 (the reference's SYM_LEN), both OpenCL codes: _ref == ldpc_oracle.c == CUDA on seeded frames."""
    import torch
    from ldpc_erasure_codes_b200.codec import unpack_mask
    S, B = 1024, 24
    codec = codecs(ci, S)
    code = orc.Code.builtin(ci)
    info_h = gold.mix_bytes(B * codec.k * S, 900 + 10 * ci + P).reshape(B, codec.k, S)
    cw = codec.encode(torch.from_numpy(info_h).cuda())
    cw_ref = ref.encode(ci, info_h, top=True, nthreads=4)
    assert np.array_equal(cw.cpu().numpy(), cw_ref)
    rx = cw.clone()
    mask = codec.gen_erasures(B, 31337, P=P, payload=rx)
    flags = ref.data_in(ci, 31337, P, B)
    assert np.array_equal(unpack_mask(mask, codec.n), flags)
    for it in (2, 50):
        out, fail = codec.decode(rx, mask, max_iter=it)
        r = ref.decode(ci, rx.cpu().numpy(), flags, num_iter=it, variant="canon", top=True, nthreads=4)
        o = orc.decode(code, rx.cpu().numpy(), flags, max_iter=it)
        assert np.array_equal(out.cpu().numpy(), r["out"]) and np.array_equal(fail.cpu().numpy(), r["fail_sys"])
        assert np.array_equal(o["out"], r["out"]) and np.array_equal(o["fail_sys"], r["fail_sys"])
