"""Pins the CPU oracle (oracle/ldpc_oracle.c) against everything the reference commits for this path:
Random123 KATs (Threefry), the GF(256) .mat tables, the three H matrices, plus the survey-time
cross-check values and algebraic round-trip properties."""
import json
import os
import zlib

import numpy as np
import pytest

from oracle import oracle as orc

REF_MATLAB = "/root/reference/Matlab"


def info_pattern(k, S=8):
    """info symbol i = uint64((i+1) * 0x9E3779B97F4A7C15 mod 2^64), little-endian, S = 8 B."""
    v = (np.arange(1, k + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
    b = v.view(np.uint8).reshape(k, 8)
    if S == 8:
        return b.copy()
    return np.tile(b, (1, S // 8)).copy()


# ----------------------------------------------------------------------------- Threefry
def test_threefry_kats():
    # Random123 kat_vectors, threefry4x32 20 rounds
    assert [hex(x) for x in orc.threefry4x32_20([0, 0, 0, 0], [0, 0, 0, 0])] == \
        ["0x9c6ca96a", "0xe17eae66", "0xfc10ecd4", "0x5256a7d8"]
    f = 0xFFFFFFFF
    assert [hex(x) for x in orc.threefry4x32_20([f] * 4, [f] * 4)] == \
        ["0x2a881696", "0x57012287", "0xf6c7446e", "0xa16a6732"]
    assert [hex(x) for x in orc.threefry4x32_20([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                                                [0xa4093822, 0x299f31d0, 0x082efa98, 0xec4e6c89])] == \
        ["0x59cd1dbb", "0xb8879579", "0x86b5d00c", "0xac8b6d84"]


def test_generator_survey_values():
    # seed 12345: first ten (x0 & 63) for counters 1..10
    vals = [int(orc.threefry4x32_20([i, 0, 0, 0], [1, 12345, 0, 0])[0] & 63) for i in range(1, 11)]
    assert vals == [36, 9, 17, 33, 28, 46, 16, 34, 50, 43]
    f = orc.gen_erasures_iid(2000, 12345, 1, P=19)[0]
    assert f.sum() == 624 and list(np.nonzero(f)[0][:8]) == [1, 2, 6, 11, 12, 15, 16, 20]
    f = orc.gen_erasures_iid(2040, 12345, 1, P=13)[0]
    assert f.sum() == 447 and list(np.nonzero(f)[0][:8]) == [1, 11, 15, 16, 20, 24, 25, 28]
    f = orc.gen_erasures_iid(4000, 12345, 1, P=19)[0]
    assert f.sum() == 1178


def test_generator_counter_continuity():
    # counter never resets between frames (decoder_top.cl:75,96): frame f symbol s uses 1 + f*n + s
    a = orc.gen_erasures_iid(2040, 7, 5, P=13)
    b = orc.gen_erasures_iid(2040, 7, 2, P=13, frame0=3)
    assert np.array_equal(a[3:], b)
    c = orc.gen_erasures_iid(2040, 7, 5, p32=int(0.2 * 2 ** 32))
    assert abs(c.mean() - 0.2) < 0.02


# ----------------------------------------------------------------------------- GF(256)
def test_gf256_tables_match_reference_mat(golden_dir):
    g = np.load(os.path.join(golden_dir, "gf256_tables.npz"))
    t = orc.gf256_tables()
    assert np.array_equal(t["mul"], g["mul"])
    assert np.array_equal(t["inv"], g["inv"])
    a = np.arange(256, dtype=np.uint8)
    assert np.array_equal(g["add"], a[:, None] ^ a[None, :])
    assert t["mul"][2, 0x80] == 0x71
    assert list(t["inv"][:8]) == [1, 184, 208, 92, 159, 104, 134, 46]


@pytest.mark.skipif(not os.path.isdir(REF_MATLAB), reason="reference not mounted")
def test_gf256_fixture_is_the_reference_file(golden_dir):
    import scipy.io as sio
    g = np.load(os.path.join(golden_dir, "gf256_tables.npz"))
    r = sio.loadmat(os.path.join(REF_MATLAB, "GF_256_add_mult_inv_tables.mat"))
    assert np.array_equal(g["mul"], r["GF_mult_lookup"])
    assert np.array_equal(g["inv"], r["GF_inv_lookup"].reshape(-1))
    assert np.array_equal(g["add"], r["GF_add_lookup"])


# ----------------------------------------------------------------------------- H
@pytest.mark.parametrize("ci", [0, 1, 2])
def test_codes_digest_and_triangular_form(ci, golden_dir):
    code = orc.Code.builtin(ci)
    d = json.load(open(os.path.join(golden_dir, "codes_digest.json")))[orc.CODE_TABLE[ci]["name"]]
    assert (code.n, code.k, code.m) == (d["n"], d["k"], d["m"])
    assert zlib.crc32(code.row_ptr.astype("<i4").tobytes()) == d["crc32_row_ptr"]
    assert zlib.crc32(code.col_idx.astype("<i4").tobytes()) == d["crc32_col_idx"]
    # last entry of row r is the diagonal k + r (SURVEY a-3)
    last = code.col_idx[code.row_ptr[1:] - 1]
    assert np.array_equal(last, code.k + np.arange(code.m))


@pytest.mark.skipif(not os.path.isdir(REF_MATLAB), reason="reference not mounted")
@pytest.mark.parametrize("ci", [0, 1, 2])
def test_exported_codes_equal_reference_mat(ci):
    ref_name = {0: "n2000_k1000_no6cycles_triangleForm_OpenCL_H.mat",
                1: "n2040_k1530_irreg_H_no6cycles_triangleForm.mat",
                2: "n4000_k2000_no6cycles_triangleForm.mat"}[ci]
    a = orc.Code.builtin(ci)
    b = orc.Code.from_mat(os.path.join(REF_MATLAB, ref_name))
    assert np.array_equal(a.row_ptr, b.row_ptr) and np.array_equal(a.col_idx, b.col_idx)


@pytest.mark.skipif(not os.path.isdir("/root/reference/OpenCL/device"), reason="reference not mounted")
def test_codes_equal_reference_vlist_header():
    """LDPC_Vlist_data.h rows 0-999 / 1000-1509 are the (2000,1000) / (2040,1530) codes."""
    import re
    txt = open("/root/reference/OpenCL/device/LDPC_Vlist_data.h").read()
    body = txt[txt.index("parity_check_mat_Vlist_master"):]
    rows = re.findall(r"\{([0-9,\s]+)\}", body)
    rows = [[int(x) for x in r.split(",")] for r in rows]
    assert len(rows) == 1510 and all(len(r) == 20 for r in rows)
    for ci, (lo, hi) in {0: (0, 1000), 1: (1000, 1510)}.items():
        code = orc.Code.builtin(ci)
        for r in range(lo, hi):
            w = rows[r][0]
            mine = code.col_idx[code.row_ptr[r - lo]:code.row_ptr[r - lo + 1]] + 1
            assert w == len(mine) and rows[r][1:1 + w] == list(mine) and not any(rows[r][1 + w:])


# ----------------------------------------------------------------------------- encoder
ENC_XCHECK = {  # parity[0], parity[m-1], XOR of all parities, crc32(codeword)
    0: ("a67b72cd5f1686bf", "918b44765d60493b", "51fb1978ee4133bc", 0x8673d3b5),
    1: ("8b900b08ad986692", "ff5dfa5139cb3011", "34cc18a03e98ff35", 0x91491f2f),
    2: ("38afafaae57988c4", "056a305906d0edd4", "b245738405d81d68", 0xa70ddfec),
}


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_encoder_survey_crosscheck(ci):
    code = orc.Code.builtin(ci)
    cw = orc.encode(code, info_pattern(code.k)[None])[0]
    par = cw[code.k:].view("<u8").reshape(-1)
    p0, pl, px, crc = ENC_XCHECK[ci]
    assert f"{int(par[0]):016x}" == p0
    assert f"{int(par[-1]):016x}" == pl
    assert f"{int(np.bitwise_xor.reduce(par)):016x}" == px
    assert zlib.crc32(cw.tobytes()) == crc


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_encoder_output_satisfies_all_checks(ci):
    code = orc.Code.builtin(ci)
    rng = np.random.default_rng(ci)
    cw = orc.encode(code, rng.integers(0, 256, (3, code.k, 16), dtype=np.uint8))
    for r in range(code.m):
        cols = code.col_idx[code.row_ptr[r]:code.row_ptr[r + 1]]
        assert not np.bitwise_xor.reduce(cw[:, cols, :], axis=1).any()


# ----------------------------------------------------------------------------- peeling
def _frame(code, seed, P, frame, S=8):
    cw = orc.encode(code, info_pattern(code.k, S)[None])[0]
    er = orc.gen_erasures_iid(code.n, seed, 1, P=P, frame0=frame)[0]
    rx = cw.copy()
    rx[er == 1] = 0
    return cw, rx, er


def test_peel_survey_frames():
    c0 = orc.Code.builtin(0)
    cw, rx, er = _frame(c0, 12345, 19, 0)
    p, e, it = orc.peel_single(c0, rx, er, max_iter=50)
    assert e.sum() == 0 and np.array_equal(p, cw) and it == 4
    c1 = orc.Code.builtin(1)
    cw, rx, er = _frame(c1, 12345, 13, 0)
    p, e, it = orc.peel_single(c1, rx, er, max_iter=50)
    assert e.sum() == 212  # stalls: a GE test case
    p10, e10, _ = orc.peel_single(c1, rx, er, max_iter=10)
    assert e10.sum() == 232
    c2 = orc.Code.builtin(2)
    cw, rx, er = _frame(c2, 12345, 19, 0)
    p, e, it = orc.peel_single(c2, rx, er, max_iter=50)
    assert e.sum() == 0 and np.array_equal(p, cw) and it == 3


def test_peel_early_stop_is_output_neutral_and_u64_equal():
    code = orc.Code.builtin(1)
    rng = np.random.default_rng(5)
    for P in (9, 12, 13):
        cw = orc.encode(code, rng.integers(0, 256, (1, code.k, 16), dtype=np.uint8))[0]
        er = orc.gen_erasures_iid(code.n, 99, 1, P=P, frame0=P)[0]
        rx = cw.copy(); rx[er == 1] = 0
        a = orc.peel_single(code, rx, er, 50, early_stop=False)
        b = orc.peel_single(code, rx, er, 50, early_stop=True)
        c = orc.peel_single(code, rx, er, 50, early_stop=True, u64=True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert np.array_equal(b[0], c[0]) and np.array_equal(b[1], c[1]) and b[2] == c[2]
        known = a[1] == 0
        assert np.array_equal(a[0][known], cw[known])


def test_peel_edge_cases():
    code = orc.Code.builtin(1)
    cw = orc.encode(code, info_pattern(code.k)[None])[0]
    # no erasures
    p, e, it = orc.peel_single(code, cw, np.zeros(code.n, np.uint8), 50)
    assert it == 0 and np.array_equal(p, cw)
    # everything erased: nothing can be recovered
    p, e, it = orc.peel_single(code, np.zeros_like(cw), np.ones(code.n, np.uint8), 50)
    assert e.sum() == code.n and not p.any()
    # only parity erased: one sweep recovers all (each row's diagonal is its last member)
    er = np.zeros(code.n, np.uint8); er[code.k:] = 1
    rx = cw.copy(); rx[code.k:] = 0
    p, e, it = orc.peel_single(code, rx, er, 50)
    assert e.sum() == 0 and np.array_equal(p, cw)


# ----------------------------------------------------------------------------- hybrid
def test_hybrid_survey_frames():
    code = orc.Code.builtin(1)
    expect = {0: (447, 1), 1: (387, 0), 2: (432, 1)}
    for f, (n_er, status) in expect.items():
        cw, rx, er = _frame(code, 12345, 13, f)
        assert er.sum() == n_er
        p, e, st, rowops = orc.hybrid_single(code, rx, er, peel_iter=10)
        assert st == status and e.sum() == 0 and np.array_equal(p, cw)
        # GE from the peeling fixed point gives the same (unique) answer
        p2, e2, st2, _ = orc.hybrid_single(code, rx, er, peel_iter=1000)
        assert np.array_equal(p2, cw)


def test_hybrid_rank_deficient_contract():
    code = orc.Code.builtin(1)
    cw = orc.encode(code, info_pattern(code.k)[None])[0]
    er = orc.gen_erasures_iid(code.n, 4, 1, P=20)[0]   # 31 %: hopeless
    rx = cw.copy(); rx[er == 1] = 0
    p, e, st, _ = orc.hybrid_single(code, rx, er, peel_iter=10)
    pp, ee, _ = orc.peel_single(code, rx, er, 10)
    assert st == 2 and np.array_equal(p, pp) and np.array_equal(e, ee)


# ----------------------------------------------------------------------------- bursty
def test_bursty_state_chain_and_mean():
    fl, st = orc.gen_erasures_bursty(4000, 3, 50, 0.001, 0.1, 10.0)
    a, s1 = orc.gen_erasures_bursty(4000, 3, 20, 0.001, 0.1, 10.0)
    b, s2 = orc.gen_erasures_bursty(4000, 3, 30, 0.001, 0.1, 10.0, frame0=20, state=s1)
    assert np.array_equal(fl, np.concatenate([a, b])) and s2 == st
    big, _ = orc.gen_erasures_bursty(4000, 11, 500, 0.02, 0.4, 10.0)
    mean = (1 / (1 + 1 / 10.0)) * 0.02 + (1 - 1 / (1 + 1 / 10.0)) * 0.4   # Bursty_Error_Channel_Model.m:69
    assert abs(big.mean() - mean) < 0.01


# ----------------------------------------------------------------------------- RS
RS_XCHECK = {
    (255, 191): ("0378cbca5629", "28964632c889", 0x8cd07b24, "9f20d0b6f76b", 0xe49085a0),
    (255, 192): ("8ee9897489f8", "521eee1e600e", 0xe751cdd2, "5dfc8bdb61b7", 0x7ddaba94),
}


@pytest.mark.parametrize("nk", [(255, 191), (255, 192)])
def test_rs_generator_survey_crosscheck(nk):
    n, k = nk
    G = orc.rs_gsys(n, k)
    assert np.array_equal(G[:, :k], np.eye(k, dtype=np.uint8))
    P = G[:, k:]
    p0, pl, crcP, par, crcC = RS_XCHECK[nk]
    assert P[0, :6].tobytes().hex() == p0 and P[-1, :6].tobytes().hex() == pl
    assert zlib.crc32(P.tobytes()) == crcP
    u = ((37 * np.arange(k) + 11) & 255).astype(np.uint8)[:, None]
    cw = orc.rs_encode(G, u)[:, 0]
    assert cw[k:k + 6].tobytes().hex() == par and zlib.crc32(cw.tobytes()) == crcC


def test_rs_small_and_250():
    G = orc.rs_gsys(7, 5)
    assert G[0, 5:].tobytes().hex() == "c188" and G[4, 5:].tobytes().hex() == "3e1b"
    G = orc.rs_gsys(250, 125)
    assert zlib.crc32(G[:, 125:].tobytes()) == 0x7ff4fca0


@pytest.mark.parametrize("nk", [(7, 5), (255, 191), (255, 192)])
def test_rs_roundtrip(nk):
    """Mirror of Matlab/Test_My_RS_Decode.m:45-58: random k-subsets of received symbols decode to the source."""
    n, k = nk
    G = orc.rs_gsys(n, k)
    rng = np.random.default_rng(n * 1000 + k)
    for trial in range(200 if n == 7 else 6):
        u = rng.integers(0, 256, (k, 4), dtype=np.uint8)
        cw = orc.rs_encode(G, u)
        idx = np.sort(rng.permutation(n)[:k]).astype(np.int32)
        out, rd = orc.rs_decode(G, idx, cw[idx])
        assert rd == 0 and np.array_equal(out, u)


def test_rs_first_k_received_rule():
    n, k = 255, 191
    G = orc.rs_gsys(n, k)
    rng = np.random.default_rng(1)
    u = rng.integers(0, 256, (k, 8), dtype=np.uint8)
    cw = orc.rs_encode(G, u)
    er = rng.random(n) < 0.2
    rec = np.nonzero(~er)[0]
    if len(rec) >= k:
        out, rd = orc.rs_decode(G, rec[:k].astype(np.int32), cw[rec[:k]])
        assert rd == 0 and np.array_equal(out, u)


def test_mds_count():
    er = np.zeros(2040, np.uint8)
    er[:64] = 1          # block 0: 64 > 63 -> fail
    er[255:255 + 63] = 1  # block 1: 63 -> ok
    assert orc.rs_mds_count(2040, 255, 192, er) == 1


# ---------------------------------------------------------------------------- FEC packet front-ends
def test_fec_packet_format_and_reassembly():
    """Header word = [class:8|block:8|symbol:16] in both halves (encoder_VITA_in_UDP_out.cl:100-104,170-175); the receiver
    places a symbol by its header, clears its flag, drops foreign blocks (decoder_with_reordering_logic.cl:77-131)."""
    rng = np.random.default_rng(7)
    B, n, S = 4, 37, 16
    cw = rng.integers(0, 256, (B, n, S), dtype=np.uint8)
    pk = orc.packetize(cw, block0=254)                       # blocks 254, 255, 0, 1: the counter is modulo 256
    assert pk.shape == (B * n, 8 + S)
    hdr = pk[:, :8].copy().view("<u8").ravel()
    assert hdr[0] == 0x01FE000001FE0000 and hdr[n + 5] == 0x01FF000501FF0005 and hdr[2 * n + 36] == 0x0100002401000024
    assert np.array_equal(pk[:, 8:].reshape(B, n, S), cw)
    # arrival order does not matter; lost packets stay erased and zero; duplicates count twice; foreign packets are dropped
    order = rng.permutation(B * n)
    lost = rng.random(B * n) < 0.2
    stream = np.concatenate([pk[order][~lost[order]], pk[:3], orc.packetize(cw[:1], block0=77)[:5]])
    bad = pk[10:11].copy(); bad[0, 3] ^= 0xFF                # class byte of the low half corrupted: halves differ
    stream = np.concatenate([stream, bad])
    got, flags, counts = orc.depacketize(stream, n, 254, B)
    want = cw.copy(); want.reshape(B * n, S)[lost & (np.arange(B * n) >= 3)] = 0
    assert np.array_equal(got, want)
    assert np.array_equal(flags.ravel(), (lost & (np.arange(B * n) >= 3)).astype(np.uint8))
    assert counts[B] == 6 and counts[:B].sum() == (~lost).sum() + 3


def test_receiver_hand_off_rule():
    """decoder_with_reordering_logic.cl:54-55,139 for the reference's (2000,1000): desired 800, minimum 200."""
    r = lambda cur, nxt: orc.ready_to_decode(2000, 1000, cur, nxt)
    assert r(2000, 0) and not r(1999, 0)
    assert r(1801, 11) and not r(1800, 11) and not r(1801, 10)
    assert r(1201, 101) and not r(1200, 101) and not r(1201, 100)
