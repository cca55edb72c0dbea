"""ctypes/numpy front-end of the CPU oracle (oracle/ldpc_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(ldpc_erasure_codes_b200) never imports this module.

H matrices are read here with scipy (independently of the library's own C++
MAT-v5 loader, so the two cross-check each other).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libldpc_oracle.so")
_REPO = os.path.dirname(_HERE)
CODES_DIR = os.path.join(_REPO, "ldpc_erasure_codes_b200", "codes")

# code table: reference ldpc_params rows (OpenCL/device/LDPC_Vlist_data.h:10-14) + the .mat-only code
CODE_TABLE = {
    0: dict(name="n2000_k1000", n=2000, k=1000, rs_n=250, rs_k=125),
    1: dict(name="n2040_k1530", n=2040, k=1530, rs_n=255, rs_k=192),
    2: dict(name="n4000_k2000", n=4000, k=2000, rs_n=250, rs_k=125),
}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ldpc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_ldpc_peel.restype = C.c_int
        _lib.orc_ldpc_peel_u64.restype = C.c_int
        _lib.orc_ldpc_hybrid.restype = C.c_int
        _lib.orc_rs_mds_count.restype = C.c_int
        _lib.orc_rs_gsys.restype = C.c_int
        _lib.orc_rs_decode.restype = C.c_int
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Code:
    """H in CSR form (0-based, ascending) = the reference's Vlist rows."""

    def __init__(self, H, k=None, name=""):
        import scipy.sparse as sp

        R = sp.csr_matrix(H)
        R.sort_indices()
        self.m, self.n = R.shape
        self.k = self.n - self.m if k is None else k
        self.row_ptr = np.ascontiguousarray(R.indptr, dtype=np.int32)
        self.col_idx = np.ascontiguousarray(R.indices, dtype=np.int32)
        self.name = name

    @staticmethod
    def from_mat(path, name=""):
        import scipy.io as sio

        H = sio.loadmat(path, spmatrix=True)["H_sparse"]
        return Code(H, name=name or os.path.basename(path))

    @staticmethod
    def builtin(code_ind: int) -> "Code":
        t = CODE_TABLE[code_ind]
        return Code.from_mat(os.path.join(CODES_DIR, t["name"] + ".mat"), name=t["name"])


def threefry4x32_20(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().orc_threefry4x32_20(_p(c), _p(k), _p(o))
    return o


def gen_erasures_iid(n, seed, nframes, P=None, p32=None, frame0=0):
    """flags [nframes][n] u8.  P = numerator/64 (reference rule) or p32 = 32-bit threshold (extension)."""
    flags = np.zeros((nframes, n), dtype=np.uint8)
    mode, thr = (0, int(P)) if P is not None else (1, int(p32))
    lib().orc_gen_erasures_iid(C.c_int(n), C.c_uint32(seed & 0xFFFFFFFF), C.c_int(mode), C.c_uint32(thr),
                               C.c_uint64(frame0), C.c_int64(nframes), _p(flags))
    return flags


def gen_erasures_bursty(n, seed, nframes, alpha, beta, bias, frame0=0, state=0):
    flags = np.zeros((nframes, n), dtype=np.uint8)
    st = C.c_int(state)
    lib().orc_gen_erasures_bursty(C.c_int(n), C.c_uint32(seed & 0xFFFFFFFF), C.c_double(alpha), C.c_double(beta),
                                  C.c_double(bias), C.c_uint64(frame0), C.c_int64(nframes), C.byref(st), _p(flags))
    return flags, st.value


def encode(code: Code, info: np.ndarray, nthreads=0) -> np.ndarray:
    """info [B][k][S] u8 -> codewords [B][n][S] u8."""
    info = np.ascontiguousarray(info, dtype=np.uint8)
    B, k, S = info.shape
    assert k == code.k
    cw = np.zeros((B, code.n, S), dtype=np.uint8)
    lib().orc_ldpc_encode_batch(C.c_int(code.n), C.c_int(code.k), _p(code.row_ptr), _p(code.col_idx), C.c_int(S),
                                C.c_int64(B), _p(info), _p(cw), C.c_int(nthreads))
    return cw


def decode(code: Code, payload: np.ndarray, erased: np.ndarray, max_iter=50, early_stop=True, mode="peel",
           nthreads=0, inplace=False):
    """payload [B][n][S], erased [B][n] (u8 flags).  Returns dict(out, payload, erased, fail_sys, iters, status)."""
    payload = np.array(payload, dtype=np.uint8, order="C", copy=not inplace)
    erased = np.array(erased, dtype=np.uint8, order="C", copy=not inplace)
    B, n, S = payload.shape
    assert n == code.n and erased.shape == (B, n)
    out = np.zeros((B, code.k, S), dtype=np.uint8)
    fail = np.zeros(B, dtype=np.uint8)
    iters = np.zeros(B, dtype=np.int32)
    status = np.zeros(B, dtype=np.int32)
    lib().orc_ldpc_decode_batch(C.c_int(code.n), C.c_int(code.k), _p(code.row_ptr), _p(code.col_idx), C.c_int(S),
                                C.c_int64(B), _p(payload), _p(erased), _p(out), _p(fail), _p(iters), _p(status),
                                C.c_int(max_iter), C.c_int(1 if early_stop else 0),
                                C.c_int({"peel": 0, "hybrid": 1}[mode]), C.c_int(nthreads))
    return dict(out=out, payload=payload, erased=erased, fail_sys=fail, iters=iters, status=status)


def peel_single(code: Code, payload, erased, max_iter=50, early_stop=True, u64=False):
    payload = np.array(payload, dtype=np.uint8, order="C")
    erased = np.array(erased, dtype=np.uint8, order="C")
    n, S = payload.shape
    fn = lib().orc_ldpc_peel_u64 if u64 else lib().orc_ldpc_peel
    it = fn(C.c_int(code.n), C.c_int(code.k), _p(code.row_ptr), _p(code.col_idx), C.c_int(S), _p(payload), _p(erased),
            C.c_int(max_iter), C.c_int(1 if early_stop else 0))
    return payload, erased, it


def hybrid_single(code: Code, payload, erased, peel_iter=10, abort_writeback=False):
    payload = np.array(payload, dtype=np.uint8, order="C")
    erased = np.array(erased, dtype=np.uint8, order="C")
    n, S = payload.shape
    ro = C.c_int(0)
    st = lib().orc_ldpc_hybrid(C.c_int(code.n), C.c_int(code.k), _p(code.row_ptr), _p(code.col_idx), C.c_int(S),
                               _p(payload), _p(erased), C.c_int(peel_iter), C.c_int(1 if abort_writeback else 0),
                               C.byref(ro))
    return payload, erased, st, ro.value


def rs_mds_count(n, rs_n, rs_k, erased):
    erased = np.ascontiguousarray(erased, dtype=np.uint8)
    return lib().orc_rs_mds_count(C.c_int(n), C.c_int(rs_n), C.c_int(rs_k), _p(erased))


def gf256_tables():
    mul = np.zeros((256, 256), dtype=np.uint8)
    inv = np.zeros(255, dtype=np.uint8)
    log = np.zeros(256, dtype=np.uint8)
    alog = np.zeros(255, dtype=np.uint8)
    lib().orc_gf256_tables(_p(mul), _p(inv), _p(log), _p(alog))
    return dict(mul=mul, inv=inv, log=log, alog=alog)


def rs_gsys(n, k):
    G = np.zeros((k, n), dtype=np.uint8)
    rc = lib().orc_rs_gsys(C.c_int(n), C.c_int(k), _p(G))
    assert rc == 0
    return G


def rs_encode(G, info):
    """info [k][S] -> cw [n][S]."""
    k, n = G.shape
    info = np.ascontiguousarray(info, dtype=np.uint8)
    S = info.shape[1]
    cw = np.zeros((n, S), dtype=np.uint8)
    lib().orc_rs_encode(C.c_int(n), C.c_int(k), C.c_int(S), _p(G), _p(info), _p(cw))
    return cw


def rs_decode(G, recv_idx, recv_val):
    """recv_idx: first k received positions (0-based, ascending); recv_val [k][S] -> (info [k][S], rank_deficient)."""
    k, n = G.shape
    recv_idx = np.ascontiguousarray(recv_idx, dtype=np.int32)
    recv_val = np.ascontiguousarray(recv_val, dtype=np.uint8)
    S = recv_val.shape[1]
    out = np.zeros((k, S), dtype=np.uint8)
    rc = lib().orc_rs_decode(C.c_int(n), C.c_int(k), C.c_int(S), _p(G), _p(recv_idx), _p(recv_val), _p(out))
    return out, rc


# ---- FEC packet front-ends ---------------------------------------------------------
def packetize(cw: np.ndarray, block0: int) -> np.ndarray:
    """cw [B][n][S] -> packets [B*n][8+S] (encoder_VITA_in_UDP_out.cl:100-104,170-175)."""
    B, n, S = cw.shape
    out = np.zeros((B * n, 8 + S), dtype=np.uint8)
    lib().orc_packetize(C.c_int(n), C.c_int(S), C.c_uint32(block0), C.c_int64(B), _p(np.ascontiguousarray(cw)), _p(out))
    return out


def depacketize(packets: np.ndarray, n: int, block0: int, B: int):
    """packets [N][8+S] in arrival order -> (cw [B][n][S], flags [B][n], counts [B+1]) (decoder_with_reordering_logic.cl:59-131)."""
    packets = np.ascontiguousarray(packets)
    N, ps = packets.shape
    S = ps - 8
    cw = np.zeros((B, n, S), dtype=np.uint8)
    flags = np.zeros((B, n), dtype=np.uint8)
    counts = np.zeros(B + 1, dtype=np.uint32)
    lib().orc_depacketize(C.c_int(n), C.c_int(S), C.c_uint32(block0), C.c_int64(B), _p(packets), C.c_int64(N), _p(cw), _p(flags), _p(counts))
    return cw, flags, counts


def ready_to_decode(n, k, cur_cnt, next_cnt) -> bool:
    return bool(lib().orc_ready_to_decode(C.c_int(n), C.c_int(k), C.c_int(cur_cnt), C.c_int(next_cnt)))


def num_threads():
    return lib().orc_num_threads()


# ---- non-binary GF(256) LDPC code ---------------------------------------------------
def nb_coefficients(code: Code, seed: int) -> np.ndarray:
    """One nonzero GF(256) coefficient per nonzero of H in CSR order (ErasureCodes_NonBinaryLDPCSim.m:51-58, seeded)."""
    coef = np.zeros(len(code.col_idx), dtype=np.uint8)
    lib().orc_nb_coefficients(C.c_int64(len(coef)), C.c_uint32(seed & 0xFFFFFFFF), _p(coef))
    return coef


def nb_encode(code: Code, coef: np.ndarray, info: np.ndarray, nthreads=0) -> np.ndarray:
    info = np.ascontiguousarray(info, dtype=np.uint8)
    B, k, S = info.shape
    assert k == code.k
    cw = np.zeros((B, code.n, S), dtype=np.uint8)
    lib().orc_nb_encode_batch(C.c_int(code.n), C.c_int(code.k), _p(code.row_ptr), _p(code.col_idx), _p(np.ascontiguousarray(coef)),
                              C.c_int(S), C.c_int64(B), _p(info), _p(cw), C.c_int(nthreads))
    return cw


def nb_decode(code: Code, coef: np.ndarray, payload: np.ndarray, erased: np.ndarray, max_iter=10, mode="hybrid", nthreads=0):
    """My_LDPC_HybridML_NonBinary_Erasure_Decoder.m.  status: 0 clean after the sweeps, 1 elimination succeeded,
    2 elimination met a column without pivot, 3 (mode peel) erasures left."""
    payload = np.array(payload, dtype=np.uint8, order="C")
    erased = np.array(erased, dtype=np.uint8, order="C")
    B, n, S = payload.shape
    out = np.zeros((B, code.k, S), dtype=np.uint8)
    fail = np.zeros(B, dtype=np.uint8)
    iters = np.zeros(B, dtype=np.int32)
    status = np.zeros(B, dtype=np.int32)
    lib().orc_nb_decode_batch(C.c_int(code.n), C.c_int(code.k), _p(code.row_ptr), _p(code.col_idx), _p(np.ascontiguousarray(coef)),
                              C.c_int(S), C.c_int64(B), _p(payload), _p(erased), _p(out), _p(fail), _p(iters), _p(status),
                              C.c_int(max_iter), C.c_int({"peel": 0, "hybrid": 1}[mode]), C.c_int(nthreads))
    return dict(out=out, payload=payload, erased=erased, fail_sys=fail, iters=iters, status=status)


# ---- variable payload length and the receiver's two-buffer state machine (the rest) ----------------
def packetize_var(cw: np.ndarray, len8: np.ndarray, block0: int) -> np.ndarray:
    """Sender with num_longs_used (encoder_VITA_in_UDP_out.cl:162,186-197): header, len8 payload words, zeros."""
    pk = packetize(cw, block0)
    for i, l in enumerate(np.asarray(len8).reshape(-1)):
        pk[i, 8 + 8 * int(l):] = 0
    return pk


def rx_stream(code: Code, packets: np.ndarray, len8=None, max_iter=50, mode="peel", flush=True):
    """The per-packet loop of ldpc_erasure_decoder_with_reordering_logic.cl:71-142, one packet at a time in plain Python:
    two block buffers {current, next} (:45-50), the first usable packet names the current block (:88-91; next = current + 1
    modulo 256), a packet of any other block is dropped (:105,124), its payload words are placed and its flag cleared
    (:94-131), and after every packet the hand-off rule (:139).  On hand-off the current block is decoded as it stands and
    emitted; next becomes current, keeps its counter, and the freed buffer is cleared.  Returns [(block, out [k][S], fail)]."""
    n, k = code.n, code.k
    S = packets.shape[1] - 8
    buf = np.zeros((2, n, S), dtype=np.uint8)
    er = np.ones((2, n), dtype=np.uint8)
    cur, nxt, cur_cnt, nxt_cnt = -1, -1, 0, 0
    out = []

    def hand_off():
        nonlocal cur, nxt, cur_cnt, nxt_cnt
        d = decode(code, buf[0:1], er[0:1], max_iter=max_iter, mode=mode)
        out.append((cur, d["out"][0], int(d["fail_sys"][0])))
        buf[0], er[0] = buf[1], er[1]
        buf[1], er[1] = 0, 1
        cur, nxt = nxt, (nxt + 1) & 0xFF
        cur_cnt, nxt_cnt = nxt_cnt, 0

    for i in range(packets.shape[0]):
        h = int(np.frombuffer(packets[i, :8].tobytes(), dtype=np.uint64)[0])
        lo, hi = h & 0xFFFFFFFF, h >> 32
        cls, blk, sym = (lo >> 24) & 0xFF, (lo >> 16) & 0xFF, lo & 0xFFFF
        valid = lo == hi and cls == 1 and sym < n
        if cur < 0:
            if not valid:
                continue
            cur, nxt = blk, (blk + 1) & 0xFF
        used = S if len8 is None else 8 * int(len8[i])
        if valid and blk == cur:
            buf[0, sym, :used] = packets[i, 8:8 + used]
            er[0, sym] = 0
            cur_cnt += 1
        elif valid and blk == nxt:
            buf[1, sym, :used] = packets[i, 8:8 + used]
            er[1, sym] = 0
            nxt_cnt += 1
        if ready_to_decode(n, k, cur_cnt, nxt_cnt):
            hand_off()
    if flush:
        for _ in range(2):
            if cur >= 0 and cur_cnt > 0:
                hand_off()
    return out
