"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.  Bit-exact everywhere (byte / integer arithmetic only)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import oracle as orc  # noqa: E402


@pytest.fixture(scope="module")
def codecs():
    from ldpc_erasure_codes_b200.codec import LdpcCodec
    cache = {}

    def get(ci, S, max_batch=512):
        key = (ci, S, max_batch)
        if key not in cache:
            cache[key] = LdpcCodec(code=ci, symbol_bytes=S, device=0, max_batch=max_batch)
        return cache[key]
    yield get
    for c in cache.values():
        c.close()


def _np(t):
    return t.cpu().numpy()


def _rand_info(B, k, S, seed):
    from ldpc_erasure_codes_b200.codec import fill_random
    t = torch.empty((B, k, S), dtype=torch.uint8, device="cuda")
    fill_random(t, seed=seed)
    return t


# ------------------------------------------------------------------------------- loader
@pytest.mark.parametrize("ci", [0, 1, 2])
def test_h_loader_matches_scipy(codecs, ci):
    codec = codecs(ci, 64)
    code = orc.Code.builtin(ci)
    rp, cidx = codec.csr()
    assert np.array_equal(rp, code.row_ptr) and np.array_equal(cidx, code.col_idx)
    assert (codec.n, codec.k, codec.m) == (code.n, code.k, code.m)
    assert codec.info.encode_levels == {0: 60, 1: 27, 2: 77}[ci]   # SURVEY a-3


# ------------------------------------------------------------------------------- channel
@pytest.mark.parametrize("ci,P,frame0", [(0, 19, 0), (1, 13, 0), (1, 9, 1000003), (2, 19, 5), (1, 0, 0), (1, 64, 0)])
def test_iid64_generator_bit_exact(codecs, ci, P, frame0):
    from ldpc_erasure_codes_b200.codec import unpack_mask
    codec = codecs(ci, 64)
    B = 37
    mask = codec.gen_erasures(B, 12345, P=P, frame0=frame0)
    ref = orc.gen_erasures_iid(codec.n, 12345, B, P=P, frame0=frame0)
    assert np.array_equal(unpack_mask(mask, codec.n), ref)
    # padding bits above n are zero
    m = _np(mask).view(np.uint32)
    if codec.n % 32:
        assert not (m[:, -1] >> (codec.n % 32)).any()


def test_iid32_and_counter_wrap(codecs):
    from ldpc_erasure_codes_b200.codec import unpack_mask
    codec = codecs(1, 64)
    thr = int(0.2 * 2 ** 32)
    frame0 = (2 ** 32) // 2040 - 3          # counter wraps inside this batch
    mask = codec.gen_erasures(8, 777, p=0.2, frame0=frame0)
    ref = orc.gen_erasures_iid(2040, 777, 8, p32=thr, frame0=frame0)
    assert np.array_equal(unpack_mask(mask, 2040), ref)


@pytest.mark.parametrize("params", [(0.001, 0.1, 10.0), (0.02, 0.4, 10.0), (0.3, 0.9, 2.0)])
def test_bursty_generator_bit_exact_and_shard_independent(codecs, params):
    from ldpc_erasure_codes_b200.codec import unpack_mask
    codec = codecs(2, 64)
    B = 24
    ref, _ = orc.gen_erasures_bursty(4000, 99, B, *params)
    mask = codec.gen_erasures(B, 99, bursty=params)
    assert np.array_equal(unpack_mask(mask, 4000), ref)
    # a shard that starts in the middle of the stream reproduces the same frames
    part = codec.gen_erasures(B - 10, 99, bursty=params, frame0=10)
    assert np.array_equal(unpack_mask(part, 4000), ref[10:])


# ------------------------------------------------------------------------------- encoder
@pytest.mark.parametrize("ci,S,B", [(0, 64, 20), (1, 64, 300), (2, 64, 20), (1, 16, 33), (1, 128, 9), (0, 1024, 5)])
def test_encode_bit_exact(codecs, ci, S, B):
    codec = codecs(ci, S)
    info = _rand_info(B, codec.k, S, seed=ci * 100 + S)
    cw = codec.encode(info)
    ref = orc.encode(orc.Code.builtin(ci), _np(info))
    assert np.array_equal(_np(cw), ref)


def test_encode_survey_crosscheck(codecs):
    import zlib
    codec = codecs(1, 16)
    v = (np.arange(1, codec.k + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)).view(np.uint8).reshape(codec.k, 8)
    info = np.tile(v, (1, 2))[None]
    cw = _np(codec.encode(torch.from_numpy(info).cuda()))[0]
    assert zlib.crc32(np.ascontiguousarray(cw[:, :8]).tobytes()) == 0x91491f2f   # SURVEY 8(c), (2040,1530)


# ------------------------------------------------------------------------------- peel decode
def _scenario(codec, ci, B, P, seed, S, valid=True):
    code = orc.Code.builtin(ci)
    if valid:
        info = _rand_info(B, codec.k, S, seed=seed)
        cw = codec.encode(info)
    else:   # arbitrary bytes, not a codeword: only an exact replay of the serial schedule matches
        cw = _rand_info(B, codec.n, S, seed=seed + 1)
    rx = cw.clone()
    mask = codec.gen_erasures(B, seed, P=P, payload=rx)
    flags = orc.gen_erasures_iid(code.n, seed, B, P=P)
    return code, cw, rx, mask, flags


@pytest.mark.parametrize("ci,S,P", [(1, 64, 6), (1, 64, 10), (1, 64, 13), (1, 64, 16), (1, 64, 19),
                                    (0, 64, 19), (0, 64, 24), (2, 64, 19), (1, 16, 13), (1, 32, 13), (1, 128, 12),
                                    (0, 1024, 19)])
def test_peel_decode_bit_exact(codecs, ci, S, P):
    codec = codecs(ci, S)
    B = 12 if S >= 1024 else 200
    code, cw, rx, mask, flags = _scenario(codec, ci, B, P, 4242 + P, S)
    out, fail = codec.decode(rx, mask, max_iter=50)
    ref = orc.decode(code, _np(rx), flags, max_iter=50)
    assert np.array_equal(_np(fail), ref["fail_sys"])
    assert np.array_equal(_np(out), ref["out"])
    good = ref["fail_sys"] == 0
    assert np.array_equal(_np(out)[good], _np(cw)[good][:, :code.k])


@pytest.mark.parametrize("max_iter", [0, 1, 2, 3, 7, 10])
def test_peel_decode_iteration_cap_is_exact(codecs, max_iter):
    """A binding num_iter gives the reference's partial result: the serial sweep order is replayed exactly."""
    codec = codecs(1, 64)
    code, cw, rx, mask, flags = _scenario(codec, 1, 150, 12, 99, 64)
    out, fail = codec.decode(rx, mask, max_iter=max_iter)
    ref = orc.decode(code, _np(rx), flags, max_iter=max_iter, early_stop=False)
    assert np.array_equal(_np(fail), ref["fail_sys"]) and np.array_equal(_np(out), ref["out"])


@pytest.mark.parametrize("ci,P", [(1, 12), (0, 22)])
def test_peel_decode_exact_on_non_codeword_input(codecs, ci, P):
    codec = codecs(ci, 64)
    code, cw, rx, mask, flags = _scenario(codec, ci, 120, P, 31337, 64, valid=False)
    out, fail = codec.decode(rx, mask, max_iter=50)
    ref = orc.decode(code, _np(rx), flags, max_iter=50)
    assert np.array_equal(_np(fail), ref["fail_sys"]) and np.array_equal(_np(out), ref["out"])


def test_peel_edge_cases(codecs):
    from ldpc_erasure_codes_b200.codec import pack_mask
    codec = codecs(1, 64)
    code = orc.Code.builtin(1)
    info = _rand_info(4, codec.k, 64, seed=5)
    cw = codec.encode(info)
    flags = np.zeros((4, code.n), np.uint8)
    flags[1, :] = 1                      # everything erased
    flags[2, code.k:] = 1                # all parities erased
    flags[3, :code.k:2] = 1              # every other systematic symbol
    rx = _np(cw).copy()
    rx[flags == 1] = 0
    out, fail = codec.decode(torch.from_numpy(rx).cuda(), torch.from_numpy(pack_mask(flags)).cuda(), max_iter=50)
    ref = orc.decode(code, rx, flags, max_iter=50)
    assert np.array_equal(_np(fail), ref["fail_sys"]) and np.array_equal(_np(out), ref["out"])
    assert list(_np(fail)) == [0, 1, 0, int(ref["fail_sys"][3])]
    # empty batch is a no-op
    e_out, e_fail = codec.decode(torch.empty((0, code.n, 64), dtype=torch.uint8, device="cuda"),
                                 torch.empty((0, codec.mask_words), dtype=torch.int32, device="cuda"))
    assert e_out.shape[0] == 0


@pytest.mark.parametrize("B", [1, 147, 149, 700])
def test_ragged_batches_and_chunking(codecs, B):
    codec = codecs(1, 64, max_batch=256)     # 700 > max_batch: internal chunking
    code, cw, rx, mask, flags = _scenario(codec, 1, B, 13, 1000 + B, 64)
    out, fail = codec.decode(rx, mask)
    ref = orc.decode(code, _np(rx), flags, max_iter=50)
    assert np.array_equal(_np(fail), ref["fail_sys"]) and np.array_equal(_np(out), ref["out"])


def test_chunk_pipeline_matches_serial(codecs, monkeypatch):
    """A batch of several max_batch chunks alternates between two internal streams; same bytes as chunk after chunk,
    and the caller's stream sees the results when the call's work is done."""
    codec = codecs(1, 32, 512)
    code = orc.Code.builtin(1)
    B = 512 * 5 + 77
    info = _rand_info(B, codec.k, 32, seed=21)
    cw = codec.encode(info)
    rx = cw.clone()
    mask = codec.gen_erasures(B, 2121, P=12, payload=rx)
    out1, f1 = codec.decode(rx, mask)
    chk = out1.sum(dtype=torch.int64)                  # queued on the caller's stream right behind the call
    monkeypatch.setenv("LDPC_CUDA_CHUNK_PIPELINE", "0")
    out2, f2 = codec.decode(rx, mask)
    assert int(chk) == int(out2.sum(dtype=torch.int64))
    assert torch.equal(out1, out2) and torch.equal(f1, f2)
    flags = orc.gen_erasures_iid(code.n, 2121, B, P=12)
    ref = orc.decode(code, _np(rx), flags, max_iter=50, mode="peel")
    assert np.array_equal(_np(f1), ref["fail_sys"]) and np.array_equal(_np(out1), ref["out"])


@pytest.mark.parametrize("W,slots", [(16, 0), (32, 0), (64, 0), (32, 1), (32, 2), (16, 2)])
def test_executor_geometries_agree(codecs, W, slots):
    from ldpc_erasure_codes_b200.codec import LdpcCodec
    codec = LdpcCodec(code=1, symbol_bytes=64, device=0, max_batch=512)
    codec.set_exec_geometry(W, slots)
    assert codec.info.slice_bytes == W
    code, cw, rx, mask, flags = _scenario(codec, 1, 300, 13, 77, 64)
    out, fail = codec.decode(rx, mask)
    ref = orc.decode(code, _np(rx), flags, max_iter=50)
    assert np.array_equal(_np(out), ref["out"]) and np.array_equal(_np(fail), ref["fail_sys"])
    cw2 = codec.encode(_rand_info(50, codec.k, 64, seed=3))
    assert np.array_equal(_np(cw2), orc.encode(code, _np(_rand_info(50, codec.k, 64, seed=3))))
    codec.close()


def test_statistics_match_reference_counters(codecs):
    codec = codecs(1, 64)
    code, cw, rx, mask, flags = _scenario(codec, 1, 400, 13, 2024, 64)
    codec.reset_stats()
    out, fail = codec.decode(rx, mask)
    st = codec.stats()
    ref = orc.decode(code, _np(rx), flags, max_iter=50)
    assert st["frames"] == 400 and st["ldpc_errors"] == int(ref["fail_sys"].sum())
    assert st["rs_errors"] == sum(orc.rs_mds_count(code.n, 255, 192, f) for f in flags)   # perf_tests.cl:70-80


def test_host_buffer_entry_points(codecs):
    codec = codecs(1, 64, max_batch=256)
    code, cw, rx, mask, flags = _scenario(codec, 1, 600, 12, 8, 64)
    h_out, h_fail = codec.decode_host(rx.cpu().pin_memory(), mask.cpu().pin_memory())
    ref = orc.decode(code, _np(rx), flags, max_iter=50)
    assert np.array_equal(h_out.numpy(), ref["out"]) and np.array_equal(h_fail.numpy(), ref["fail_sys"])
    info = _rand_info(600, codec.k, 64, seed=21)
    h_cw = codec.encode_host(info.cpu().pin_memory())
    assert np.array_equal(h_cw.numpy(), orc.encode(code, _np(info)))


def test_large_batch_round_trip_properties(codecs):
    """Full-size style check that does not need the oracle: encode -> erase -> decode returns the input wherever
    the decoder reports success; decoding is idempotent; XOR-linearity of the encoder."""
    from ldpc_erasure_codes_b200.codec import LdpcCodec
    codec = LdpcCodec(code=1, symbol_bytes=64, device=0, max_batch=16384)
    B = 20000
    a = _rand_info(B, codec.k, 64, seed=1)
    b = _rand_info(B, codec.k, 64, seed=2)
    ca, cb = codec.encode(a), codec.encode(b)
    assert bool((codec.encode(a ^ b) == (ca ^ cb)).all())
    rx = ca.clone()
    mask = codec.gen_erasures(B, 5, P=12, payload=rx)
    out, fail = codec.decode(rx, mask)
    good = fail == 0
    assert 0.9 < float(good.float().mean()) <= 1.0
    assert bool((out[good] == a[good]).all())
    out2, fail2 = codec.decode(rx, mask)
    assert bool((out2 == out).all()) and bool((fail2 == fail).all())
    codec.close()


# ------------------------------------------------------------------------------- a code the user brings
def _random_h(n, m, dv, tri, seed):
    """Sparse H: [random part | lower-triangular (tri) or random parity part], column weight ~dv."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    k = n - m
    H = np.zeros((m, n), np.uint8)
    for v in range(k):
        H[rng.choice(m, dv, replace=False), v] = 1
    for r in range(m):
        if tri:
            H[r, k + r] = 1
            if r:
                lo = max(0, r - 6)                  # banded: keeps the column weights small
                H[r, k + lo + rng.choice(r - lo, min(r - lo, 2), replace=False)] = 1
        else:
            H[rng.choice(m, dv, replace=False), k + r] = 1
    for r in range(m):                       # no empty rows
        if not H[r, :k].any():
            H[r, rng.integers(k)] = 1
    return sp.csc_matrix(H.astype(np.float64))


@pytest.mark.parametrize("n,m,S", [(96, 48, 16), (333, 100, 48), (1000, 250, 64)])
def test_user_supplied_triangular_code(tmp_path, n, m, S):
    """ldpc_ctx_create(path): a MAT-v5 file that is not one of the built-ins; encode, erase, peel and hybrid decode."""
    import scipy.io as sio
    from ldpc_erasure_codes_b200.codec import LdpcCodec
    H = _random_h(n, m, 3, True, seed=n)
    path = str(tmp_path / "user_code.mat")
    sio.savemat(path, {"H_sparse": H}, do_compression=True)
    codec = LdpcCodec(code=path, symbol_bytes=S, device=0, max_batch=128)
    assert (codec.n, codec.k, codec.m) == (n, n - m, m)
    code = orc.Code(H)
    B = 77
    info = _rand_info(B, codec.k, S, seed=n)
    cw = codec.encode(info)
    assert np.array_equal(_np(cw), orc.encode(code, _np(info)))
    for P, mode, it in ((8, "peel", 50), (20, "peel", 3), (18, "hybrid", 10)):
        rx = cw.clone()
        mask = codec.gen_erasures(B, 31 + P, P=P, payload=rx)
        flags = orc.gen_erasures_iid(n, 31 + P, B, P=P)
        out, fail = codec.decode(rx, mask, max_iter=it, mode=mode)
        ref = orc.decode(code, _np(rx), flags, max_iter=it, mode=mode)
        assert np.array_equal(_np(fail), ref["fail_sys"]) and np.array_equal(_np(out), ref["out"])
    codec.close()


def test_generated_code_round_trip(tmp_path):
    """a code from the library's own girth-8 generator (Hgen_irregularDegree...m), saved as the
    MAT-v5 file the reference saves, through the loader, the encoder and both decoders against the oracle."""
    from ldpc_erasure_codes_b200 import hgen
    from ldpc_erasure_codes_b200.codec import LdpcCodec
    H, _ = hgen.generate([(300, 4)], [(600, 2)], seed=5)
    assert hgen.count_short_cycles(H) == (0, 0)
    path = str(tmp_path / "generated.mat")
    hgen.save_mat(path, H)
    S, B = 32, 90
    codec = LdpcCodec(code=path, symbol_bytes=S, device=0, max_batch=64)
    assert (codec.n, codec.k) == (600, 300)
    code = orc.Code(H)
    info = _rand_info(B, codec.k, S, seed=3)
    cw = codec.encode(info)
    assert np.array_equal(_np(cw), orc.encode(code, _np(info)))
    for P, mode, it in ((12, "peel", 50), (26, "hybrid", 10)):
        rx = cw.clone()
        mask = codec.gen_erasures(B, 100 + P, P=P, payload=rx)
        flags = orc.gen_erasures_iid(codec.n, 100 + P, B, P=P)
        out, fail = codec.decode(rx, mask, max_iter=it, mode=mode)
        ref = orc.decode(code, _np(rx), flags, max_iter=it, mode=mode)
        assert np.array_equal(_np(fail), ref["fail_sys"]) and np.array_equal(_np(out), ref["out"])
    codec.close()


def test_user_supplied_non_triangular_code_decodes_but_does_not_encode(tmp_path):
    import scipy.io as sio
    from ldpc_erasure_codes_b200 import _lib as L
    from ldpc_erasure_codes_b200.codec import LdpcCodec
    n, m, S = 200, 80, 32
    H = _random_h(n, m, 3, False, seed=5)
    path = str(tmp_path / "user_code.mat")
    sio.savemat(path, {"H_sparse": H})
    codec = LdpcCodec(code=path, symbol_bytes=S, device=0, max_batch=64)
    info = _rand_info(4, codec.k, S, seed=1)
    with pytest.raises(L.LdpcCudaError) as ei:
        codec.encode(info)
    assert ei.value.code == -7                                  # LDPC_ERR_NOT_TRIANGULAR
    code = orc.Code(H)
    rx = _rand_info(40, n, S, seed=2)                           # arbitrary words: the replay is exact on any input
    mask = codec.gen_erasures(40, 9, P=10, payload=rx)
    flags = orc.gen_erasures_iid(n, 9, 40, P=10)
    out, fail = codec.decode(rx, mask, max_iter=50, mode="peel")
    ref = orc.decode(code, _np(rx), flags, max_iter=50, mode="peel")
    assert np.array_equal(_np(fail), ref["fail_sys"]) and np.array_equal(_np(out), ref["out"])
    codec.close()


def test_hybrid_does_not_rely_on_zeroed_erasures(codecs, monkeypatch):
    """Erased positions of the input hold arbitrary bytes instead of zeros: the decoders never read them (the
    executor's syndromes include them, the elimination stage XORs them out again)."""
    codec = codecs(1, 64)
    code = orc.Code.builtin(1)
    B = 120
    info = _rand_info(B, codec.k, 64, seed=8)
    cw = codec.encode(info)
    mask = codec.gen_erasures(B, 808, P=13)
    flags = orc.gen_erasures_iid(code.n, 808, B, P=13)
    rx = _np(cw).copy()
    junk = np.random.default_rng(3).integers(0, 256, rx.shape, dtype=np.uint8)
    rx[flags == 1] = junk[flags == 1]
    rxd = torch.from_numpy(rx).cuda()
    ref = orc.decode(code, rx, flags, max_iter=10, mode="hybrid")
    for stages in ("3", "2", "0"):
        monkeypatch.setenv("LDPC_CUDA_GE_STAGES", stages)
        out, fail = codec.decode(rxd, mask, max_iter=10, mode="hybrid")
        good = ref["fail_sys"] == 0
        assert np.array_equal(_np(fail), ref["fail_sys"])
        assert np.array_equal(_np(out)[good], _np(info)[good])


def test_hybrid_reference_symbol_size(codecs):
    """S = 1024 (the reference's 128 x u64 symbols): the elimination stage walks the payload 64 bytes at a time."""
    codec = codecs(1, 1024, 64)
    code, cw, rx, mask, flags = _scenario(codec, 1, 24, 13, 2025, 1024)
    ref = _hybrid_check(codec, code, rx, mask, flags)
    assert (ref["status"] == 1).sum() > 2


# ------------------------------------------------------------------------------- hybrid-ML (GF(2) elimination)
def _hybrid_check(codec, code, rx, mask, flags, max_iter=10):
    codec.reset_stats()
    out, fail = codec.decode(rx, mask, max_iter=max_iter, mode="hybrid")
    st = codec.stats()
    ref = orc.decode(code, _np(rx), flags, max_iter=max_iter, mode="hybrid")
    assert np.array_equal(_np(fail), ref["fail_sys"])
    assert np.array_equal(_np(out), ref["out"])
    assert st["ml_attempts"] == int((ref["status"] > 0).sum())
    assert st["ml_failures"] == int((ref["status"] == 2).sum())
    return ref


@pytest.mark.parametrize("P", [10, 12, 13, 14, 16])
def test_hybrid_decode_bit_exact(codecs, P):
    """Matlab/My_LDPC_HybridML_Erasure_Decoder.m: 10 sweeps, then elimination on the residual set."""
    codec = codecs(1, 64)
    code, cw, rx, mask, flags = _scenario(codec, 1, 160, P, 500 + P, 64)
    ref = _hybrid_check(codec, code, rx, mask, flags)
    good = ref["fail_sys"] == 0
    assert np.array_equal(ref["out"][good], _np(cw)[good][:, :code.k])
    if P == 13:
        assert (ref["status"] == 1).sum() > 20      # the elimination stage really ran and succeeded


@pytest.mark.parametrize("mode,it", [("peel", 50), ("hybrid", 10)])
def test_executor_plain_form_matches(monkeypatch, mode, it):
    """The executor's fallback for a blob that leaves no room for its pass table (plain level-by-level gathers), forced."""
    from ldpc_erasure_codes_b200.codec import LdpcCodec
    monkeypatch.setenv("LDPC_CUDA_EXEC_PLAIN", "1")
    codec = LdpcCodec(code=1, symbol_bytes=64, device=0, max_batch=128)
    code, cw, rx, mask, flags = _scenario(codec, 1, 150, 13, 91, 64)
    out, fail = codec.decode(rx, mask, max_iter=it, mode=mode)
    ref = orc.decode(code, _np(rx), flags, max_iter=it, mode=mode)
    assert np.array_equal(_np(fail), ref["fail_sys"]) and np.array_equal(_np(out), ref["out"])
    codec.close()


def test_hybrid_batch_larger_than_small_max_batch():
    """A context created with max_batch < 256: hybrid chunks must not exceed it (the schedule scratch is sized for
    max_batch codewords -- a round-2 regression found by tools/sanitize_pass.py: illegal address)."""
    from ldpc_erasure_codes_b200.codec import LdpcCodec
    codec = LdpcCodec(code=1, symbol_bytes=64, device=0, max_batch=16)
    code, cw, rx, mask, flags = _scenario(codec, 1, 40, 13, 77, 64)
    _hybrid_check(codec, code, rx, mask, flags)
    codec.close()


def test_hybrid_survey_frames(codecs):
    """(2040,1530), P=13/64, seed 12345: frame 0 (447 erasures) and frame 2 need elimination."""
    codec = codecs(1, 16)
    code = orc.Code.builtin(1)
    v = (np.arange(1, code.k + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)).view(np.uint8).reshape(code.k, 8)
    info = torch.from_numpy(np.tile(np.tile(v, (1, 2))[None], (3, 1, 1)).copy()).cuda()
    cw = codec.encode(info)
    rx = cw.clone()
    mask = codec.gen_erasures(3, 12345, P=13, payload=rx)
    flags = orc.gen_erasures_iid(code.n, 12345, 3, P=13)
    assert list(flags.sum(axis=1)) == [447, 387, 432]
    out, fail = codec.decode(rx, mask, max_iter=10, mode="hybrid")
    assert not _np(fail).any() and np.array_equal(_np(out), _np(info))
    # the same frames from the peeling fixed point (cap not binding) give the same unique answer
    out2, fail2 = codec.decode(rx, mask, max_iter=1000, mode="hybrid")
    assert np.array_equal(_np(out2), _np(info))


@pytest.mark.parametrize("S", [16, 128])
def test_hybrid_other_symbol_sizes(codecs, S):
    codec = codecs(1, S)
    code, cw, rx, mask, flags = _scenario(codec, 1, 60, 13, 77, S)
    _hybrid_check(codec, code, rx, mask, flags)


def test_hybrid_n2000_global_workspace(codecs):
    codec = codecs(0, 64)
    code, cw, rx, mask, flags = _scenario(codec, 0, 40, 27, 5, 64)     # 42 %: beyond the peeling threshold
    ref = _hybrid_check(codec, code, rx, mask, flags)
    assert (ref["status"] > 0).any()


def test_hybrid_n4000_bursty_stress(codecs):
    """Config 3: n4000/k2000 under the two-state channel, parameters that do reach stopping sets."""
    from ldpc_erasure_codes_b200.codec import unpack_mask
    codec = codecs(2, 64)
    code = orc.Code.builtin(2)
    B = 10
    info = _rand_info(B, codec.k, 64, seed=3)
    cw = codec.encode(info)
    rx = cw.clone()
    mask = codec.gen_erasures(B, 2024, bursty=(0.38, 0.9, 10.0), payload=rx)
    flags = unpack_mask(mask, code.n)
    ref = _hybrid_check(codec, code, rx, mask, flags)
    assert (ref["status"] > 0).any()


def test_hybrid_warp_and_cta_kernels_agree(codecs, monkeypatch):
    """Inactivation decoding, per-warp elimination and the CTA-per-codeword kernel solve the same systems."""
    codec = codecs(1, 64)
    code, cw, rx, mask, flags = _scenario(codec, 1, 300, 13, 4711, 64)
    ref = _hybrid_check(codec, code, rx, mask, flags)
    assert (ref["status"] == 1).sum() > 40
    for stages in ("0", "2", "1"):          # CTA kernel only / per-warp elimination / inactivation decoding
        monkeypatch.setenv("LDPC_CUDA_GE_STAGES", stages)
        _hybrid_check(codec, code, rx, mask, flags)
    monkeypatch.setenv("LDPC_CUDA_GE_STAGES", "3")
    monkeypatch.setenv("LDPC_CUDA_GE_SPLIT", "0")   # inactivation stage as one kernel (pattern + payload together)
    _hybrid_check(codec, code, rx, mask, flags)


@pytest.mark.parametrize("wpc", ["16", "7"])
def test_hybrid_fall_through_when_slots_are_small(monkeypatch, wpc):
    """Small shared-memory slots: stage 1 defers what does not fit to stage 2 / the CTA kernel; same answers."""
    from ldpc_erasure_codes_b200.codec import LdpcCodec
    monkeypatch.setenv("LDPC_CUDA_GE_WPC", wpc)
    codec = LdpcCodec(code=1, symbol_bytes=32, device=0, max_batch=256)
    code, cw, rx, mask, flags = _scenario(codec, 1, 200, 13, 99, 32)
    ref = _hybrid_check(codec, code, rx, mask, flags)
    assert (ref["status"] == 1).sum() > 20
    codec.close()


# ------------------------------------------------------------------------------- Reed-Solomon GF(256)
@pytest.fixture(scope="module")
def rs_codecs():
    from ldpc_erasure_codes_b200.codec import RsCodec
    cache = {}

    def get(n, k, S):
        if (n, k, S) not in cache:
            cache[(n, k, S)] = RsCodec(n=n, k=k, symbol_bytes=S, device=0, max_batch=256)
        return cache[(n, k, S)]
    yield get
    for c in cache.values():
        c.close()


@pytest.mark.parametrize("nk", [(255, 191), (255, 192), (250, 125), (7, 5)])
def test_rs_generator_matches_oracle(rs_codecs, nk):
    n, k = nk
    assert np.array_equal(rs_codecs(n, k, 16).generator(), orc.rs_gsys(n, k))


@pytest.mark.parametrize("n,k,S", [(255, 191, 1024), (255, 192, 64), (7, 5, 16), (250, 125, 48), (100, 90, 80)])
def test_rs_encode_bit_exact(rs_codecs, n, k, S):
    codec = rs_codecs(n, k, S)
    info = _rand_info(5, k, S, seed=n + k)
    cw = _np(codec.encode(info))
    G = orc.rs_gsys(n, k)
    for b in range(5):
        assert np.array_equal(cw[b], orc.rs_encode(G, _np(info)[b]))


@pytest.mark.parametrize("n,k,S,p", [(255, 191, 1024, 0.2), (255, 192, 64, 0.2), (255, 191, 64, 0.05), (7, 5, 16, 0.25),
                                     (250, 125, 48, 0.45), (100, 90, 80, 0.08)])
def test_rs_decode_bit_exact(rs_codecs, n, k, S, p):
    """Decoder == restated Matlab/My_RS_Decode_Optimize_With_GFTables.m on the first k received symbols."""
    from ldpc_erasure_codes_b200.codec import pack_mask
    codec = rs_codecs(n, k, S)
    B = 24
    rng = np.random.default_rng(n * 7 + k)
    info = _rand_info(B, k, S, seed=k)
    cw = codec.encode(info)
    flags = (rng.random((B, n)) < p).astype(np.uint8)
    flags[0] = 0                                  # nothing erased
    flags[1] = 0; flags[1, :n - k] = 1            # exactly n - k systematic symbols erased: still decodable
    flags[2] = 0; flags[2, :n - k + 1] = 1        # one too many: undecodable
    flags[3] = 0; flags[3, k:] = 1                # all repair symbols erased
    rx = _np(cw).copy()
    rx[flags == 1] = 0
    out, fail = codec.decode(torch.from_numpy(rx).cuda(), torch.from_numpy(pack_mask(flags)).cuda())
    out, fail = _np(out), _np(fail)
    G = orc.rs_gsys(n, k)
    for b in range(B):
        rec = np.nonzero(flags[b] == 0)[0]
        if len(rec) >= k:
            ref, rd = orc.rs_decode(G, rec[:k].astype(np.int32), rx[b][rec[:k]])
            assert rd == 0 and fail[b] == 0
            assert np.array_equal(out[b], ref) and np.array_equal(ref, _np(info)[b])
        else:
            assert fail[b] == 1
            want = _np(info)[b].copy()
            want[flags[b, :k] == 1] = 0
            assert np.array_equal(out[b], want)
    assert fail[2] == 1 and fail[1] == 0


# ------------------------------------------------------------------------------- error-rate run (counters only)
@pytest.mark.parametrize("mode,max_iter", [("peel", 50), ("hybrid", 10)])
def test_simulate_fer_matches_explicit_decode_and_oracle(codecs, mode, max_iter):
    """ldpc_simulate_fer = the reference's committed flow (all-zero codewords, counters only): the same counters as
    generating the masks, decoding zero payloads explicitly, and as the oracle on those masks."""
    codec = codecs(1, 16)
    code = orc.Code.builtin(1)
    B, seed, P = 700, 424242, 13        # 700 > max_batch 512: chunked, frame counter continues across chunks
    codec.reset_stats()
    codec.simulate_fer(B, seed, P=P, max_iter=max_iter, mode=mode)
    sim = codec.stats()
    rx = torch.zeros((B, code.n, 16), dtype=torch.uint8, device="cuda")
    mask = codec.gen_erasures(B, seed, P=P)
    codec.reset_stats()
    out, fail = codec.decode(rx, mask, max_iter=max_iter, mode=mode)
    assert codec.stats() == sim
    flags = orc.gen_erasures_iid(code.n, seed, B, P=P)
    ref = orc.decode(code, np.zeros((B, code.n, 16), np.uint8), flags, max_iter=max_iter, mode=mode)
    assert sim["frames"] == B
    assert sim["ldpc_errors"] - sim["ml_recovered"] == int(ref["fail_sys"].sum()) == int(_np(fail).sum())
    assert sim["rs_errors"] == sum(orc.rs_mds_count(code.n, 255, 192, f) for f in flags)
    assert not _np(out).any()           # the all-zero codeword decodes to zeros


def test_simulate_fer_bursty_channel_and_host_hybrid(codecs):
    """The two-state channel through the counters-only run, and the hybrid decoder through the host-buffer entry point."""
    from ldpc_erasure_codes_b200.codec import unpack_mask
    codec = codecs(1, 32)
    code = orc.Code.builtin(1)
    B, seed, chan = 300, 77, (0.02, 0.08, 10.0)
    codec.reset_stats()
    codec.simulate_fer(B, seed, bursty=chan, max_iter=10, mode="hybrid")
    sim = codec.stats()
    mask = codec.gen_erasures(B, seed, bursty=chan)
    flags = unpack_mask(mask, code.n)
    assert np.array_equal(flags, orc.gen_erasures_bursty(code.n, seed, B, *chan)[0])
    ref = orc.decode(code, np.zeros((B, code.n, 32), np.uint8), flags, max_iter=10, mode="hybrid")
    assert sim["frames"] == B and sim["ml_attempts"] == int((ref["status"] > 0).sum())
    assert sim["ldpc_errors"] - sim["ml_recovered"] == int(ref["fail_sys"].sum())
    # host buffers in, host buffers out, hybrid mode
    info = _rand_info(B, codec.k, 32, seed=5)
    cw = codec.encode(info)
    rx = cw.clone()
    mask = codec.gen_erasures(B, seed, bursty=chan, payload=rx)
    out_h, fail_h = codec.decode_host(rx.cpu().pin_memory(), mask.cpu().pin_memory(), max_iter=10, mode="hybrid")
    ref = orc.decode(code, _np(rx), flags, max_iter=10, mode="hybrid")
    assert np.array_equal(out_h.numpy(), ref["out"]) and np.array_equal(fail_h.numpy(), ref["fail_sys"])


# ------------------------------------------------------------------------------- FEC packet front-ends
@pytest.mark.parametrize("ci,S,B,block0", [(1, 64, 6, 0), (1, 16, 9, 250), (0, 1024, 2, 255), (2, 32, 3, 17)])
def test_packetize_and_depacketize_bit_exact(codecs, ci, S, B, block0):
    """Sender / receiver data formats of the reference's front-end kernels against the oracle: any arrival order, losses,
    duplicates, packets of foreign blocks, a corrupted header; block numbers wrap modulo 256."""
    from ldpc_erasure_codes_b200.codec import unpack_mask
    codec = codecs(ci, S)
    n = codec.n
    cw = _rand_info(B, n, S, seed=5 + ci)
    pk = codec.packetize(cw, block0)
    ref_pk = orc.packetize(_np(cw), block0)
    assert np.array_equal(_np(pk), ref_pk)
    rng = np.random.default_rng(9)
    order = rng.permutation(B * n)
    keep = rng.random(B * n) >= 0.25
    foreign = orc.packetize(_np(cw[:1]), (block0 + B + 40) & 0xFF)[:50]
    bad = ref_pk[7:8].copy(); bad[0, 7] ^= 0x01
    stream = np.concatenate([ref_pk[order][keep[order]], ref_pk[:11], foreign, bad])
    got_cw, got_mask, got_cnt = codec.depacketize(torch.from_numpy(stream).cuda(), block0, B)
    ref_cw, ref_flags, ref_cnt = orc.depacketize(stream, n, block0, B)
    assert np.array_equal(_np(got_cw), ref_cw)
    assert np.array_equal(unpack_mask(got_mask, n), ref_flags)
    assert np.array_equal(_np(got_cnt).astype(np.uint32), ref_cnt) and ref_cnt[B] == 51
    # nothing received at all: everything erased, zero payload
    e_cw, e_mask, e_cnt = codec.depacketize(torch.zeros((0, 8 + S), dtype=torch.uint8, device="cuda"), block0, B)
    assert not _np(e_cw).any() and unpack_mask(e_mask, n).all() and not _np(e_cnt).any()


def test_packet_stream_end_to_end(codecs):
    """encode -> packets -> a lossy, reordering network -> reassembly -> decode: the codec behind its front-ends."""
    from ldpc_erasure_codes_b200.codec import unpack_mask
    codec = codecs(1, 64)
    code = orc.Code.builtin(1)
    B, block0 = 40, 236                                   # wraps past 255
    info = _rand_info(B, codec.k, 64, seed=77)
    packets = _np(codec.packetize(codec.encode(info), block0))
    rng = np.random.default_rng(1)
    arrived = packets[rng.permutation(len(packets))]
    arrived = arrived[rng.random(len(arrived)) >= 0.15]   # 15 % packet loss
    rx, mask, counts = codec.depacketize(torch.from_numpy(arrived).cuda(), block0, B)
    assert int(_np(counts)[:B].sum()) == len(arrived) and int(_np(counts)[B]) == 0
    out, fail = codec.decode(rx, mask)
    flags = unpack_mask(mask, code.n)
    ref = orc.decode(code, _np(rx), flags, max_iter=50, mode="peel")
    assert np.array_equal(_np(out), ref["out"]) and np.array_equal(_np(fail), ref["fail_sys"])
    good = _np(fail) == 0
    assert good.sum() >= B - 1 and np.array_equal(_np(out)[good], _np(info)[good])
    n, k = codec.n, codec.k
    assert codec.ready_to_decode(n, 0) and not codec.ready_to_decode(k + 408, 11) and codec.ready_to_decode(k + 409, 11)
    assert all(codec.ready_to_decode(c, x) == orc.ready_to_decode(n, k, c, x) for c in range(1500, 2041, 7) for x in (0, 10, 11, 100, 101))


# ------------------------------------------------------------------------------- second failure criterion (a-8)
@pytest.mark.parametrize("mode,it,P", [("peel", 50, 13), ("peel", 2, 10), ("hybrid", 10, 13), ("hybrid", 10, 16)])
def test_fail_any_matches_oracle(codecs, mode, it, P):
    """d_fail_any = any of the n symbols still unknown (LDPCErasureCodes_MessagePassingAlgSim.m:229-236);
    d_fail = any of the first k (perf_tests.cl:215-228)."""
    codec = codecs(1, 64)
    code, cw, rx, mask, flags = _scenario(codec, 1, 300, P, 606 + P, 64)
    fail_any = torch.full((300,), 7, dtype=torch.uint8, device="cuda")
    codec.reset_stats()
    out, fail = codec.decode(rx, mask, max_iter=it, mode=mode, fail_any=fail_any)
    ref = orc.decode(code, _np(rx), flags, max_iter=it, mode=mode)
    ref_any = (ref["erased"].max(axis=1) > 0).astype(np.uint8)
    assert np.array_equal(_np(fail), ref["fail_sys"])
    assert np.array_equal(_np(fail_any), ref_any)
    assert (ref_any >= ref["fail_sys"]).all()
    st = codec.stats()
    assert st["any_errors"] == int(ref_any.sum())
    if mode == "peel" and P == 13:
        assert ref_any.sum() > ref["fail_sys"].sum() or ref_any.sum() > 0    # the case is not vacuous


# ------------------------------------------------------------------------------- the benchmark's geometry
@pytest.mark.parametrize("mode,it", [("peel", 50), ("hybrid", 10)])
def test_benchmark_geometry_against_oracle(mode, it):
    """bench.py's launch shape -- max_batch 65536, a call of 2 x 65536 codewords (two chunks: two internal streams and two
    scratch sets in peel mode), 148 persistent CTAs claiming work dynamically -- compared with the ORACLE on a strided
    sample of 4096 codewords (not a round-trip property)."""
    from ldpc_erasure_codes_b200.codec import LdpcCodec
    S, P, sub = 64, 13, 65536
    B = 2 * sub
    codec = LdpcCodec(code=1, symbol_bytes=S, device=0, max_batch=sub)
    code = orc.Code.builtin(1)
    info = torch.empty((sub, codec.k, S), dtype=torch.uint8, device="cuda")
    rx = torch.empty((B, codec.n, S), dtype=torch.uint8, device="cuda")
    from ldpc_erasure_codes_b200.codec import fill_random
    for h in range(2):
        fill_random(info, seed=90 + h)
        codec.encode(info, out=rx[h * sub:(h + 1) * sub])
    mask = codec.gen_erasures(B, 4711, P=P, payload=rx)
    out, fail = codec.decode(rx, mask, max_iter=it, mode=mode)
    torch.cuda.synchronize()
    idx = torch.arange(5, B, B // 4096, device="cuda")[:4096]
    assert idx.numel() == 4096 and int(idx[-1]) >= sub          # the sample spans both chunks
    flags = orc.gen_erasures_iid(code.n, 4711, B, P=P)[_np(idx)]
    from ldpc_erasure_codes_b200.codec import unpack_mask
    assert np.array_equal(unpack_mask(mask[idx], code.n), flags)
    ref = orc.decode(code, _np(rx[idx]), flags, max_iter=it, mode=mode)
    assert np.array_equal(_np(fail[idx]), ref["fail_sys"])
    assert np.array_equal(_np(out[idx]), ref["out"])
    # and the whole batch through the round-trip property
    good = fail == 0
    for h in range(2):
        fill_random(info, seed=90 + h)
        g = good[h * sub:(h + 1) * sub]
        assert bool((out[h * sub:(h + 1) * sub][g] == info[g]).all())
    codec.close()


# ------------------------------------------------------------------------------- several contexts in one process
def _two_context_run(dev_a, dev_b, threaded):
    import threading
    from ldpc_erasure_codes_b200.codec import LdpcCodec, fill_random
    code = orc.Code.builtin(1)
    res = {}

    def work(dev, seed):
        with torch.cuda.device(dev):
            codec = LdpcCodec(code=1, symbol_bytes=64, device=dev, max_batch=256)
            info = torch.empty((300, codec.k, 64), dtype=torch.uint8, device=f"cuda:{dev}")
            fill_random(info, seed=seed)
            cw = codec.encode(info)
            rx = cw.clone()
            mask = codec.gen_erasures(300, seed, P=13, payload=rx)
            out_p, fail_p = codec.decode(rx, mask, max_iter=50, mode="peel")
            out_h, fail_h = codec.decode(rx, mask, max_iter=10, mode="hybrid")
            torch.cuda.synchronize(dev)
            res[(dev, seed)] = (_np(info), _np(cw), _np(rx), _np(out_p), _np(fail_p), _np(out_h), _np(fail_h))
            codec.close()

    jobs = [(dev_a, 11), (dev_b, 12)]
    if threaded:
        ts = [threading.Thread(target=work, args=j) for j in jobs]
        [t.start() for t in ts]
        [t.join() for t in ts]
    else:
        for j in jobs:
            work(*j)
    assert len(res) == 2
    for (dev, seed), (info, cw, rx, out_p, fail_p, out_h, fail_h) in res.items():
        assert np.array_equal(cw, orc.encode(code, info))
        flags = orc.gen_erasures_iid(code.n, seed, 300, P=13)
        rp = orc.decode(code, rx, flags, max_iter=50, mode="peel")
        rh = orc.decode(code, rx, flags, max_iter=10, mode="hybrid")
        assert np.array_equal(out_p, rp["out"]) and np.array_equal(fail_p, rp["fail_sys"])
        assert np.array_equal(out_h, rh["out"]) and np.array_equal(fail_h, rh["fail_sys"])


@pytest.mark.parametrize("threaded", [False, True])
def test_two_contexts_one_process(threaded):
    """Two ldpc_ctx in one process -- on devices 0 and 1 when the box has two GPUs (the shared-memory opt-in is per
    device), else both on device 0 -- from one thread and from one thread each."""
    second = 1 if torch.cuda.device_count() > 1 else 0
    _two_context_run(0, second, threaded)


def test_decode_host_multi_fans_out(codecs):
    """ldpc_decode_host_multi / ldpc_encode_host_multi: one host batch over several contexts (one per GPU where
    the box has several), one host thread per context inside the library."""
    from ldpc_erasure_codes_b200.codec import LdpcCodec, decode_host_multi, encode_host_multi
    ndev = max(1, min(torch.cuda.device_count(), 4))
    devs = list(range(ndev)) if ndev > 1 else [0, 0]
    cs = [LdpcCodec(code=1, symbol_bytes=64, device=d, max_batch=128) for d in devs]
    code = orc.Code.builtin(1)
    B = 517
    info = _rand_info(B, code.k, 64, seed=77).cpu().pin_memory()
    h_cw = encode_host_multi(cs, info)
    assert np.array_equal(h_cw.numpy(), orc.encode(code, info.numpy()))
    flags = orc.gen_erasures_iid(code.n, 4, B, P=13)
    rx = h_cw.numpy().copy()
    rx[flags == 1] = 0
    from ldpc_erasure_codes_b200.codec import pack_mask
    h_rx = torch.from_numpy(rx).pin_memory()
    h_mask = torch.from_numpy(pack_mask(flags)).pin_memory()
    for mode, it in (("peel", 50), ("hybrid", 10)):
        fail_any = torch.zeros(B, dtype=torch.uint8).pin_memory()
        out, fail = decode_host_multi(cs, h_rx, h_mask, max_iter=it, mode=mode, fail_any=fail_any)
        ref = orc.decode(code, rx, flags, max_iter=it, mode=mode)
        assert np.array_equal(out.numpy(), ref["out"]) and np.array_equal(fail.numpy(), ref["fail_sys"])
        assert np.array_equal(fail_any.numpy(), (ref["erased"].max(axis=1) > 0).astype(np.uint8))
    assert sum(c.stats()["frames"] for c in cs) == 2 * B
    from ldpc_erasure_codes_b200 import _lib as L
    with pytest.raises(L.LdpcCudaError):
        decode_host_multi([cs[0], cs[0]], h_rx, h_mask)            # a context listed twice
    for c in cs:
        c.close()


# ---- SURVEY 8(f) rank 1, the rest: variable payload length, two-buffer receiver -------------------------------------------
def test_variable_payload_length_packets(codecs):
    """num_longs_used (encoder_VITA_in_UDP_out.cl:162,186-197; receiver :94-111): per-packet payload length."""
    import torch
    from ldpc_erasure_codes_b200.codec import unpack_mask
    codec = codecs(1, 64)
    code = orc.Code.builtin(1)
    B = 3
    cw = codec.encode(_rand_info(B, codec.k, 64, seed=21))
    rng = np.random.default_rng(4)
    len8 = rng.integers(0, 9, B * codec.n).astype(np.int16)
    d_len = torch.from_numpy(len8).cuda()
    pk = codec.packetize(cw, block0=254, len8=d_len)
    assert np.array_equal(_np(pk), orc.packetize_var(_np(cw), len8, 254))
    keep = rng.permutation(pk.shape[0])[: int(0.9 * pk.shape[0])]
    sel = torch.from_numpy(keep).cuda()
    noisy = pk[sel].clone()
    cw2, mask2, counts = codec.depacketize(noisy.contiguous(), 254, B, len8=d_len[sel].contiguous())
    ref = np.zeros((B, codec.n, 64), np.uint8)
    flags = np.ones((B, codec.n), np.uint8)
    h_pk = _np(noisy)
    for i, src in enumerate(keep):
        b, sym = divmod(int(src), codec.n)
        ref[b, sym, : 8 * len8[src]] = h_pk[i, 8: 8 + 8 * len8[src]]
        flags[b, sym] = 0
    assert np.array_equal(_np(cw2), ref) and np.array_equal(unpack_mask(mask2, codec.n), flags)
    # garbage behind the valid words of a slot must not reach the symbol
    junk = pk[sel].clone()
    junk[:, 8 + 16:] = 0xAB
    short = torch.full((len(keep),), 2, dtype=torch.int16, device="cuda")
    cw3, _, _ = codec.depacketize(junk.contiguous(), 254, B, len8=short)
    assert int(_np(cw3)[:, :, 16:].max()) == 0


@pytest.mark.parametrize("mode,it,loss", [("peel", 50, 0.08), ("hybrid", 10, 0.17)])
def test_rx_stream_two_buffer_state_machine(codecs, mode, it, loss):
    """ldpc_erasure_decoder_with_reordering_logic.cl:45-142 against its plain-Python restatement: five blocks whose numbers
    wrap past 255, packets of neighbouring blocks interleaved, losses, duplicates, a foreign block, pushed in ragged batches."""
    import torch
    from ldpc_erasure_codes_b200.codec import RxStream
    codec = codecs(1, 64)
    code = orc.Code.builtin(1)
    NB, block0 = 5, 253
    info = _rand_info(NB, codec.k, 64, seed=77)
    cw = codec.encode(info)
    pk = _np(codec.packetize(cw, block0=block0))
    rng = np.random.default_rng(5)
    order = []
    for b in range(NB):                       # block b's packets, the tail of each block mixed with the head of the next
        idx = b * codec.n + rng.permutation(codec.n)
        idx = idx[rng.random(codec.n) >= loss]
        order.append(idx)
    stream = []
    for b in range(NB):
        head, tail = order[b][:-300], order[b][-300:]
        stream.extend(head.tolist())
        nxt_head = order[b + 1][:200].tolist() if b + 1 < NB else []
        mix = tail.tolist() + nxt_head
        rng.shuffle(mix)
        stream.extend(mix)
        if b + 1 < NB:
            order[b + 1] = order[b + 1][200:]
    stream = np.array(stream)
    arr = pk[stream]
    arr = np.concatenate([arr[:500], arr[100:130], arr[500:]])                 # duplicates
    foreign = pk[:40].copy()
    foreign[:, 2] = foreign[:, 6] = (block0 + 77) & 0xFF                        # a block outside the window
    arr = np.concatenate([arr[:900], foreign, arr[900:]])
    ref = orc.rx_stream(code, arr, max_iter=it, mode=mode)
    rx = RxStream(codec, max_iter=it, mode=mode, cap=4)
    got = []
    d_arr = torch.from_numpy(arr).cuda()
    pos = 0
    while pos < len(arr):
        step = int(rng.integers(1, 1500))
        got += rx.push(d_arr[pos:pos + step].contiguous())
        pos += step
    got += rx.flush()
    rx.close()
    assert [g[0] for g in got] == [r[0] for r in ref] and len(ref) >= NB
    for g, r in zip(got, ref):
        assert g[2] == r[2] and np.array_equal(_np(g[1]), r[1])
    decoded = {g[0]: _np(g[1]) for g in got if g[2] == 0}
    assert len(decoded) >= NB - 1
    for b in range(NB):
        if (block0 + b) & 0xFF in decoded:
            assert np.array_equal(decoded[(block0 + b) & 0xFF], _np(info)[b])


# ---- in-place host entry points: only the recovered symbols come back over PCIe ------------------------------------------
@pytest.mark.parametrize("ci,S,B,P,mode,it", [(1, 64, 700, 13, "peel", 50), (1, 64, 300, 13, "hybrid", 10), (0, 1024, 40, 19, "peel", 50),
                                              (1, 48, 130, 12, "peel", 3), (2, 16, 90, 14, "hybrid", 10)])
def test_decode_host_inplace_repairs_the_callers_buffer(codecs, ci, S, B, P, mode, it):
    """ldpc_decode_host_inplace: after the call the systematic rows of the pinned codeword buffer equal the decoder output of
    the oracle, the parity rows and every received symbol are untouched; chunked (max_batch < B) like the copying form."""
    from ldpc_erasure_codes_b200.codec import pack_mask
    codec = codecs(ci, S, max_batch=256)
    code = orc.Code.builtin(ci)
    info = _rand_info(B, codec.k, S, seed=31)
    cw = orc.encode(code, _np(info))
    flags = orc.gen_erasures_iid(code.n, 9, B, P=P)
    rx = cw.copy()
    rx[flags == 1] = 0
    h_rx = torch.from_numpy(rx.copy()).pin_memory()
    h_mask = torch.from_numpy(pack_mask(flags)).pin_memory()
    fail_any = torch.zeros(B, dtype=torch.uint8).pin_memory()
    fail, _ = codec.decode_host_inplace(h_rx, h_mask, max_iter=it, mode=mode, fail_any=fail_any)
    ref = orc.decode(code, rx, flags, max_iter=it, mode=mode)
    got = h_rx.numpy()
    assert np.array_equal(got[:, :code.k], ref["out"])
    assert np.array_equal(got[:, code.k:], rx[:, code.k:])
    assert np.array_equal(fail.numpy(), ref["fail_sys"])
    assert np.array_equal(fail_any.numpy(), (ref["erased"].max(axis=1) > 0).astype(np.uint8))
    # the copying form returns the same bytes
    out2, fail2 = codec.decode_host(torch.from_numpy(rx).pin_memory(), h_mask, max_iter=it, mode=mode)
    assert np.array_equal(out2.numpy(), got[:, :code.k]) and np.array_equal(fail2.numpy(), fail.numpy())


def test_host_inplace_encode_multi_and_pageable_memory(codecs):
    from ldpc_erasure_codes_b200 import _lib as L
    from ldpc_erasure_codes_b200.codec import LdpcCodec, decode_host_inplace_multi, pack_mask
    codec = codecs(1, 64, max_batch=256)
    code = orc.Code.builtin(1)
    B = 600
    info = _rand_info(B, codec.k, 64, seed=8)
    # encoder in place: information rows in, parity rows filled; pageable and pinned memory both work (plain copies)
    ref_cw = orc.encode(code, _np(info))
    for pin in (True, False):
        h = torch.zeros((B, code.n, 64), dtype=torch.uint8)
        h[:, :code.k] = info.cpu()
        h = h.pin_memory() if pin else h
        codec.encode_host_inplace(h)
        assert np.array_equal(h.numpy(), ref_cw)
    # the in-place decoder writes from the device: pageable memory is refused, nothing is touched
    flags = orc.gen_erasures_iid(code.n, 3, B, P=12)
    rx = ref_cw.copy()
    rx[flags == 1] = 0
    h_mask = torch.from_numpy(pack_mask(flags)).pin_memory()
    pageable = torch.from_numpy(rx.copy())
    with pytest.raises(L.LdpcCudaError):
        codec.decode_host_inplace(pageable, h_mask)
    assert np.array_equal(pageable.numpy(), rx)
    # several contexts on one pinned buffer
    ndev = max(1, min(torch.cuda.device_count(), 4))
    devs = list(range(ndev)) if ndev > 1 else [0, 0]
    cs = [LdpcCodec(code=1, symbol_bytes=64, device=d, max_batch=128) for d in devs]
    h_rx = torch.from_numpy(rx.copy()).pin_memory()
    fail, _ = decode_host_inplace_multi(cs, h_rx, h_mask)
    ref = orc.decode(code, rx, flags, max_iter=50)
    assert np.array_equal(h_rx.numpy()[:, :code.k], ref["out"]) and np.array_equal(fail.numpy(), ref["fail_sys"])
    assert np.array_equal(h_rx.numpy()[:, code.k:], rx[:, code.k:])
    for c in cs:
        c.close()


@pytest.mark.parametrize("env", [{"LDPC_CUDA_HOST_GATHER": "0"}, {"LDPC_CUDA_HOST_STAGES": "2"}, {"LDPC_CUDA_HOST_STAGES": "1", "LDPC_CUDA_HOST_GATHER": "4"},
                                 {"LDPC_CUDA_ENC_SPLIT_STORE": "0"}])
def test_host_path_switches_give_the_same_bytes(monkeypatch, env):
    """The A/B switches of the host path and of the encoder's store select other code paths, not other results."""
    from ldpc_erasure_codes_b200.codec import LdpcCodec, pack_mask
    for name, v in env.items():
        monkeypatch.setenv(name, v)
    codec = LdpcCodec(code=1, symbol_bytes=64, device=0, max_batch=200)      # a fresh context: the pipeline is built under these settings
    code = orc.Code.builtin(1)
    B = 900
    info = _rand_info(B, codec.k, 64, seed=4)
    ref_cw = orc.encode(code, _np(info))
    assert np.array_equal(_np(codec.encode(info)), ref_cw)
    h = torch.zeros((B, code.n, 64), dtype=torch.uint8)
    h[:, :code.k] = info.cpu()
    h = h.pin_memory()
    codec.encode_host_inplace(h)
    assert np.array_equal(h.numpy(), ref_cw)
    flags = orc.gen_erasures_iid(code.n, 6, B, P=13)
    rx = ref_cw.copy()
    rx[flags == 1] = 0
    h_rx = torch.from_numpy(rx.copy()).pin_memory()
    h_mask = torch.from_numpy(pack_mask(flags)).pin_memory()
    fail, _ = codec.decode_host_inplace(h_rx, h_mask)
    ref = orc.decode(code, rx, flags, max_iter=50)
    assert np.array_equal(h_rx.numpy()[:, :code.k], ref["out"]) and np.array_equal(h_rx.numpy()[:, code.k:], rx[:, code.k:])
    assert np.array_equal(fail.numpy(), ref["fail_sys"])
    out2, fail2 = codec.decode_host(torch.from_numpy(rx).pin_memory(), h_mask)
    assert np.array_equal(out2.numpy(), ref["out"]) and np.array_equal(fail2.numpy(), ref["fail_sys"])
    codec.close()


def test_decode_host_inplace_on_registered_memory(codecs):
    """INTEGRATION.md's route for an existing host array: cudaHostRegister it, repair it in place, unregister."""
    from ldpc_erasure_codes_b200.codec import pack_mask
    codec = codecs(1, 64, max_batch=256)
    code = orc.Code.builtin(1)
    B = 300
    info = _rand_info(B, codec.k, 64, seed=12)
    cw = orc.encode(code, _np(info))
    flags = orc.gen_erasures_iid(code.n, 21, B, P=12)
    rx = cw.copy()
    rx[flags == 1] = 0
    h_rx = torch.from_numpy(rx.copy())                    # pageable
    h_mask = torch.from_numpy(pack_mask(flags))
    rt = torch.cuda.cudart()
    assert int(rt.cudaHostRegister(h_rx.data_ptr(), h_rx.numel(), 0)) == 0
    try:
        fail, _ = codec.decode_host_inplace(h_rx, h_mask)
    finally:
        assert int(rt.cudaHostUnregister(h_rx.data_ptr())) == 0
    ref = orc.decode(code, rx, flags, max_iter=50)
    assert np.array_equal(h_rx.numpy()[:, :code.k], ref["out"]) and np.array_equal(fail.numpy(), ref["fail_sys"])
    assert np.array_equal(h_rx.numpy()[:, code.k:], rx[:, code.k:])
