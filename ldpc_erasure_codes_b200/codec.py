"""Host-side mirror of the reference's host program for the codec path.

The reference drives its kernels from one C++ translation unit (OpenCL/host/src/main.cpp):
`init_opencl()` -> `run()` (write buffer, set args, enqueue data_in / codec / data_out, finish,
read buffer) -> `cleanup()`, parameterised by the CLI flags -c (code), -p (PER numerator / 64),
-n (frames), -i (iterations).  `LdpcCodec` keeps those names and meanings on top of the C ABI
(include/ldpc_cuda.h); torch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

MODE_PEEL, MODE_HYBRID = 0, 1


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class LdpcCodec:
    """One context on one GPU (ldpc_ctx).  code = built-in index (the host's -c flag: 0 = (2000,1000),
    1 = (2040,1530); 2 = (4000,2000)) or a path to a MAT-v5 file holding `H_sparse`."""

    def __init__(self, code=1, symbol_bytes=64, device=0, max_batch=65536):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("LdpcCodec needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device)
        h = C.c_void_p()
        path = code.encode() if isinstance(code, str) else None
        ind = code if isinstance(code, int) else 0
        _lib.check(self.lib.ldpc_ctx_create(C.byref(h), path, ind, symbol_bytes, device, max_batch))
        self._h = h
        info = _lib.CodeInfo()
        _lib.check(self.lib.ldpc_ctx_info(self._h, C.byref(info)))
        self.info = info
        self.n, self.k, self.m, self.S = info.n, info.k, info.m, info.symbol_bytes
        self.mask_words = info.mask_words

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ldpc_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- code tables -------------------------------------------------------------
    def csr(self):
        rp = np.zeros(self.m + 1, np.int32)
        ci = np.zeros(self.info.nnz, np.int32)
        _lib.check(self.lib.ldpc_ctx_get_csr(self._h, rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p)))
        return rp, ci

    def set_exec_geometry(self, slice_bytes=0, slots=0):
        _lib.check(self.lib.ldpc_ctx_set_exec_geometry(self._h, slice_bytes, slots))
        _lib.check(self.lib.ldpc_ctx_info(self._h, C.byref(self.info)))

    # ---- encoder trio (encoder_top.cl data_in -> ldpc_erasure_encoder -> data_out) ----
    def encode(self, info: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        assert info.is_cuda and info.dtype == torch.uint8 and info.is_contiguous()
        B = info.shape[0]
        assert info.shape[1:] == (self.k, self.S)
        if out is None:
            out = torch.empty((B, self.n, self.S), dtype=torch.uint8, device=info.device)
        _lib.check(self.lib.ldpc_encode(self._h, _ptr(info), _ptr(out), B, _stream()))
        return out

    # ---- decoder-side data_in: Threefry erasure channel ---------------------------------
    def gen_erasures(self, B, seed, P=None, p=None, bursty=None, frame0=0, payload: torch.Tensor | None = None,
                     mask: torch.Tensor | None = None) -> torch.Tensor:
        """P = PER numerator / 64 (host flag -p); p = exact rate (32-bit threshold extension);
        bursty = (alpha, beta, bias).  Zeroes erased symbols of `payload` in place when given."""
        m = _lib.ErasureModel()
        if P is not None:
            m.model, m.per_numerator_div_64 = 0, int(P)
        elif p is not None:
            m.model, m.threshold32 = 1, min(int(p * 2 ** 32), 2 ** 32 - 1)
        else:
            m.model = 2
            m.alpha, m.beta, m.bias = bursty
        if mask is None:
            mask = torch.empty((B, self.mask_words), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.ldpc_gen_erasures(self._h, C.byref(m), seed & 0xFFFFFFFF, frame0, B, _ptr(mask),
                                              _ptr(payload), _stream()))
        return mask

    @staticmethod
    def _model(P=None, p=None, bursty=None):
        m = _lib.ErasureModel()
        if P is not None:
            m.model, m.per_numerator_div_64 = 0, int(P)
        elif p is not None:
            m.model, m.threshold32 = 1, min(int(p * 2 ** 32), 2 ** 32 - 1)
        else:
            m.model = 2
            m.alpha, m.beta, m.bias = bursty
        return m

    # ---- the reference's committed run: all-zero codewords, counters only -----------------------
    def simulate_fer(self, frames, seed, P=None, p=None, bursty=None, max_iter=50, mode="peel", frame0=0):
        """host -c <code> -p <P> -n <frames> -i <max_iter>: adds to the counters read by stats()."""
        m = self._model(P, p, bursty)
        _lib.check(self.lib.ldpc_simulate_fer(self._h, C.byref(m), seed & 0xFFFFFFFF, frame0, frames, max_iter,
                                              {"peel": MODE_PEEL, "hybrid": MODE_HYBRID}[mode], _stream()))

    # ---- ldpc_erasure_decoder(num_iter, code_ind) + data_out ----------------------------
    def decode(self, cw: torch.Tensor, mask: torch.Tensor, max_iter=50, mode="peel", out=None, fail=None, fail_any=None):
        """-> (out [B][k][S], fail [B]); pass a [B] uint8 tensor as fail_any to also get the MATLAB harness's
        criterion (any of the n symbols still unknown, LDPCErasureCodes_MessagePassingAlgSim.m:229-236)."""
        assert cw.is_cuda and cw.dtype == torch.uint8 and cw.is_contiguous() and mask.is_contiguous()
        B = cw.shape[0]
        assert cw.shape[1:] == (self.n, self.S) and mask.shape == (B, self.mask_words)
        if out is None:
            out = torch.empty((B, self.k, self.S), dtype=torch.uint8, device=cw.device)
        if fail is None:
            fail = torch.empty((B,), dtype=torch.uint8, device=cw.device)
        _lib.check(self.lib.ldpc_decode_ex(self._h, _ptr(cw), _ptr(mask), _ptr(out), _ptr(fail), _ptr(fail_any), max_iter,
                                           {"peel": MODE_PEEL, "hybrid": MODE_HYBRID}[mode], B, _stream()))
        return out, fail

    # ---- FEC packet front-ends (encoder_VITA_in_UDP_out.cl / decoder_with_reordering_logic.cl) ----
    def packetize(self, cw: torch.Tensor, block0: int = 0, len8: torch.Tensor | None = None) -> torch.Tensor:
        """cw [B][n][S] -> packets [B*n][8+S]: 64-bit FEC header word [class|block|symbol] x2, then the symbol.
        len8 [B*n] (int16, 8-byte words): variable payload length, the rest of a packet's slot is zero-filled."""
        assert cw.is_cuda and cw.dtype == torch.uint8 and cw.is_contiguous() and cw.shape[1:] == (self.n, self.S)
        B = cw.shape[0]
        pk = torch.empty((B * self.n, 8 + self.S), dtype=torch.uint8, device=cw.device)
        if len8 is None:
            _lib.check(self.lib.ldpc_packetize(self._h, _ptr(cw), block0 & 0xFFFFFFFF, B, _ptr(pk), _stream()))
        else:
            assert len8.is_cuda and len8.dtype == torch.int16 and len8.is_contiguous() and len8.numel() == B * self.n
            _lib.check(self.lib.ldpc_packetize_var(self._h, _ptr(cw), _ptr(len8), block0 & 0xFFFFFFFF, B, _ptr(pk), _stream()))
        return pk

    def depacketize(self, packets: torch.Tensor, block0: int, B: int, len8: torch.Tensor | None = None):
        """packets [N][8+S] in arrival order -> (cw [B][n][S], mask [B][mask_words], counts [B+1]).
        len8 [N] (int16): valid payload words per packet (the rest of the symbol stays zero)."""
        assert packets.is_cuda and packets.dtype == torch.uint8 and packets.is_contiguous() and packets.shape[1] == 8 + self.S
        cw = torch.empty((B, self.n, self.S), dtype=torch.uint8, device=packets.device)
        mask = torch.empty((B, self.mask_words), dtype=torch.int32, device=packets.device)
        counts = torch.empty((B + 1,), dtype=torch.int32, device=packets.device)
        if len8 is None:
            _lib.check(self.lib.ldpc_depacketize(self._h, _ptr(packets), packets.shape[0], block0 & 0xFFFFFFFF, B, _ptr(cw), _ptr(mask),
                                                 _ptr(counts), _stream()))
        else:
            assert len8.is_cuda and len8.dtype == torch.int16 and len8.is_contiguous() and len8.numel() == packets.shape[0]
            _lib.check(self.lib.ldpc_depacketize_var(self._h, _ptr(packets), _ptr(len8), packets.shape[0], block0 & 0xFFFFFFFF, B, _ptr(cw),
                                                     _ptr(mask), _ptr(counts), _stream()))
        return cw, mask, counts

    def ready_to_decode(self, cur_block_cnt: int, next_block_cnt: int) -> bool:
        return bool(self.lib.ldpc_ready_to_decode(self._h, cur_block_cnt, next_block_cnt))

    # ---- run(): host buffers in, host buffers out ---------------------------------------
    def decode_host(self, cw: torch.Tensor, mask: torch.Tensor, max_iter=50, mode="peel", out=None, fail=None):
        """cw / mask are HOST tensors (pinned for overlap); copies happen inside the call."""
        assert not cw.is_cuda and cw.dtype == torch.uint8 and cw.is_contiguous()
        B = cw.shape[0]
        if out is None:
            out = torch.empty((B, self.k, self.S), dtype=torch.uint8, pin_memory=True)
        if fail is None:
            fail = torch.empty((B,), dtype=torch.uint8, pin_memory=True)
        _lib.check(self.lib.ldpc_decode_host(self._h, _ptr(cw), _ptr(mask), _ptr(out), _ptr(fail), max_iter,
                                             {"peel": MODE_PEEL, "hybrid": MODE_HYBRID}[mode], B))
        return out, fail

    def decode_host_inplace(self, cw: torch.Tensor, mask: torch.Tensor, max_iter=50, mode="peel", fail=None, fail_any=None):
        """ldpc_decode_host_inplace: cw is a PINNED host tensor [B, n, S]; its erased systematic symbols are overwritten with
        what decode_host would have returned for them, nothing else crosses PCIe on the way back.  Returns (fail, fail_any)."""
        assert not cw.is_cuda and cw.dtype == torch.uint8 and cw.is_contiguous()
        B = cw.shape[0]
        if fail is None:
            fail = torch.empty((B,), dtype=torch.uint8, pin_memory=True)
        _lib.check(self.lib.ldpc_decode_host_inplace(self._h, _ptr(cw), _ptr(mask), _ptr(fail), _ptr(fail_any), max_iter,
                                                     {"peel": MODE_PEEL, "hybrid": MODE_HYBRID}[mode], B))
        return fail, fail_any

    def encode_host_inplace(self, cw: torch.Tensor):
        """ldpc_encode_host_inplace: cw [B, n, S] on the host with the information symbols in rows 0..k-1; the parity rows are filled in."""
        assert not cw.is_cuda and cw.dtype == torch.uint8 and cw.is_contiguous() and cw.shape[1] == self.n
        _lib.check(self.lib.ldpc_encode_host_inplace(self._h, _ptr(cw), cw.shape[0]))
        return cw

    def encode_host(self, info: torch.Tensor, out=None):
        assert not info.is_cuda and info.dtype == torch.uint8 and info.is_contiguous()
        B = info.shape[0]
        if out is None:
            out = torch.empty((B, self.n, self.S), dtype=torch.uint8, pin_memory=True)
        _lib.check(self.lib.ldpc_encode_host(self._h, _ptr(info), _ptr(out), B))
        return out

    # ---- ERROR_STAT / data_out report ----------------------------------------------------
    def stats(self):
        s = _lib.Stats()
        _lib.check(self.lib.ldpc_get_stats(self._h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in s._fields_}

    def reset_stats(self):
        _lib.check(self.lib.ldpc_reset_stats(self._h))

    # ---- CL_QUEUE_PROFILING_ENABLE / getStartEndTime ---------------------------------------
    def profile_enable(self, on=True):
        _lib.check(self.lib.ldpc_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, reset=True):
        p = _lib.Profile()
        _lib.check(self.lib.ldpc_profile_read(self._h, C.byref(p), 1 if reset else 0))
        d = {name: dict(ms=p.ms[i], launches=p.launches[i]) for i, name in enumerate(_lib.KIND_NAMES)}
        d["exec_phase_cycles"] = list(p.exec_phase_cycles)
        d["ge_phase_cycles"] = list(p.ge_phase_cycles)
        d["apply_phase_cycles"] = list(p.apply_phase_cycles)
        return d


def decode_host_multi(codecs, cw: torch.Tensor, mask: torch.Tensor, max_iter=50, mode="peel", out=None, fail=None, fail_any=None):
    """One host batch over several GPUs (ldpc_decode_host_multi): codecs = one LdpcCodec per device, same code and S;
    cw / mask are HOST tensors; frames [g*B/G, (g+1)*B/G) go to GPU g, one host thread per GPU inside the library."""
    c0 = codecs[0]
    assert not cw.is_cuda and cw.dtype == torch.uint8 and cw.is_contiguous()
    B = cw.shape[0]
    if out is None:
        out = torch.empty((B, c0.k, c0.S), dtype=torch.uint8, pin_memory=True)
    if fail is None:
        fail = torch.empty((B,), dtype=torch.uint8, pin_memory=True)
    arr = (C.c_void_p * len(codecs))(*[c._h for c in codecs])
    _lib.check(c0.lib.ldpc_decode_host_multi(arr, len(codecs), _ptr(cw), _ptr(mask), _ptr(out), _ptr(fail), _ptr(fail_any), max_iter,
                                             {"peel": MODE_PEEL, "hybrid": MODE_HYBRID}[mode], B))
    return out, fail


def decode_host_inplace_multi(codecs, cw: torch.Tensor, mask: torch.Tensor, max_iter=50, mode="peel", fail=None, fail_any=None):
    """ldpc_decode_host_inplace_multi: the in-place form of decode_host_multi (cw pinned, repaired where it lies)."""
    c0 = codecs[0]
    assert not cw.is_cuda and cw.dtype == torch.uint8 and cw.is_contiguous()
    B = cw.shape[0]
    if fail is None:
        fail = torch.empty((B,), dtype=torch.uint8, pin_memory=True)
    arr = (C.c_void_p * len(codecs))(*[c._h for c in codecs])
    _lib.check(c0.lib.ldpc_decode_host_inplace_multi(arr, len(codecs), _ptr(cw), _ptr(mask), _ptr(fail), _ptr(fail_any), max_iter,
                                                     {"peel": MODE_PEEL, "hybrid": MODE_HYBRID}[mode], B))
    return fail, fail_any


def encode_host_multi(codecs, info: torch.Tensor, out=None):
    c0 = codecs[0]
    assert not info.is_cuda and info.dtype == torch.uint8 and info.is_contiguous()
    B = info.shape[0]
    if out is None:
        out = torch.empty((B, c0.n, c0.S), dtype=torch.uint8, pin_memory=True)
    arr = (C.c_void_p * len(codecs))(*[c._h for c in codecs])
    _lib.check(c0.lib.ldpc_encode_host_multi(arr, len(codecs), _ptr(info), _ptr(out), B))
    return out


def read_h_file(path: str):
    """Parse a MAT-v5 code file with the library's own C++ loader (no GPU needed) -> (m, n, triangular, row_ptr, col_idx)."""
    lib = _lib.load()
    dims = (C.c_int32 * 4)()
    _lib.check(lib.ldpc_read_h_file(path.encode(), C.byref(dims), None, None))
    rp = np.zeros(dims[0] + 1, np.int32)
    ci = np.zeros(dims[2], np.int32)
    _lib.check(lib.ldpc_read_h_file(path.encode(), C.byref(dims), rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p)))
    return dims[0], dims[1], bool(dims[3]), rp, ci


def fill_random(t: torch.Tensor, seed: int, block0: int = 0):
    """Counter-based synthetic payload (Threefry key {2, seed}); t must be a CUDA byte tensor."""
    lib = _lib.load()
    _lib.check(lib.ldpc_fill_random(_ptr(t), t.numel() * t.element_size(), seed & 0xFFFFFFFF, block0,
                                    t.device.index or 0, _stream()))
    return t


def unpack_mask(mask, n):
    """[B][mask_words] int32 bit mask -> [B][n] uint8 flags (numpy)."""
    m = np.ascontiguousarray(mask.cpu().numpy() if hasattr(mask, "cpu") else mask).view(np.uint32)
    bits = np.unpackbits(m.view(np.uint8), axis=1, bitorder="little")
    return bits[:, :n].copy()


def pack_mask(flags):
    """[B][n] uint8 flags -> [B][ceil(n/32)] int32 bit mask (numpy)."""
    B, n = flags.shape
    nw = (n + 31) // 32
    pad = np.zeros((B, nw * 32), np.uint8)
    pad[:, :n] = flags
    return np.packbits(pad, axis=1, bitorder="little").view(np.uint32).astype(np.int32, copy=False).reshape(B, nw)


class RxStream:
    """The receiver's two-buffer state machine (ldpc_rx_stream; ldpc_erasure_decoder_with_reordering_logic.cl:45-142):
    push packets in arrival order, get the blocks that became ready -- [(block number, out [k][S], fail)]."""

    def __init__(self, codec: LdpcCodec, max_iter=50, mode="peel", cap=8):
        self.lib = _lib.load()
        self.codec = codec
        self.cap = cap
        h = C.c_void_p()
        _lib.check(self.lib.ldpc_rx_stream_create(C.byref(h), codec._h, max_iter, {"peel": MODE_PEEL, "hybrid": MODE_HYBRID}[mode]))
        self._h = h
        self._out = torch.empty((cap, codec.k, codec.S), dtype=torch.uint8, device=codec.device)
        self._fail = torch.empty((cap,), dtype=torch.uint8, device=codec.device)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ldpc_rx_stream_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _collect(self, nd, blocks):
        torch.cuda.synchronize()
        return [(int(blocks[i]), self._out[i].clone(), int(self._fail[i].item())) for i in range(nd)]

    def push(self, packets: torch.Tensor, len8: torch.Tensor | None = None):
        c = self.codec
        assert packets.is_cuda and packets.dtype == torch.uint8 and packets.is_contiguous() and packets.shape[1] == 8 + c.S
        blocks = (C.c_int32 * self.cap)()
        nd = C.c_int(0)
        _lib.check(self.lib.ldpc_rx_stream_push(self._h, _ptr(packets), _ptr(len8), packets.shape[0], _ptr(self._out), _ptr(self._fail),
                                                blocks, self.cap, C.byref(nd), _stream()))
        return self._collect(nd.value, blocks)

    def flush(self):
        blocks = (C.c_int32 * self.cap)()
        nd = C.c_int(0)
        _lib.check(self.lib.ldpc_rx_stream_flush(self._h, _ptr(self._out), _ptr(self._fail), blocks, self.cap, C.byref(nd), _stream()))
        return self._collect(nd.value, blocks)

    def state(self):
        st = (C.c_int32 * 4)()
        _lib.check(self.lib.ldpc_rx_stream_state(self._h, C.byref(st)))
        return dict(cur=st[0], next=st[1], cur_cnt=st[2], next_cnt=st[3])


class NbLdpcCodec:
    """Non-binary GF(256) LDPC code over the structure of a binary LdpcCodec (ldpc_nb_ctx; SURVEY 8(f) rank 3:
    Matlab/ErasureCodes_NonBinaryLDPCSim.m, My_LDPC_HybridML_NonBinary_Erasure_Decoder.m).  coef = one nonzero field
    element per nonzero of H in CSR order (numpy u8), or None to draw them from `seed` as the simulation does."""

    def __init__(self, base: LdpcCodec, coef=None, seed=1):
        import numpy as np
        self.lib = _lib.load()
        self.base = base
        h = C.c_void_p()
        cptr = None
        if coef is not None:
            coef = np.ascontiguousarray(coef, dtype=np.uint8)
            assert coef.size == base.info.nnz
            cptr = coef.ctypes.data_as(C.c_void_p)
        _lib.check(self.lib.ldpc_nb_ctx_create(C.byref(h), base._h, cptr, seed & 0xFFFFFFFF))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ldpc_nb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def coefficients(self):
        import numpy as np
        out = np.zeros(self.base.info.nnz, dtype=np.uint8)
        _lib.check(self.lib.ldpc_nb_get_coefficients(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def encode(self, info: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        b = self.base
        B = info.shape[0]
        assert info.dtype == torch.uint8 and info.is_contiguous() and tuple(info.shape[1:]) == (b.k, b.S)
        if out is None:
            out = torch.empty((B, b.n, b.S), dtype=torch.uint8, device=b.device)
        _lib.check(self.lib.ldpc_nb_encode(self._h, _ptr(info), _ptr(out), B, _stream()))
        return out

    def decode(self, cw: torch.Tensor, mask: torch.Tensor, max_iter=10, mode="hybrid", out=None, fail=None):
        b = self.base
        B = cw.shape[0]
        assert cw.dtype == torch.uint8 and cw.is_contiguous() and tuple(cw.shape[1:]) == (b.n, b.S)
        assert mask.is_contiguous() and tuple(mask.shape) == (B, b.mask_words)
        if out is None:
            out = torch.empty((B, b.k, b.S), dtype=torch.uint8, device=b.device)
        if fail is None:
            fail = torch.empty((B,), dtype=torch.uint8, device=b.device)
        _lib.check(self.lib.ldpc_nb_decode(self._h, _ptr(cw), _ptr(mask), _ptr(out), _ptr(fail), max_iter,
                                           {"peel": MODE_PEEL, "hybrid": MODE_HYBRID}[mode], B, _stream()))
        return out, fail


class RsCodec:
    """Reed-Solomon GF(2^8) erasure codec (rs_ctx): field 0x171, G[i][j] = alpha^(i*j) systematised,
    decode from the first k received symbols -- Matlab/ReedSolomonErasureCodes.m, Test_My_RS_Decode.m."""

    def __init__(self, n=255, k=191, symbol_bytes=1024, device=0, max_batch=4096):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("RsCodec needs a CUDA device (sm_100a); there is no CPU fallback")
        h = C.c_void_p()
        _lib.check(self.lib.rs_ctx_create(C.byref(h), n, k, symbol_bytes, device, max_batch))
        self._h = h
        self.n, self.k, self.S = n, k, symbol_bytes
        self.mask_words = (n + 31) // 32
        self.device = torch.device("cuda", device)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.rs_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def generator(self):
        g = np.zeros((self.k, self.n), np.uint8)
        _lib.check(self.lib.rs_ctx_get_generator(self._h, g.ctypes.data_as(C.c_void_p)))
        return g

    def encode(self, info: torch.Tensor, out=None):
        assert info.is_cuda and info.dtype == torch.uint8 and info.is_contiguous()
        B = info.shape[0]
        assert info.shape[1:] == (self.k, self.S)
        if out is None:
            out = torch.empty((B, self.n, self.S), dtype=torch.uint8, device=info.device)
        _lib.check(self.lib.rs_encode(self._h, _ptr(info), _ptr(out), B, _stream()))
        return out

    def decode(self, cw: torch.Tensor, mask: torch.Tensor, out=None, fail=None):
        assert cw.is_cuda and cw.dtype == torch.uint8 and cw.is_contiguous() and mask.is_contiguous()
        B = cw.shape[0]
        assert cw.shape[1:] == (self.n, self.S) and mask.shape == (B, self.mask_words)
        if out is None:
            out = torch.empty((B, self.k, self.S), dtype=torch.uint8, device=cw.device)
        if fail is None:
            fail = torch.empty((B,), dtype=torch.uint8, device=cw.device)
        _lib.check(self.lib.rs_decode(self._h, _ptr(cw), _ptr(mask), _ptr(out), _ptr(fail), B, _stream()))
        return out, fail
