"""ctypes binding of libldpc_cuda.so (include/ldpc_cuda.h).  No CPU fallback: if the CUDA
library is missing or a call fails, this module raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libldpc_cuda.so")

# every symbol include/ldpc_cuda.h declares (tests check the library exports all of them)
EXPORTS = [
    "ldpc_ctx_create", "ldpc_ctx_destroy", "ldpc_ctx_info", "ldpc_read_h_file", "ldpc_ctx_get_csr", "ldpc_ctx_set_exec_geometry",
    "ldpc_encode", "ldpc_gen_erasures", "ldpc_decode", "ldpc_decode_ex", "ldpc_simulate_fer", "ldpc_get_stats", "ldpc_reset_stats",
    "ldpc_encode_host", "ldpc_decode_host", "ldpc_decode_host_ex", "ldpc_encode_host_multi", "ldpc_decode_host_multi", "ldpc_decode_host_inplace", "ldpc_decode_host_inplace_multi", "ldpc_encode_host_inplace", "ldpc_fill_random", "ldpc_packetize", "ldpc_depacketize", "ldpc_ready_to_decode", "ldpc_packetize_var", "ldpc_depacketize_var", "ldpc_rx_stream_create", "ldpc_rx_stream_destroy", "ldpc_rx_stream_push", "ldpc_rx_stream_flush", "ldpc_rx_stream_state", "ldpc_profile_enable", "ldpc_profile_read",
    "rs_ctx_create", "rs_ctx_destroy", "rs_ctx_get_generator", "rs_encode", "rs_decode",
    "ldpc_nb_ctx_create", "ldpc_nb_ctx_destroy", "ldpc_nb_get_coefficients", "ldpc_nb_encode", "ldpc_nb_decode",
    "ldpc_h_generate", "ldpc_h_count_short_cycles", "ldpc_h_last_error_string",
    "ldpc_last_error_string", "ldpc_cuda_abi_version",
]


class LdpcCudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libldpc_cuda error {code}: {msg}")
        self.code = code


class CodeInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("n", "k", "m", "nnz", "symbol_bytes", "mask_words", "rs_n", "rs_k", "max_row_weight",
                 "max_col_weight", "encode_levels", "slice_bytes", "exec_slots", "device")] + [("max_batch", C.c_int64)]


class ErasureModel(C.Structure):
    _fields_ = [("model", C.c_int32), ("per_numerator_div_64", C.c_int32), ("threshold32", C.c_uint32),
                ("alpha", C.c_double), ("beta", C.c_double), ("bias", C.c_double)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("frames", "ldpc_errors", "rs_errors", "ml_attempts", "ml_failures", "ml_recovered", "any_errors")]


K_KINDS = 8
KIND_NAMES = ["peel", "exec_decode", "exec_encode", "hybrid", "channel", "hybrid_warp", "hybrid_cta", "hybrid_apply"]


class Profile(C.Structure):
    _fields_ = [("ms", C.c_double * K_KINDS), ("launches", C.c_int64 * K_KINDS), ("exec_phase_cycles", C.c_uint64 * 8), ("ge_phase_cycles", C.c_uint64 * 8), ("apply_phase_cycles", C.c_uint64 * 8)]


_lib = None


def load():
    """Loads libldpc_cuda.so; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} not found: build it with `python -m ldpc_erasure_codes_b200.build` "
                          "(the codec has no CPU fallback)")
    lib = C.CDLL(SO_PATH)
    vp, i32, i64, u32, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_uint64
    lib.ldpc_ctx_create.argtypes = [C.POINTER(vp), C.c_char_p, i32, i32, i32, i64]
    lib.ldpc_ctx_destroy.argtypes = [vp]
    lib.ldpc_ctx_info.argtypes = [vp, C.POINTER(CodeInfo)]
    lib.ldpc_read_h_file.argtypes = [C.c_char_p, C.POINTER(C.c_int32 * 4), vp, vp]
    lib.ldpc_ctx_get_csr.argtypes = [vp, vp, vp]
    lib.ldpc_ctx_set_exec_geometry.argtypes = [vp, i32, i32]
    lib.ldpc_encode.argtypes = [vp, vp, vp, i64, vp]
    lib.ldpc_gen_erasures.argtypes = [vp, C.POINTER(ErasureModel), u32, u64, i64, vp, vp, vp]
    lib.ldpc_decode.argtypes = [vp, vp, vp, vp, vp, i32, i32, i64, vp]
    lib.ldpc_decode_ex.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i64, vp]
    lib.ldpc_simulate_fer.argtypes = [vp, C.POINTER(ErasureModel), u32, u64, i64, i32, i32, vp]
    lib.ldpc_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.ldpc_reset_stats.argtypes = [vp]
    lib.ldpc_encode_host.argtypes = [vp, vp, vp, i64]
    lib.ldpc_decode_host.argtypes = [vp, vp, vp, vp, vp, i32, i32, i64]
    lib.ldpc_decode_host_ex.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i64]
    lib.ldpc_encode_host_multi.argtypes = [C.POINTER(vp), i32, vp, vp, i64]
    lib.ldpc_decode_host_multi.argtypes = [C.POINTER(vp), i32, vp, vp, vp, vp, vp, i32, i32, i64]
    lib.ldpc_decode_host_inplace.argtypes = [vp, vp, vp, vp, vp, i32, i32, i64]
    lib.ldpc_decode_host_inplace_multi.argtypes = [C.POINTER(vp), i32, vp, vp, vp, vp, i32, i32, i64]
    lib.ldpc_encode_host_inplace.argtypes = [vp, vp, i64]
    lib.ldpc_fill_random.argtypes = [vp, i64, u32, u64, i32, vp]
    lib.ldpc_packetize.argtypes = [vp, vp, u32, i64, vp, vp]
    lib.ldpc_depacketize.argtypes = [vp, vp, i64, u32, i64, vp, vp, vp, vp]
    lib.ldpc_ready_to_decode.argtypes = [vp, i32, i32]
    lib.ldpc_packetize_var.argtypes = [vp, vp, vp, u32, i64, vp, vp]
    lib.ldpc_depacketize_var.argtypes = [vp, vp, vp, i64, u32, i64, vp, vp, vp, vp]
    lib.ldpc_rx_stream_create.argtypes = [C.POINTER(vp), vp, i32, i32]
    lib.ldpc_rx_stream_destroy.argtypes = [vp]
    lib.ldpc_rx_stream_push.argtypes = [vp, vp, vp, i64, vp, vp, vp, i32, C.POINTER(i32), vp]
    lib.ldpc_rx_stream_flush.argtypes = [vp, vp, vp, vp, i32, C.POINTER(i32), vp]
    lib.ldpc_rx_stream_state.argtypes = [vp, C.POINTER(C.c_int32 * 4)]
    lib.ldpc_profile_enable.argtypes = [vp, i32]
    lib.ldpc_profile_read.argtypes = [vp, C.POINTER(Profile), i32]
    lib.rs_ctx_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32, i64]
    lib.rs_ctx_destroy.argtypes = [vp]
    lib.rs_ctx_get_generator.argtypes = [vp, vp]
    lib.rs_encode.argtypes = [vp, vp, vp, i64, vp]
    lib.rs_decode.argtypes = [vp, vp, vp, vp, vp, i64, vp]
    lib.ldpc_nb_ctx_create.argtypes = [C.POINTER(vp), vp, vp, u32]
    lib.ldpc_nb_ctx_destroy.argtypes = [vp]
    lib.ldpc_nb_get_coefficients.argtypes = [vp, vp]
    lib.ldpc_nb_encode.argtypes = [vp, vp, vp, i64, vp]
    lib.ldpc_nb_decode.argtypes = [vp, vp, vp, vp, vp, i32, i32, i64, vp]
    lib.ldpc_h_generate.argtypes = [vp, i32, vp, i32, u64, i32, C.POINTER(C.c_int32 * 4), vp, vp, i64, vp]
    lib.ldpc_h_count_short_cycles.argtypes = [vp, vp, i32, i32, C.POINTER(i64), C.POINTER(i64)]
    lib.ldpc_h_last_error_string.restype = C.c_char_p
    lib.ldpc_last_error_string.restype = C.c_char_p
    lib.ldpc_cuda_abi_version.restype = i32
    for name in EXPORTS:  # every entry point returns an int status except the error string
        fn = getattr(lib, name)
        if name not in ("ldpc_last_error_string", "ldpc_h_last_error_string"):
            fn.restype = i32
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise LdpcCudaError(rc, load().ldpc_last_error_string().decode())
