#!/usr/bin/env python3
"""Re-derives the reference paper's block-error-rate table (Latex/Milcom_2022_ErasureCodes.tex:197-210) with the
B200 codec: all-zero codewords, Threefry erasures at P/64, peeling decoder with the host default of 50 sweeps, the
RS-equivalent MDS count -- the flow of the reference's committed host program, via ldpc_simulate_fer.  A statistical
parity check against numbers the reference itself published (different RNG seeds, so agreement is within sampling
error).  Also reports the hybrid-ML decoder, which the paper only ran in MATLAB."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ldpc_erasure_codes_b200.codec import LdpcCodec

PAPER = {  # (code, P): (LDPC BLER, RS BLER, N_T)
    (0, 24): (2.2e-5, 2.1e-5, 2e6), (0, 23): (0.0, 2e-6, 1e7), (0, 22): (0.0, 0.0, 2e8),
    (1, 12): (0.02, 7.3e-3, 1e6), (1, 11): (1.3e-4, 9.3e-4, 1e6), (1, 10): (0.0, 6.3e-5, 1e7), (1, 9): (0.0, 2e-6, 1e8),
}

frames = int(float(os.environ.get("FRAMES", "2e6")))
hybrid_frames = int(float(os.environ.get("HYBRID_FRAMES", "2e5")))
for (ci, P), (bler, rs_bler, nt) in PAPER.items():
    codec = LdpcCodec(code=ci, symbol_bytes=16, device=0, max_batch=1 << 16)
    mult = codec.n // codec.info.rs_n
    codec.reset_stats()
    t0 = time.perf_counter()
    codec.simulate_fer(frames, seed=20221100 + P, P=P, max_iter=50, mode="peel")
    st = codec.stats()
    dt = time.perf_counter() - t0
    row = dict(code=f"({codec.n},{codec.k})", per=f"{P}/64={P/64:.4f}", frames=st["frames"],
               ldpc_bler=st["ldpc_errors"] / st["frames"], paper_ldpc_bler=bler,
               rs_bler=st["rs_errors"] / (mult * st["frames"]), paper_rs_bler=rs_bler, paper_frames=nt,
               seconds=round(dt, 2), frames_per_s=round(st["frames"] / dt))
    codec.reset_stats()
    codec.simulate_fer(hybrid_frames, seed=20221100 + P, P=P, max_iter=10, mode="hybrid")
    sh = codec.stats()
    row.update(hybrid_frames=sh["frames"], hybrid_bler=(sh["ldpc_errors"] - sh["ml_recovered"]) / sh["frames"],
               hybrid_ml_attempt_rate=sh["ml_attempts"] / sh["frames"])
    print(json.dumps(row), flush=True)
    codec.close()
