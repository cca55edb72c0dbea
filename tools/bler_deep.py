#!/usr/bin/env python3
"""The points of the paper's BLER table where it observed no error, at 1e9 frames each (peeling, 50 sweeps)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldpc_erasure_codes_b200.codec import LdpcCodec
frames = int(float(os.environ.get("FRAMES", "1e9")))
for ci, P in ((1, 10), (1, 9), (0, 23), (0, 22)):
    codec = LdpcCodec(code=ci, symbol_bytes=16, device=0, max_batch=1 << 16)
    mult = codec.n // codec.info.rs_n
    codec.reset_stats()
    t0 = time.perf_counter()
    codec.simulate_fer(frames, seed=424200 + P, P=P, max_iter=50, mode="peel")
    st = codec.stats()
    dt = time.perf_counter() - t0
    print(json.dumps(dict(code=f"({codec.n},{codec.k})", per=f"{P}/64", frames=st["frames"], ldpc_errors=st["ldpc_errors"],
                          ldpc_bler=st["ldpc_errors"] / st["frames"], rs_bler=st["rs_errors"] / (mult * st["frames"]),
                          seconds=round(dt, 1))), flush=True)
    codec.close()
