#!/usr/bin/env python3
"""The points of the paper's BLER table where it observed no error, at 1e9 frames each (peeling, 50 sweeps).

The reference's erasure generator counts symbols in a 32-bit counter that wraps (decoder_top.cl:75,96), so for one seed the
frame sequence repeats every 2^32 / gcd(2^32, n) frames (2^28 for n = 2000, 2^29 for n = 2040).  A long run is therefore cut
into segments of at most one period, each with its own seed: every frame counted is a distinct draw."""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldpc_erasure_codes_b200.codec import LdpcCodec
frames = int(float(os.environ.get("FRAMES", "1e9")))
for ci, P in ((1, 10), (1, 9), (0, 23), (0, 22)):
    codec = LdpcCodec(code=ci, symbol_bytes=16, device=0, max_batch=1 << 16)
    mult = codec.n // codec.info.rs_n
    period = (1 << 32) // math.gcd(1 << 32, codec.n)
    codec.reset_stats()
    t0 = time.perf_counter()
    done, seg = 0, 0
    while done < frames:
        nf = min(period, frames - done)
        codec.simulate_fer(nf, seed=424200 + 1000 * P + seg, P=P, max_iter=50, mode="peel")
        done += nf
        seg += 1
    st = codec.stats()
    dt = time.perf_counter() - t0
    print(json.dumps(dict(code=f"({codec.n},{codec.k})", per=f"{P}/64", frames=st["frames"], segments=seg, period=period,
                          ldpc_errors=st["ldpc_errors"], ldpc_bler=st["ldpc_errors"] / st["frames"],
                          rs_bler=st["rs_errors"] / (mult * st["frames"]), seconds=round(dt, 1))), flush=True)
    codec.close()
