#!/usr/bin/env python3
"""Measurement of the non-binary GF(256) LDPC code on one B200: encode, sweeps-only and hybrid decode of
B codewords (device resident, CUDA events), with the oracle's restatement of the MATLAB decoder on the host cores beside it.
One JSON line per operation."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ldpc_erasure_codes_b200.codec import LdpcCodec, NbLdpcCodec, fill_random
from oracle import oracle as orc

ci = int(os.environ.get("CODE", "1")); S = int(os.environ.get("S", "64")); B = int(os.environ.get("B", "8192")); P = int(os.environ.get("P", "13"))
base = LdpcCodec(code=ci, symbol_bytes=S, device=0, max_batch=B)
nb = NbLdpcCodec(base, seed=1)
n, k = base.n, base.k
info = torch.empty((B, k, S), dtype=torch.uint8, device="cuda"); fill_random(info, 3)
cw = nb.encode(info)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


peak = 6554.2
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
ms = timed(lambda: nb.encode(info, out=cw))
print(json.dumps(dict(op="nb_encode", code=ci, S=S, codewords=B, ms=round(ms, 3), info_gbps=round(B * k * S * 8 / ms / 1e6, 1),
                      hbm_frac=round((k + n) * S * B / ms / 1e6 / peak, 3))), flush=True)
rx = cw.clone()
mask = base.gen_erasures(B, 12345, P=P, payload=rx)
out = torch.empty_like(info); fail = torch.empty(B, dtype=torch.uint8, device="cuda")
for mode, it in (("peel", 50), ("hybrid", 10)):
    ms = timed(lambda: nb.decode(rx, mask, max_iter=it, mode=mode, out=out, fail=fail))
    fer = float(fail.float().mean())
    ok = fail == 0
    assert bool((out[ok] == info[ok]).all())
    print(json.dumps(dict(op=f"nb_decode_{mode}", code=ci, S=S, P=P, max_iter=it, codewords=B, ms=round(ms, 3), info_gbps=round(B * k * S * 8 / ms / 1e6, 1),
                          hbm_frac=round((n * S + k * S + n // 8 + 1) * B / ms / 1e6 / peak, 3), frame_error_rate=round(fer, 5))), flush=True)
# CPU restatement (oracle), all host threads, bounded sample
code = orc.Code.builtin(ci); coef = nb.coefficients()
Bc = 256
h_rx = rx[:Bc].cpu().numpy(); flags = orc.gen_erasures_iid(n, 12345, Bc, P=P)
t0 = time.perf_counter(); orc.nb_decode(code, coef, h_rx, flags, max_iter=10, mode="hybrid"); dt = time.perf_counter() - t0
print(json.dumps(dict(op="cpu_nb_decode_hybrid", kind="port (oracle/ldpc_oracle.c, restating the MATLAB decoder)", codewords=Bc, threads=orc.num_threads(),
                      info_gbps=round(Bc * k * S * 8 / dt / 1e9, 3))), flush=True)
