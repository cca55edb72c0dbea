#!/usr/bin/env python3
"""Kernel-time sweep on a GPU box: per-launch device time of the peel compiler and the payload executor for
several erasure rates and executor geometries (slice width W, slots)."""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldpc_erasure_codes_b200.codec import LdpcCodec, fill_random

ci = int(os.environ.get("CODE", "1")); S = int(os.environ.get("S", "64")); B = int(os.environ.get("B", "32768"))
geoms = [tuple(map(int, g.split("x"))) for g in os.environ.get("GEOMS", "32x3,32x2,32x1,64x1,16x4,16x2").split(",")]
rates = [int(x) for x in os.environ.get("RATES", "0,6,10,13").split(",")]
codec = LdpcCodec(code=ci, symbol_bytes=S, device=0, max_batch=B)
n, k = codec.n, codec.k
info = torch.empty((B, k, S), dtype=torch.uint8, device="cuda"); fill_random(info, 1)
cw = codec.encode(info)
alg = (n * S + (n + 7) // 8 + k * S + 1) * B
res = []
for P in rates:
    rx = cw.clone()
    mask = codec.gen_erasures(B, 12345, P=P, payload=rx)
    out = torch.empty((B, k, S), dtype=torch.uint8, device="cuda"); fail = torch.empty((B,), dtype=torch.uint8, device="cuda")
    for (W, slots) in geoms:
        try:
            codec.set_exec_geometry(W, slots)
        except Exception as e:
            print("skip", W, slots, e); continue
        for _ in range(2): codec.decode(rx, mask, out=out, fail=fail)
        codec.profile_read(reset=True); codec.profile_enable(True)
        for _ in range(3): codec.decode(rx, mask, out=out, fail=fail)
        pr = codec.profile_read(reset=True); codec.profile_enable(False)
        ex = pr["exec_decode"]["ms"] / pr["exec_decode"]["launches"]; pe = pr["peel"]["ms"] / pr["peel"]["launches"]
        r = dict(P=P, W=codec.info.slice_bytes, slots=codec.info.exec_slots, exec_ms=round(ex, 3), peel_ms=round(pe, 3),
                 exec_GBs=round(alg / ex / 1e6, 1), fer=round(float(fail.float().mean()), 4))
        ph = pr.get("exec_phase_cycles", [0] * 8)
        if ph[4]:
            r["phase_cyc_per_unit"] = dict(claim=ph[0] // ph[4], load=ph[1] // ph[4], bulk=ph[5] // ph[4], walk=ph[2] // ph[4], store=ph[3] // ph[4])
        print(json.dumps(r), flush=True); res.append(r)
# encode
for (W, slots) in geoms:
    try: codec.set_exec_geometry(W, slots)
    except Exception: continue
    o = torch.empty((B, n, S), dtype=torch.uint8, device="cuda")
    for _ in range(2): codec.encode(info, out=o)
    codec.profile_read(reset=True); codec.profile_enable(True)
    for _ in range(3): codec.encode(info, out=o)
    pr = codec.profile_read(reset=True); codec.profile_enable(False)
    ex = pr["exec_encode"]["ms"] / pr["exec_encode"]["launches"]
    print(json.dumps(dict(op="encode", W=codec.info.slice_bytes, slots=codec.info.exec_slots, ms=round(ex, 3),
                          GBs=round((k * S + n * S) * B / ex / 1e6, 1))), flush=True)
