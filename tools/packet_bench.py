#!/usr/bin/env python3
"""Device time of the FEC packet front-end kernels (packetize / depacketize), n2040/k1530, S=64, 256-block windows."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldpc_erasure_codes_b200.codec import LdpcCodec, fill_random
S = int(os.environ.get("S", "64")); B = 256; W = int(os.environ.get("WINDOWS", "64"))
codec = LdpcCodec(code=1, symbol_bytes=S, device=0, max_batch=B)
n = codec.n
cws = [torch.empty((B, n, S), dtype=torch.uint8, device="cuda") for _ in range(W)]
for i, t in enumerate(cws): fill_random(t, 3 + i)
pks = [codec.packetize(t, 0) for t in cws]
perm = torch.randperm(B * n, device="cuda")
shuf = [p[perm].contiguous() for p in pks]
torch.cuda.synchronize()
def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
t_pk = timed(lambda: [codec.packetize(t, 0) for t in cws])
t_de = timed(lambda: [codec.depacketize(p, 0, B) for p in shuf])
rx, mask, cnt = codec.depacketize(shuf[0], 0, B)
assert torch.equal(rx, cws[0]) and int(mask.ne(0).sum()) == 0
npk = W * B * n
byt_pk = npk * (S + 8 + S)                   # read the symbol, write header + symbol
byt_de = npk * (8 + S + S) + W * B * n * S   # read the packet, write the symbol (+ the zero fill of the block buffers)
print(json.dumps(dict(op="packetize", packets=npk, S=S, ms=round(t_pk, 3), GBs=round(byt_pk / t_pk / 1e6, 1), Mpackets_s=round(npk / t_pk / 1e3, 1))))
print(json.dumps(dict(op="depacketize (random arrival order)", packets=npk, S=S, ms=round(t_de, 3), GBs=round(byt_de / t_de / 1e6, 1), Mpackets_s=round(npk / t_de / 1e3, 1))))
