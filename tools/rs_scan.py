#!/usr/bin/env python3
"""RS(255,191), S=1024: device time of rs_decode against the erasure probability (how the time splits between the
pattern part, which grows with t^2..t^3, and the payload part, which grows with t), and of rs_encode."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldpc_erasure_codes_b200.codec import RsCodec, fill_random
n, k, S = 255, int(os.environ.get("K", "191")), int(os.environ.get("S", "1024")); B = int(os.environ.get("B", "1776"))
codec = RsCodec(n=n, k=k, symbol_bytes=S, device=0, max_batch=B)
info = torch.empty((B, k, S), dtype=torch.uint8, device="cuda"); fill_random(info, 3)
cw = codec.encode(info)
def timed(fn, reps=3):
    fn(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
print(json.dumps(dict(op="encode", B=B, ms=round(timed(lambda: codec.encode(info, out=cw)), 3))), flush=True)
for p in [float(x) for x in os.environ.get("PS", "0.0,0.02,0.05,0.1,0.2,0.25").split(",")]:
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    mask = (torch.rand((B, n), device="cuda", generator=g) < p)
    words = torch.zeros((B, codec.mask_words * 32), dtype=torch.bool, device="cuda"); words[:, :n] = mask
    packed = (words.view(B, codec.mask_words, 32).to(torch.int64) << torch.arange(32, device="cuda")).sum(-1)
    packed = packed.to(torch.int64).where(packed < 2 ** 31, packed - 2 ** 32).to(torch.int32).contiguous()
    rx = cw.clone(); rx[mask] = 0
    out, fail = codec.decode(rx, packed)
    good = fail == 0
    assert bool((out[good] == info[good]).all())
    ms = timed(lambda: codec.decode(rx, packed, out=out, fail=fail))
    print(json.dumps(dict(op="decode", p=p, B=B, t_avg=round(float(mask[:, :k].sum(1).float().mean()), 1), fer=round(float(fail.float().mean()), 4), ms=round(ms, 3))), flush=True)
