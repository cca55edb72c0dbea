import os, sys, json
sys.path.insert(0, "/root/repo")
import torch
from ldpc_erasure_codes_b200.codec import LdpcCodec, fill_random
S=64
codec = LdpcCodec(code=1, symbol_bytes=S, device=0, max_batch=4096)
n,k=codec.n,codec.k
for B in (1, 4, 8, 64, 512):
    info = torch.empty((B, k, S), dtype=torch.uint8, device="cuda"); fill_random(info, 1)
    cw = codec.encode(info); rx = cw.clone(); mask = codec.gen_erasures(B, 99, P=13, payload=rx)
    out = torch.empty((B, k, S), dtype=torch.uint8, device="cuda"); fail = torch.empty((B,), dtype=torch.uint8, device="cuda")
    for _ in range(3): codec.decode(rx, mask, out=out, fail=fail)
    codec.profile_read(reset=True); codec.profile_enable(True)
    for _ in range(20): codec.decode(rx, mask, out=out, fail=fail)
    pr = codec.profile_read(reset=True); codec.profile_enable(False)
    print(B, "peel us", round(pr["peel"]["ms"]/pr["peel"]["launches"]*1e3,1), "exec us", round(pr["exec_decode"]["ms"]/pr["exec_decode"]["launches"]*1e3,1), "fail", int(fail.sum()))
    for _ in range(3): codec.encode(info, out=cw)
    codec.profile_read(reset=True); codec.profile_enable(True)
    for _ in range(20): codec.encode(info, out=cw)
    pr = codec.profile_read(reset=True); codec.profile_enable(False)
    print(B, "encode us", round(pr["exec_encode"]["ms"]/pr["exec_encode"]["launches"]*1e3,1))
