#!/usr/bin/env python3
"""Small fixed workload for ncu: encode, erase, then three peel decodes of B codewords (n2040/k1530, S=64, P=13/64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldpc_erasure_codes_b200.codec import LdpcCodec, fill_random
B = int(os.environ.get("B", "16384")); P = int(os.environ.get("P", "13")); S = int(os.environ.get("S", "64"))
codec = LdpcCodec(code=int(os.environ.get("CODE", "1")), symbol_bytes=S, device=0, max_batch=B)
if os.environ.get("GEOM"):
    w, s = map(int, os.environ["GEOM"].split("x")); codec.set_exec_geometry(w, s)
info = torch.empty((B, codec.k, S), dtype=torch.uint8, device="cuda"); fill_random(info, 1)
cw = codec.encode(info)
mask = codec.gen_erasures(B, 12345, P=P, payload=cw)
for _ in range(3):
    out, fail = codec.decode(cw, mask, mode=os.environ.get("MODE", "peel"))
torch.cuda.synchronize()
print("ok", float(fail.float().mean()))
