#!/usr/bin/env python3
"""Small fixed RS workload for ncu: encode + decode of B codewords, RS(255,191), 1 KB symbols, 20 % erasures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ldpc_erasure_codes_b200.codec import RsCodec, fill_random, pack_mask
B = int(os.environ.get("B", "1184"))
codec = RsCodec(n=255, k=191, symbol_bytes=1024, device=0, max_batch=B)
info = torch.empty((B, 191, 1024), dtype=torch.uint8, device="cuda"); fill_random(info, 1)
cw = codec.encode(info)
flags = (np.random.default_rng(1).random((B, 255)) < 0.2).astype(np.uint8)
mask = torch.from_numpy(pack_mask(flags)).cuda()
rx = cw.clone(); rx[torch.from_numpy(flags).cuda().bool()] = 0
for _ in range(2):
    out, fail = codec.decode(rx, mask)
torch.cuda.synchronize()
good = fail == 0
print("ok", bool((out[good] == info[good]).all()), float(fail.float().mean()))
