#!/usr/bin/env python3
"""A small pass over every kernel family, checked against the oracle: what tools/sanitize.sh runs under
compute-sanitizer (encode, channel, peel, executor, hybrid stages, RS, packet front-ends, host pipeline)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ldpc_erasure_codes_b200.codec import LdpcCodec, RsCodec, fill_random, pack_mask, unpack_mask
from oracle import oracle as orc


def ok(name):
    torch.cuda.synchronize()
    print(f"[pass] {name}", flush=True)


for ci, S, B, P in ((1, 64, 40, 13), (0, 32, 24, 24), (2, 16, 12, 19)):
    codec = LdpcCodec(code=ci, symbol_bytes=S, device=0, max_batch=16)      # B > max_batch: chunking + two streams
    code = orc.Code.builtin(ci)
    info = torch.empty((B, codec.k, S), dtype=torch.uint8, device="cuda")
    fill_random(info, seed=ci + 1)
    cw = codec.encode(info)
    assert np.array_equal(cw.cpu().numpy(), orc.encode(code, info.cpu().numpy()))
    ok(f"encode code {ci} S={S}")
    rx = cw.clone()
    mask = codec.gen_erasures(B, 5 + ci, P=P, payload=rx)
    flags = orc.gen_erasures_iid(code.n, 5 + ci, B, P=P)
    assert np.array_equal(unpack_mask(mask, code.n), flags)
    ok(f"channel code {ci}")
    for mode, it in (("peel", 50), ("hybrid", 10)):
        fa = torch.zeros(B, dtype=torch.uint8, device="cuda")
        out, fail = codec.decode(rx, mask, max_iter=it, mode=mode, fail_any=fa)
        ref = orc.decode(code, rx.cpu().numpy(), flags, max_iter=it, mode=mode)
        assert np.array_equal(out.cpu().numpy(), ref["out"]) and np.array_equal(fail.cpu().numpy(), ref["fail_sys"])
        ok(f"decode {mode} code {ci}")
    if ci == 1:
        h_out, h_fail = codec.decode_host(rx.cpu().pin_memory(), mask.cpu().pin_memory())
        assert np.array_equal(h_out.numpy(), orc.decode(code, rx.cpu().numpy(), flags, max_iter=50)["out"])
        ok("host pipeline")
        pk = codec.packetize(cw[:3], block0=250)
        keep = torch.randperm(pk.shape[0], device="cuda")[: int(pk.shape[0] * 0.85)]
        cw2, m2, counts = codec.depacketize(pk[keep].contiguous(), 250, 3)
        r_cw, r_flags, r_counts = orc.depacketize(pk[keep].cpu().numpy(), code.n, 250, 3)
        assert np.array_equal(cw2.cpu().numpy(), r_cw) and np.array_equal(unpack_mask(m2, code.n), r_flags)
        ok("packet front-ends")
        codec.simulate_fer(64, 3, P=12, max_iter=10, mode="hybrid")
        ok("simulate_fer hybrid")
    codec.close()

rs = RsCodec(n=255, k=191, symbol_bytes=64, device=0, max_batch=8)
G = orc.rs_gsys(255, 191)
info = torch.empty((6, 191, 64), dtype=torch.uint8, device="cuda")
fill_random(info, seed=9)
cw = rs.encode(info)
assert np.array_equal(cw.cpu().numpy()[0], orc.rs_encode(G, info.cpu().numpy()[0]))
flags = (np.random.default_rng(3).random((6, 255)) < 0.2).astype(np.uint8)
rx = cw.cpu().numpy().copy()
rx[flags == 1] = 0
out, fail = rs.decode(torch.from_numpy(rx).cuda(), torch.from_numpy(pack_mask(flags)).cuda())
good = fail.cpu().numpy() == 0
assert np.array_equal(out.cpu().numpy()[good], info.cpu().numpy()[good])
ok("rs encode/decode")
print("[pass] all", flush=True)
