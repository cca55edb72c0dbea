#!/usr/bin/env python3
"""Randomised soak of the RS codec against the oracle (restated MATLAB encoder / decoder): random (n, k) with n - k <= 128,
symbol sizes 16 .. 2048, erasure probabilities around the code's limit, for a wall-clock budget.  Also checks that the
decoder returns the information wherever it reports success (the code is MDS: any k received symbols decode)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ldpc_erasure_codes_b200.codec import RsCodec, fill_random, pack_mask
from oracle import oracle as orc

budget = float(os.environ.get("FUZZ_SECONDS", "120")); rng = np.random.default_rng(int(os.environ.get("SEED", "1")))
t0 = time.time(); runs = 0; bad = 0; frames = 0; undecodable = 0
while time.time() - t0 < budget:
    n = int(rng.integers(3, 256)); r = int(rng.integers(1, min(128, n - 1) + 1)); k = n - r
    S = int(rng.choice([16, 32, 48, 64, 128, 256, 1024, 2048])); B = int(rng.integers(1, 40))
    p = float(np.clip(rng.uniform(0.0, 1.3) * r / n, 0.0, 0.95))
    codec = RsCodec(n=n, k=k, symbol_bytes=S, device=0, max_batch=64)
    G = orc.rs_gsys(n, k)
    assert np.array_equal(codec.generator(), G)
    info = torch.empty((B, k, S), dtype=torch.uint8, device="cuda"); fill_random(info, int(rng.integers(1, 2**31)))
    cw = codec.encode(info).cpu().numpy(); inf = info.cpu().numpy()
    ok = all(np.array_equal(cw[b], orc.rs_encode(G, inf[b])) for b in range(min(B, 3)))
    flags = (rng.random((B, n)) < p).astype(np.uint8)
    rx = cw.copy(); rx[flags == 1] = 0
    out, fail = codec.decode(torch.from_numpy(rx).cuda(), torch.from_numpy(pack_mask(flags)).cuda())
    out, fail = out.cpu().numpy(), fail.cpu().numpy()
    for b in range(B):
        rec = np.nonzero(flags[b] == 0)[0]
        if len(rec) >= k:
            if b < 4:
                ref, rd = orc.rs_decode(G, rec[:k].astype(np.int32), rx[b][rec[:k]])
                ok = ok and rd == 0 and np.array_equal(out[b], ref)
            ok = ok and fail[b] == 0 and np.array_equal(out[b], inf[b])
        else:
            want = inf[b].copy(); want[flags[b, :k] == 1] = 0
            ok = ok and fail[b] == 1 and np.array_equal(out[b], want)
            undecodable += 1
    frames += B; runs += 1
    if not ok:
        bad += 1
        print(f"MISMATCH n={n} k={k} S={S} B={B} p={p:.3f}", flush=True)
    codec.close()
print(f"rs fuzz: {runs} runs, {frames} codewords ({undecodable} undecodable), {bad} mismatches, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
