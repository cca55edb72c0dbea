#!/usr/bin/env python3
"""Randomised soak test: CUDA path vs the oracle over random codes, symbol sizes, erasure rates, channels, batch sizes,
modes and iteration caps, for a given wall-clock budget.  Prints one line per mismatch (none expected) and a summary."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ldpc_erasure_codes_b200.codec import LdpcCodec, fill_random, unpack_mask
from oracle import oracle as orc

budget = float(os.environ.get("FUZZ_SECONDS", os.environ.get("SECONDS", "240"))); seed0 = int(os.environ.get("SEED", "1"))
rng = np.random.default_rng(seed0)
codes = {ci: orc.Code.builtin(ci) for ci in (0, 1, 2)}
cache = {}
def codec_for(ci, S):
    if (ci, S) not in cache:
        if len(cache) > 6: cache.pop(next(iter(cache))).close()
        cache[(ci, S)] = LdpcCodec(code=ci, symbol_bytes=S, device=0, max_batch=96)
    return cache[(ci, S)]
t0 = time.time(); runs = 0; bad = 0; ge_frames = 0; ge_fail = 0
while time.time() - t0 < budget:
    ci = int(rng.choice([0, 1, 1, 1, 2])); S = int(rng.choice([16, 32, 48, 64, 64, 128])); B = int(rng.integers(1, 200))
    mode = str(rng.choice(["peel", "hybrid", "hybrid"])); it = int(rng.choice([0, 1, 2, 5, 10, 50]))
    code = codes[ci]; codec = codec_for(ci, S)
    thr = {0: 24, 1: 13, 2: 24}[ci]
    seed = int(rng.integers(1, 2**31))
    valid = mode == "hybrid" or rng.random() < 0.7
    if valid:
        info = torch.empty((B, codec.k, S), dtype=torch.uint8, device="cuda"); fill_random(info, seed)
        cw = codec.encode(info)
    else:
        cw = torch.empty((B, codec.n, S), dtype=torch.uint8, device="cuda"); fill_random(cw, seed)
    rx = cw.clone()
    if rng.random() < 0.75:
        P = int(np.clip(thr + rng.integers(-6, 5), 0, 64))
        mask = codec.gen_erasures(B, seed, P=P, payload=rx); desc = f"P={P}"
    else:
        a, b = float(rng.uniform(0.0, 0.45)), float(rng.uniform(0.1, 0.95))
        mask = codec.gen_erasures(B, seed, bursty=(a, b, 10.0), payload=rx); desc = f"bursty({a:.2f},{b:.2f})"
    flags = unpack_mask(mask, code.n)
    codec.reset_stats()
    out, fail = codec.decode(rx, mask, max_iter=it, mode=mode)
    st = codec.stats()
    ref = orc.decode(code, rx.cpu().numpy(), flags, max_iter=it, mode=mode)
    ok = np.array_equal(fail.cpu().numpy(), ref["fail_sys"])
    if mode == "peel":
        ok = ok and np.array_equal(out.cpu().numpy(), ref["out"])
    else:
        good = ref["fail_sys"] == 0
        ok = ok and np.array_equal(out.cpu().numpy()[good], ref["out"][good])
        ok = ok and st["ml_attempts"] == int((ref["status"] > 0).sum()) and st["ml_failures"] == int((ref["status"] == 2).sum())
        ge_frames += int((ref["status"] > 0).sum()); ge_fail += int((ref["status"] == 2).sum())
    runs += 1
    if not ok:
        bad += 1
        print("MISMATCH", json.dumps(dict(code=ci, S=S, B=B, mode=mode, max_iter=it, chan=desc, seed=seed, valid=bool(valid))), flush=True)
print(json.dumps(dict(runs=runs, mismatches=bad, seconds=round(time.time() - t0, 1), frames_eliminated=ge_frames, rank_deficient=ge_fail)))
