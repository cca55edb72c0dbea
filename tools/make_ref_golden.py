#!/usr/bin/env python3
"""Generates tests/golden/ref_vectors.json from oracle/_ref -- the reference's OWN device sources
(ldpc_erasure_encoder.cl, ldpc_erasure_decoder.cl, data_in of ldpc_erasure_decoder_top.cl) compiled
unmodified by gcc (oracle/Makefile).  Run in the authoring container, where /root/reference exists:

    python tools/make_ref_golden.py

The file holds, per seeded case, digests of what the reference produced: the encoder's codewords, the
generator's erasure flags, the decoder's k output symbols and their is_erasure flags.  Inputs are a pure
function of the case (see `case_inputs`), so the tests rebuild them anywhere and compare the oracle
restatement and the CUDA path against these digests without the reference tree.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [dict(code=ci, S=S, P=P, num_iter=it, B=6, seed=1000 * ci + 10 * P + it + S)
         for ci in (0, 1) for S in (16, 64, 1024) for P in (9, 13, 19, 24) for it in (1, 3, 50)]


def mix_bytes(count, seed):
    """`count` pseudo-random bytes as a pure function of (index, seed): splitmix64 finaliser over a counter."""
    nw = (count + 7) // 8
    with np.errstate(over="ignore"):
        v = (np.arange(nw, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x632BE59BD9B4E019)) * np.uint64(0x9E3779B97F4A7C15)
        v ^= v >> np.uint64(30)
        v *= np.uint64(0xBF58476D1CE4E5B9)
        v ^= v >> np.uint64(27)
        v *= np.uint64(0x94D049BB133111EB)
        v ^= v >> np.uint64(31)
    return v.view(np.uint8)[:count].copy()


def case_inputs(case, k):
    return mix_bytes(case["B"] * k * case["S"], case["seed"]).reshape(case["B"], k, case["S"])


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:32]


def main():
    from oracle import ref
    assert ref.available(), "oracle/_ref could not be built (no reference tree?)"
    out = []
    for case in CASES:
        ci, S, P, it = case["code"], case["S"], case["P"], case["num_iter"]
        n, k = ref.code_params(ci)[:2]
        info = case_inputs(case, k)
        cw = ref.encode(ci, info)
        flags = ref.data_in(ci, case["seed"], P, case["B"])
        rx = cw.copy()
        rx[flags == 1] = 0                       # "erased = all zero" (ldpc_erasure_decoder.cl:17-20)
        dec = ref.decode(ci, rx, flags, num_iter=it, variant="canon")
        rec = dict(case)
        rec.update(cw=digest(cw), flags=digest(flags), erased=int(flags.sum()), out=digest(dec["out"]),
                   out_flags=digest(dec["out_flags"]), fail_sys=[int(x) for x in dec["fail_sys"]])
        if S == 1024:                            # the same through the committed top files (SYM_LEN = 128)
            assert digest(ref.encode(ci, info, top=True)) == rec["cw"]
            assert digest(ref.decode(ci, rx, flags, num_iter=it, variant="canon", top=True)["out"]) == rec["out"]
        out.append(rec)
    path = os.path.join(ROOT, "tests", "golden", "ref_vectors.json")
    with open(path, "w") as f:
        json.dump(dict(generator="tools/make_ref_golden.py", source="oracle/_ref (reference .cl compiled unmodified)", cases=out), f, indent=0)
    print(f"wrote {len(out)} cases to {path}")


if __name__ == "__main__":
    main()
