#!/usr/bin/env python3
"""Step-by-step bring-up check on a GPU box: each stage is compared with the oracle and
synchronised separately (LDPC_CUDA_DEBUG_SYNC=1 names the kernel that faults)."""
import os
import sys

os.environ.setdefault("LDPC_CUDA_DEBUG_SYNC", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ldpc_erasure_codes_b200.codec import LdpcCodec, fill_random, unpack_mask
from oracle import oracle as orc

ci = int(sys.argv[1]) if len(sys.argv) > 1 else 1
S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
B = int(sys.argv[3]) if len(sys.argv) > 3 else 40
P = int(sys.argv[4]) if len(sys.argv) > 4 else 13
code = orc.Code.builtin(ci)
codec = LdpcCodec(code=ci, symbol_bytes=S, device=0, max_batch=1024)
print("info:", {f: getattr(codec.info, f) for f, _ in codec.info._fields_}, flush=True)
rp, ci_ = codec.csr()
assert np.array_equal(rp, code.row_ptr) and np.array_equal(ci_, code.col_idx)
print("H loader == scipy", flush=True)

mask = codec.gen_erasures(B, 12345, P=P)
torch.cuda.synchronize()
flags = orc.gen_erasures_iid(code.n, 12345, B, P=P)
assert np.array_equal(flags, unpack_mask(mask, code.n)), "mask mismatch"
print("gen_erasures ok", flush=True)

info = torch.empty((B, codec.k, S), dtype=torch.uint8, device="cuda")
fill_random(info, seed=7)
torch.cuda.synchronize()
print("fill ok", info.flatten()[:8].tolist(), flush=True)
cw = codec.encode(info)
torch.cuda.synchronize()
cw_ref = orc.encode(code, info.cpu().numpy())
d = cw.cpu().numpy() != cw_ref
print("encode mismatching symbols:", int(d.any(axis=2).sum()), "of", d.shape[0] * d.shape[1], flush=True)
assert not d.any()

rx = cw.clone()
codec.gen_erasures(B, 12345, P=P, payload=rx, mask=mask)
torch.cuda.synchronize()
rx_ref = cw_ref.copy(); rx_ref[flags == 1] = 0
assert np.array_equal(rx.cpu().numpy(), rx_ref), "zeroing mismatch"
print("zero_erased ok", flush=True)
for max_iter in (50, 1, 2, 5):
    out, fail = codec.decode(rx, mask, max_iter=max_iter)
    torch.cuda.synchronize()
    ref = orc.decode(code, rx_ref, flags, max_iter=max_iter)
    dm = out.cpu().numpy() != ref["out"]
    print(f"max_iter={max_iter}: decode mismatching symbols:", int(dm.any(axis=2).sum()), "fail mismatch:",
          int((fail.cpu().numpy() != ref["fail_sys"]).sum()), "failed frames:", int(ref["fail_sys"].sum()), flush=True)
    assert not dm.any() and np.array_equal(fail.cpu().numpy(), ref["fail_sys"])
print("stats:", codec.stats())
print("ALL OK")
