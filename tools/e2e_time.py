#!/usr/bin/env python3
"""Time the host-buffer decode entry points (pinned host buffers, copies inside the call) for one batch: Gbit/s of decoded
information.  API=copy (ldpc_decode_host) | inplace (ldpc_decode_host_inplace); LDPC_CUDA_HOST_GATHER / LDPC_CUDA_HOST_CHUNK_MB
are read by the library."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldpc_erasure_codes_b200.codec import LdpcCodec, fill_random
B = int(os.environ.get("B", "32768")); P = 13; S = 64; R = int(os.environ.get("R", "4")); api = os.environ.get("API", "copy"); mode = os.environ.get("MODE", "peel"); it = 10 if mode == "hybrid" else 50
codec = LdpcCodec(code=1, symbol_bytes=S, device=0, max_batch=65536)
info = torch.empty((B, codec.k, S), dtype=torch.uint8, device="cuda"); fill_random(info, 1)
cw = codec.encode(info); mask = codec.gen_erasures(B, 7, P=P, payload=cw)
ref_out, ref_fail = codec.decode(cw, mask, max_iter=it, mode=mode)
h_cw = cw.cpu().pin_memory(); h_mask = mask.cpu().pin_memory()
h_out = torch.empty((B, codec.k, S), dtype=torch.uint8).pin_memory(); h_fail = torch.empty((B,), dtype=torch.uint8).pin_memory()
del cw
if api == "copy":
    run = lambda: codec.decode_host(h_cw, h_mask, max_iter=it, mode=mode, out=h_out, fail=h_fail)
else:
    run = lambda: codec.decode_host_inplace(h_cw, h_mask, max_iter=it, mode=mode, fail=h_fail)
run()
got = h_out if api == "copy" else h_cw[:, :codec.k]
assert bool((got == ref_out.cpu()).all()) and bool((h_fail == ref_fail.cpu()).all())
ts = []
for _ in range(R):
    t0 = time.perf_counter(); run(); ts.append(time.perf_counter() - t0)
ts.sort()
print(f"mode={mode} api={api} gather={os.environ.get('LDPC_CUDA_HOST_GATHER', 'default')} chunk_mb={os.environ.get('LDPC_CUDA_HOST_CHUNK_MB', 'default')} B={B}: "
      f"{B * codec.k * S * 8 / ts[len(ts)//2] / 1e9:.1f} Gbit/s (best {B * codec.k * S * 8 / ts[0] / 1e9:.1f}), {ts[len(ts)//2]*1e3:.1f} ms", flush=True)
