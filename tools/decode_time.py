#!/usr/bin/env python3
"""Time ldpc_decode (device buffers) for one batch size under the current environment: ms per call, median of R."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldpc_erasure_codes_b200.codec import LdpcCodec, fill_random
B = int(os.environ.get("B", "65536")); P = int(os.environ.get("P", "13")); S = int(os.environ.get("S", "64")); R = int(os.environ.get("R", "10")); MODE = os.environ.get("MODE", "peel"); IT = int(os.environ.get("IT", "50" if MODE == "peel" else "10"))
codec = LdpcCodec(code=int(os.environ.get("CODE", "1")), symbol_bytes=S, device=0, max_batch=B)
info = torch.empty((B, codec.k, S), dtype=torch.uint8, device="cuda"); fill_random(info, 1)
cw = codec.encode(info); mask = codec.gen_erasures(B, 4242, P=P, payload=cw)
out = torch.empty_like(info); fail = torch.empty(B, dtype=torch.uint8, device="cuda")
for _ in range(3): codec.decode(cw, mask, max_iter=IT, mode=MODE, out=out, fail=fail)
ts = []
for _ in range(R):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); codec.decode(cw, mask, max_iter=IT, mode=MODE, out=out, fail=fail); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
ok = fail == 0
assert bool((out[ok] == info[ok]).all()), "mismatch"
codec.profile_enable(True); codec.decode(cw, mask, max_iter=IT, mode=MODE, out=out, fail=fail); pr = codec.profile_read(reset=True)
hy = "/".join(f"{pr.get(k, {}).get('ms', 0.0):.3f}" for k in ("hybrid", "hybrid_apply", "hybrid_warp", "hybrid_cta"))
print(f"B={B} P={P} mode={MODE} hybrid_ms(inact/apply/warp/cta)={hy}: "
      f"{ts[len(ts)//2]:.3f} ms/call (min {ts[0]:.3f}); serial peel {pr['peel']['ms']:.3f} exec {pr['exec_decode']['ms']:.3f}  FER {float(fail.float().mean()):.4f}")
if MODE == "hybrid" and os.environ.get("LDPC_CUDA_PHASE_TIMING"):
    g = pr["apply_phase_cycles"]; n = max(1, g[6])
    print("  apply kernel, warp cycles per codeword: plan %d rhs %d replay %d dense %d out %d  (%d codewords)" % tuple([x // n for x in g[:5]] + [g[6]]))
    g = pr["ge_phase_cycles"]; n = max(1, g[6])
    print("  inactivation stage WITH payload, warp cycles per codeword: setup %d rows %d adj+synd %d peel %d dense %d out %d  (%d codewords)" % tuple([x // n for x in g[:6]] + [g[6]]))
if MODE == "hybrid":   # the same frames, pattern only (no payload): isolates the elimination from the syndrome gathers
    codec.profile_read(reset=True); codec.profile_enable(True)
    codec.simulate_fer(B, 4242, P=P, max_iter=IT, mode="hybrid"); pr = codec.profile_read(reset=True)
    hy = "/".join(f"{pr.get(k, {}).get('ms', 0.0):.3f}" for k in ("hybrid", "hybrid_apply", "hybrid_warp", "hybrid_cta"))
    print(f"  pattern-only: hybrid_ms(inact/apply/warp/cta)={hy} peel {pr['peel']['ms']:.3f}")
    if os.environ.get("LDPC_CUDA_PHASE_TIMING"):
        g = pr["ge_phase_cycles"]; n = max(1, g[6])
        print("  inactivation stage, warp cycles per codeword: setup %d rows %d adj+synd %d peel %d dense %d out %d  (%d codewords)" % tuple([x // n for x in g[:6]] + [g[6]]))
