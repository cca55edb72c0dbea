#!/usr/bin/env python3
"""Device-resident kernel timings for the BASELINE.json configurations other than the headline one
(which bench.py measures): one JSON line per measurement.  Times are CUDA-event times recorded by the
library around its own launches (ldpc_profile_*); inputs are generated on the device beforehand."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ldpc_erasure_codes_b200.codec import LdpcCodec, RsCodec, fill_random

PEAK = 6554.2
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def emit(**kw):
    print(json.dumps(kw), flush=True)


def ldpc_case(name, ci, S, B, chan, mode="peel", max_iter=50, reps=3, do_encode=True):
    codec = LdpcCodec(code=ci, symbol_bytes=S, device=0, max_batch=B)
    n, k = codec.n, codec.k
    info = torch.empty((B, k, S), dtype=torch.uint8, device="cuda"); fill_random(info, 1)
    cw = torch.empty((B, n, S), dtype=torch.uint8, device="cuda")
    out = torch.empty((B, k, S), dtype=torch.uint8, device="cuda"); fail = torch.empty((B,), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        codec.encode(info, out=cw)
    if do_encode:
        codec.profile_read(reset=True); codec.profile_enable(True)
        for _ in range(reps):
            codec.encode(info, out=cw)
        pr = codec.profile_read(reset=True); codec.profile_enable(False)
        ms = pr["exec_encode"]["ms"] / reps
        emit(config=name, op="encode", code=ci, S=S, B=B, ms=round(ms, 3), info_gbit_s=round(B * k * S * 8 / ms / 1e6, 1),
             GBs=round((k * S + n * S) * B / ms / 1e6, 1), frac_hbm=round((k * S + n * S) * B / ms / 1e6 / PEAK, 3))
    mask = codec.gen_erasures(B, 12345, payload=cw, **chan)
    for _ in range(2):
        codec.decode(cw, mask, max_iter=max_iter, mode=mode, out=out, fail=fail)
    good = fail == 0
    assert bool((out[good] == info[good]).all()), "round trip failed"
    codec.reset_stats(); codec.profile_read(reset=True); codec.profile_enable(True)
    for _ in range(reps):
        codec.decode(cw, mask, max_iter=max_iter, mode=mode, out=out, fail=fail)
    pr = codec.profile_read(reset=True); codec.profile_enable(False)
    st = codec.stats()
    ms_parts = {kk: round(pr[kk]["ms"] / reps, 3) for kk in ("peel", "exec_decode")}
    ms_parts["hybrid"] = round(sum(pr[kk]["ms"] for kk in ("hybrid", "hybrid_apply", "hybrid_warp", "hybrid_cta")) / reps, 3)   # all elimination stages
    ms_parts["hybrid_stages"] = [round(pr[kk]["ms"] / reps, 3) for kk in ("hybrid", "hybrid_apply", "hybrid_warp", "hybrid_cta")]
    ms = ms_parts["peel"] + ms_parts["exec_decode"] + ms_parts["hybrid"]
    alg = (n * S + (n + 7) // 8 + k * S + 1) * B
    emit(config=name, op="decode", mode=mode, code=ci, S=S, B=B, channel={k_: (list(v) if isinstance(v, tuple) else v) for k_, v in chan.items()},
         max_iter=max_iter, erasure_rate=round(float(torch.tensor(0.0) + sum(bin(x & 0xFFFFFFFF).count("1") for x in mask[:64].flatten().tolist()) / (64 * n)), 4),
         ms=round(ms, 3), ms_parts=ms_parts, info_gbit_s=round(B * k * S * 8 / ms / 1e6, 1),
         exec_GBs=round(alg / ms_parts["exec_decode"] / 1e6, 1), exec_frac_hbm=round(alg / ms_parts["exec_decode"] / 1e6 / PEAK, 3),
         whole_frac_hbm=round(alg / ms / 1e6 / PEAK, 3), fer=round(float(fail.float().mean()), 5),
         ml_attempt_rate=round(st["ml_attempts"] / max(1, st["frames"]), 4), ml_fail_rate=round(st["ml_failures"] / max(1, st["frames"]), 5),
         rs_equiv_block_fer=round(st["rs_errors"] / max(1, st["frames"] * (n // codec.info.rs_n)), 6))
    codec.close()
    del info, cw, out, fail, mask
    torch.cuda.empty_cache()


def rs_case(name, n, k, S, B, p, reps=3):
    codec = RsCodec(n=n, k=k, symbol_bytes=S, device=0, max_batch=B)
    info = torch.empty((B, k, S), dtype=torch.uint8, device="cuda"); fill_random(info, 3)
    cw = codec.encode(info)
    mask = (torch.rand((B, n), device="cuda") < p)
    words = torch.zeros((B, codec.mask_words * 32), dtype=torch.bool, device="cuda"); words[:, :n] = mask
    packed = (words.view(B, codec.mask_words, 32).to(torch.int64) << torch.arange(32, device="cuda")).sum(-1)
    packed = packed.to(torch.int64).where(packed < 2 ** 31, packed - 2 ** 32).to(torch.int32).contiguous()
    rx = cw.clone(); rx[mask] = 0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    out, fail = codec.decode(rx, packed)
    good = fail == 0
    assert bool((out[good] == info[good]).all()), "RS round trip failed"
    ev[0].record()
    for _ in range(reps): codec.encode(info, out=cw)
    ev[1].record()
    for _ in range(reps): codec.decode(rx, packed, out=out, fail=fail)
    ev[2].record(); torch.cuda.synchronize()
    ms_e, ms_d = ev[0].elapsed_time(ev[1]) / reps, ev[1].elapsed_time(ev[2]) / reps
    t_avg = float(mask[:, :k].sum(1).float().mean())
    emit(config=name, op="rs_encode", n=n, k=k, S=S, B=B, ms=round(ms_e, 3), info_gbit_s=round(B * k * S * 8 / ms_e / 1e6, 1),
         gf_mac_per_s=round(B * k * (n - k) * S / ms_e / 1e6, 1))
    emit(config=name, op="rs_decode", n=n, k=k, S=S, B=B, p=p, ms=round(ms_d, 3), info_gbit_s=round(B * k * S * 8 / ms_d / 1e6, 1),
         GBs=round((n * S + 32 + k * S + 1) * B / ms_d / 1e6, 1), frac_hbm=round((n * S + 32 + k * S + 1) * B / ms_d / 1e6 / PEAK, 4),
         gf_gmac_per_s=round(B * k * t_avg * S / ms_d / 1e6, 1), block_fer=round(float(fail.float().mean()), 5))
    codec.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["1", "2h", "3", "4", "5"]
    if "1" in which:   # config 1: n2000_k1000 triangular H, reference symbol size (1024 B), 30 % -> 19/64
        ldpc_case("config1_n2000_k1000_S1024_P19", 0, 1024, 4096, dict(P=19))
    if "2h" in which:  # config 2 in hybrid mode (10 sweeps + elimination)
        ldpc_case("config2_hybrid_P13", 1, 64, 32768, dict(P=13), mode="hybrid", max_iter=10, do_encode=False)
        ldpc_case("config2_hybrid_P12", 1, 64, 32768, dict(P=12), mode="hybrid", max_iter=10, do_encode=False)
    if "3" in which:   # config 3: n4000_k2000 under the bursty channel, peel and hybrid
        ldpc_case("config3_n4000_bursty_repo_params", 2, 64, 16384, dict(bursty=(0.001, 0.1, 10.0)), mode="hybrid", max_iter=10)
        ldpc_case("config3_n4000_bursty_alpha.1_beta.4", 2, 64, 16384, dict(bursty=(0.1, 0.4, 10.0)), mode="hybrid", max_iter=10, do_encode=False)
        ldpc_case("config3_n4000_bursty_stress_.38_.9", 2, 64, 4096, dict(bursty=(0.38, 0.9, 10.0)), mode="hybrid", max_iter=10, do_encode=False)
    if "4" in which:   # config 4: RS(255,191) 1 KB symbols vs LDPC (2040,1530) at equal rate
        rs_case("config4_rs255_191_S1024", 255, 191, 1024, 2048, 0.2)
        rs_case("config4_rs255_192_S1024", 255, 192, 1024, 2048, 0.2)
        ldpc_case("config4_ldpc2040_S1024_p.2", 1, 1024, 2048, dict(p=0.2), do_encode=False)
    if "5" in which:   # config 5: n2040/k1530 S=64, erasure-rate sweep (the /64 quantisation of 10..30 %)
        for P in (6, 10, 13, 16, 19):
            ldpc_case(f"config5_P{P}", 1, 64, 32768, dict(P=P), do_encode=(P == 6))
