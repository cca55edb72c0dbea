#!/usr/bin/env python3
"""bench.py -- decoded information Gbit/s of the peeling erasure decoder on B200.

Workload (BASELINE.json configs[1]): (n=2040, k=1530) irregular code, 64-byte symbols, i.i.d.
erasures at P/64 = 13/64 = 20.3 % (the reference's own rate quantisation, decoder_top.cl:105),
one step = 1,048,576 codewords per GPU decoded through the C ABI (ldpc_decode), peeling mode,
max_iter 50 (the reference host default, main.cpp:99).

A 1 Mi-codeword batch is 130.6 GB in + 97.9 GB out and does not fit in 180 GB of HBM next to
its output, so the step walks a RESIDENT set of distinct sub-batches (default 8 x 65,536
codewords = 68 GB in + 51 GB out) twice; every sub-batch (8.5 GB) is far larger than the 126 MB
L2, so no timed launch finds its input in cache.  Inputs are encoded and erased on the device
before the timed region.

  value     = N_gpus * codewords_per_step * k * S * 8 / step time  (info bits, main.cpp:655)
  e2e       = same metric through ldpc_decode_host(): pinned HOST buffers, H2D + kernels + D2H
              inside the timed region (a smaller batch per step, stated in the JSON)
  roofline  = payload_exec_kernel (the dominant kernel): algorithmic bytes per launch /
              CUDA-event time of that kernel, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline = the reference algorithm restated in C (oracle/, OpenMP over codewords) on this
              box's host cores, bounded sample

`--impl reference` times that CPU restatement alone (PoCL / the Intel FPGA OpenCL SDK / MATLAB
are not installable offline, so the reference's own implementation cannot run here).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CODE_IND = 1          # BASELINE.json's metric is quoted on (2040,1530); -c selects another built-in code
CODE_NAMES = {0: "n2000_k1000 triangular H", 1: "n2040_k1530 irregular H", 2: "n4000_k2000 triangular H"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--symbol-bytes", type=int, default=64)
    # the reference host's own flags (main.cpp:157-170): -p PER numerator / 64, -n frames, -i iterations, -c code
    ap.add_argument("-p", "--per64", type=int, default=13, help="erasure rate numerator / 64 (reference flag -p)")
    ap.add_argument("-n", "--codewords", type=int, default=1 << 20, help="codewords (frames) per GPU per step (reference flag -n)")
    ap.add_argument("--sub-batch", type=int, default=1 << 16)
    ap.add_argument("--resident", type=int, default=8, help="distinct sub-batches kept in HBM")
    ap.add_argument("-i", "--max-iter", type=int, default=50, help="sweeps over the checks (reference flag -i, default 50)")
    ap.add_argument("-c", "--code", type=int, default=1, help="0 = (2000,1000), 1 = (2040,1530) [the benchmark], 2 = (4000,2000)")
    ap.add_argument("--mode", default="peel", choices=["peel", "hybrid"])
    ap.add_argument("--e2e-batch", type=int, default=1 << 15)
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--slice-bytes", type=int, default=0)
    ap.add_argument("--slots", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def algorithmic_bytes_decode(n, k, S):
    """n*S + ceil(n/8) + k*S + 1 per codeword."""
    return n * S + (n + 7) // 8 + k * S + 1


# ------------------------------------------------------------------------------------------
# CPU reference arm (the oracle = restatement of ldpc_erasure_decoder.cl, all host cores)
# ------------------------------------------------------------------------------------------
def cpu_decode_rate(S, P, seed, seconds, max_iter, mode="peel"):
    """Returns (info Gbit/s, threads, sample description, codewords, elapsed)."""
    import numpy as np
    from oracle import oracle as orc

    code = orc.Code.builtin(CODE_IND)
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1: ask explicitly)
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    rng = np.random.default_rng(seed)

    def make(B, frame0):
        info = rng.integers(0, 256, (B, code.k, S), dtype=np.uint8)
        cw = orc.encode(code, info)
        flags = orc.gen_erasures_iid(code.n, seed, B, P=P, frame0=frame0)
        cw[flags == 1] = 0
        return cw, flags

    # calibrate on a small batch, then size the sample for ~`seconds` of CPU work
    cw, flags = make(64 * threads, 0)
    t0 = time.perf_counter()
    orc.decode(code, cw, flags, max_iter=max_iter, mode=mode, inplace=True, nthreads=threads)
    dt = time.perf_counter() - t0
    rate = 64 * threads / dt
    B = int(max(64 * threads, min(rate * seconds, 262144)))
    cw, flags = make(B, 1 << 20)
    t0 = time.perf_counter()
    orc.decode(code, cw, flags, max_iter=max_iter, mode=mode, inplace=True, nthreads=threads)
    dt = time.perf_counter() - t0
    gbps = B * code.k * S * 8 / dt / 1e9
    return gbps, threads, f"{B} codewords {CODE_NAMES[CODE_IND].split()[0]} S={S} P={P}/64, reference sweep decoder, early stop", B, dt


def run_reference(args):
    """--impl reference: K timed steps of the CPU restatement, each a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = max(1.0, min(args.cpu_seconds, 60.0 / max(1, args.steps + args.warmup)))
    vals = []
    threads, sample = 1, ""
    ms = []
    for i in range(args.warmup + args.steps):
        g, threads, sample, B, dt = cpu_decode_rate(args.symbol_bytes, args.per64, args.seed + i, per_step,
                                                    args.max_iter, args.mode)
        if i >= args.warmup:
            vals.append(g)
            ms.append(dt * 1e3)
    v = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "decoded info Gbit/s (n2040 k1530, 20% erasures)", "value": v, "unit": "Gbit/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sum(ms) / len(ms),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{CODE_NAMES[CODE_IND]}, {args.symbol_bytes}-byte symbols, {args.per64}/64 "
                               f"({100 * args.per64 / 64:.1f}%) i.i.d. erasures, {args.mode} decode",
                   "symbol_bytes": args.symbol_bytes, "per64": args.per64, "max_iter": args.max_iter, "mode": args.mode},
        "cpu_baseline": {"value": v, "unit": "Gbit/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference algorithm restated in C (oracle/ldpc_oracle.c) on host cores; the reference's own "
                "Intel-FPGA OpenCL / MATLAB code cannot run here (no PoCL, no aoc, no MATLAB offline)",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from ldpc_erasure_codes_b200 import sharding
    from ldpc_erasure_codes_b200.codec import LdpcCodec, fill_random

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the codec has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    S, P = args.symbol_bytes, args.per64
    sub = args.sub_batch
    n_sub = (args.codewords + sub - 1) // sub
    codewords = n_sub * sub
    resident = min(args.resident, n_sub)
    codec = LdpcCodec(code=CODE_IND, symbol_bytes=S, device=local_rank, max_batch=sub)
    if args.slice_bytes or args.slots:
        codec.set_exec_geometry(args.slice_bytes, args.slots)
    n, k = codec.n, codec.k

    # ---- resident inputs: encode + erase on the device (untimed) -----------------------
    rx = torch.empty((resident, sub, n, S), dtype=torch.uint8, device=dev)
    masks = torch.empty((resident, sub, codec.mask_words), dtype=torch.int32, device=dev)
    out = torch.empty((resident, sub, k, S), dtype=torch.uint8, device=dev)
    fail = torch.empty((resident, sub), dtype=torch.uint8, device=dev)
    info = torch.empty((sub, k, S), dtype=torch.uint8, device=dev)
    frame_base = sharding.weak_frame_base(codewords, rank)   # global frame index: results do not depend on the sharding
    for r in range(resident):
        fill_random(info, seed=args.seed, block0=(frame_base + r * sub) * k * S // 16)
        codec.encode(info, out=rx[r])
        codec.gen_erasures(sub, args.seed, P=P, frame0=frame_base + r * sub, payload=rx[r], mask=masks[r])
    torch.cuda.synchronize()

    rx_all, masks_all = rx.view(resident * sub, n, S), masks.view(resident * sub, codec.mask_words)
    out_all, fail_all = out.view(resident * sub, k, S), fail.view(resident * sub)

    def step():
        # one ldpc_decode call per pass over the resident set; the library cuts it into max_batch (= sub) chunks
        for i in range(0, n_sub, resident):
            cnt = min(resident, n_sub - i) * sub
            codec.decode(rx_all[:cnt], masks_all[:cnt], max_iter=args.max_iter, mode=args.mode, out=out_all[:cnt], fail=fail_all[:cnt])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    # round-trip property on the whole resident set (size-independent check at full size):
    # every frame the decoder reports as good must equal the encoder's input
    torch.cuda.synchronize()
    fill_random(info, seed=args.seed, block0=(frame_base + 0 * sub) * k * S // 16)
    good = fail[0] == 0
    assert bool((out[0][good] == info[good]).all()), "round-trip check failed on the bench data"
    fer = float(fail.float().mean().item())
    codec.reset_stats()
    codec.profile_read(reset=True)

    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    launches = sum(v["launches"] for v in codec.profile_read(reset=True).values() if isinstance(v, dict))
    job_stats = sharding.reduce_stats(codec.stats())             # the reference's ERROR_STAT counters, summed over ranks
    ms_step = sharding.reduce_max(ms_total) / args.steps      # slowest rank, device-timed
    value = world * codewords * k * S * 8 / (ms_step * 1e-3) / 1e9

    # ---- per-kernel device time (separate pass with event brackets around each launch) ---
    codec.profile_enable(True)
    step()
    prof = codec.profile_read(reset=True)
    codec.profile_enable(False)
    exec_ms = prof["exec_decode"]["ms"] / max(1, prof["exec_decode"]["launches"])
    peel_ms = prof["peel"]["ms"] / max(1, prof["peel"]["launches"])
    alg_bytes = algorithmic_bytes_decode(n, k, S) * sub
    achieved = alg_bytes / (exec_ms * 1e-3) / 1e9
    peak, peak_src = 6650.0, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"])
            peak_src = "measured"
    except Exception:
        pass
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            per_cw = json.load(f).get("payload_exec_decode_dram_bytes_per_codeword")
            traffic = per_cw * sub if per_cw else None      # ncu capture of a smaller launch, scaled per codeword
    except Exception:
        pass

    # ---- end to end through the host-buffer entry point -------------------------------------
    e2e = None
    if not args.no_e2e:
        eb = min(args.e2e_batch, sub)
        h_cw = torch.empty((eb, n, S), dtype=torch.uint8, pin_memory=True)
        h_mask = torch.empty((eb, codec.mask_words), dtype=torch.int32, pin_memory=True)
        h_out = torch.empty((eb, k, S), dtype=torch.uint8, pin_memory=True)
        h_fail = torch.empty((eb,), dtype=torch.uint8, pin_memory=True)
        h_cw.copy_(rx[0][:eb])
        h_mask.copy_(masks[0][:eb])
        torch.cuda.synchronize()
        for _ in range(2):
            codec.decode_host(h_cw, h_mask, max_iter=args.max_iter, mode=args.mode, out=h_out, fail=h_fail)
        assert bool((h_out == out[0][:eb].cpu()).all()) and bool((h_fail == fail[0][:eb].cpu()).all())
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            codec.decode_host(h_cw, h_mask, max_iter=args.max_iter, mode=args.mode, out=h_out, fail=h_fail)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dt = sharding.reduce_max(dt)
        e2e = {"value": world * eb * args.steps * k * S * 8 / dt / 1e9, "unit": "Gbit/s",
               "h2d_bytes_per_step": eb * (n * S + codec.mask_words * 4), "d2h_bytes_per_step": eb * (k * S + 1),
               "codewords_per_step": eb, "api": "ldpc_decode_host (pinned host buffers)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        g, threads, sample, _, _ = cpu_decode_rate(S, P, args.seed, args.cpu_seconds, args.max_iter, args.mode)
        cpu = {"value": g, "unit": "Gbit/s", "cores": threads, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": "decoded info Gbit/s (n2040 k1530, 20% erasures)", "value": value, "unit": "Gbit/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {
                "workload": f"{CODE_NAMES[CODE_IND]}, {S}-byte symbols, {P}/64 ({100 * P / 64:.1f}%) i.i.d. erasures, "
                            f"{codewords} codewords per GPU per step, {args.mode} decode",
                "symbol_bytes": S, "per64": P, "max_iter": args.max_iter, "mode": args.mode,
                "codewords_per_gpu_per_step": codewords, "sub_batch": sub, "resident_sub_batches": resident,
                "calls": "one ldpc_decode call per pass over the resident sub-batches; the library runs it as max_batch-sized chunks "
                         "(one peel + one executor launch each) alternating between two internal streams",
                "l2": "every launch reads a distinct 8.5 GB sub-batch (>> 126 MB L2), no flush needed",
                "slice_bytes": codec.info.slice_bytes, "exec_slots": codec.info.exec_slots,
                "frame_error_rate": fer, "parallelism": f"codeword-sharded x{world}, no collectives",
                "counters": job_stats,
            },
            "roofline": {"bound": "hbm", "kernel": "payload_exec_kernel(decode)", "achieved": achieved, "peak": peak,
                         "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": exec_ms,
                         "peel_schedule_ms_per_launch": peel_ms,
                         "whole_step_frac": (algorithmic_bytes_decode(n, k, S) * codewords / (ms_step * 1e-3) / 1e9) / peak},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    global CODE_IND
    args = parse_args()
    CODE_IND = args.code
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
