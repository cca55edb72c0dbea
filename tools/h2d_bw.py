#!/usr/bin/env python3
"""Host <-> device copy bandwidth of this box with plain cudaMemcpyAsync, all visible GPUs at once (one host thread and two
streams per GPU, pinned buffers, H2D and D2H running concurrently): the ceiling of the host-buffer entry points
(ldpc_decode_host / ldpc_decode_host_multi), whatever the kernels do.  One JSON line."""
import json, os, sys, threading, time
import torch

mb = int(os.environ.get("MB", "2048")); reps = int(os.environ.get("REPS", "6"))
G = torch.cuda.device_count()
res = {}


def worker(g, both, barrier):
    torch.cuda.set_device(g)
    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(n, dtype=torch.uint8, device=f"cuda:{g}"); d_out = torch.empty(n, dtype=torch.uint8, device=f"cuda:{g}")
    s1, s2 = torch.cuda.Stream(g), torch.cuda.Stream(g)
    for _ in range(2):
        with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(g)
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(reps):
        with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if both:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(g)
    res[(g, both)] = reps * n / (time.perf_counter() - t0) / 1e9


out = {"gpus": G, "mb_per_copy": mb, "host_cpus": len(os.sched_getaffinity(0))}
for both in (False, True):
    for ngpu in sorted({1, G}):
        bar = threading.Barrier(ngpu)
        thr = [threading.Thread(target=worker, args=(g, both, bar)) for g in range(ngpu)]
        [t.start() for t in thr]; [t.join() for t in thr]
        key = ("h2d+d2h" if both else "h2d") + f"_{ngpu}gpu"
        out[key + "_GBs_per_direction_total"] = round(sum(res[(g, both)] for g in range(ngpu)), 1)
        out[key + "_GBs_per_gpu"] = [round(res[(g, both)], 1) for g in range(ngpu)]
print(json.dumps(out), flush=True)
