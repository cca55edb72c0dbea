#!/usr/bin/env python3
"""Small-batch latency of ldpc_decode / ldpc_encode (device resident): the reference is a streaming FPGA design that takes one
block at a time, so the cost of ONE call matters beside the batch throughput.  Per batch size: wall time per call with a
synchronise after every call (latency), wall time per call of a queue of calls (launch-bound throughput), and the same with
the call captured once in a CUDA graph and replayed (cudaGraphLaunch; the library's calls are plain stream work and can be
captured by the caller)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldpc_erasure_codes_b200.codec import LdpcCodec, fill_random

S = int(os.environ.get("S", "64")); P = int(os.environ.get("P", "13")); mode = os.environ.get("MODE", "peel")
sizes = [int(x) for x in os.environ.get("SIZES", "1,8,64,512,4096,32768").split(",")]
codec = LdpcCodec(code=int(os.environ.get("CODE", "1")), symbol_bytes=S, device=0, max_batch=max(sizes))
n, k = codec.n, codec.k
it = 10 if mode == "hybrid" else 50
for B in sizes:
    info = torch.empty((B, k, S), dtype=torch.uint8, device="cuda"); fill_random(info, 1)
    cw = codec.encode(info)
    rx = cw.clone(); mask = codec.gen_erasures(B, 99, P=P, payload=rx)
    out = torch.empty((B, k, S), dtype=torch.uint8, device="cuda"); fail = torch.empty((B,), dtype=torch.uint8, device="cuda")
    run = lambda: codec.decode(rx, mask, max_iter=it, mode=mode, out=out, fail=fail)
    for _ in range(5): run()
    torch.cuda.synchronize()
    reps = 200 if B <= 4096 else 30
    t0 = time.perf_counter()
    for _ in range(reps):
        run(); torch.cuda.synchronize()
    lat = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps): run()
    torch.cuda.synchronize()
    thr = (time.perf_counter() - t0) / reps
    rec = dict(op="decode/" + mode, B=B, S=S, P=P, us_per_call_sync=round(lat * 1e6, 1), us_per_call_queued=round(thr * 1e6, 1),
               gbit_s_queued=round(B * k * S * 8 / thr / 1e9, 1))
    ref = out.clone()
    try:
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            run()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=st):
                run()
        torch.cuda.synchronize()
        out.zero_()
        graph.replay(); torch.cuda.synchronize()
        assert torch.equal(out, ref), "graph replay output differs"
        t0 = time.perf_counter()
        for _ in range(reps):
            graph.replay(); torch.cuda.synchronize()
        glat = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(reps): graph.replay()
        torch.cuda.synchronize()
        gthr = (time.perf_counter() - t0) / reps
        rec.update(graph_us_per_call_sync=round(glat * 1e6, 1), graph_us_per_call_queued=round(gthr * 1e6, 1),
                   graph_gbit_s_queued=round(B * k * S * 8 / gthr / 1e9, 1))
    except Exception as e:  # noqa: BLE001
        rec["graph"] = "not capturable: " + str(e).splitlines()[0][:200]
    print(json.dumps(rec), flush=True)
