"""N > 1 host logic on CPU: two gloo ranks shard a frame range, run the (CPU) checker on their shard with GLOBAL
frame indices, and the gathered result equals the unsharded run -- the property the GPU sharding relies on."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ldpc_erasure_codes_b200 import sharding


def test_shard_range_covers_everything():
    for total in (0, 1, 7, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(total, g, world) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
    assert sharding.weak_frame_base(1 << 20, 3) == 3 << 20
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    code = orc.Code.builtin(1)
    b, e = sharding.shard_range(total, rank, world)
    rng = np.random.default_rng(1234)                       # same stream on every rank: the full info tensor
    info = rng.integers(0, 256, (total, code.k, 16), dtype=np.uint8)[b:e]
    cw = orc.encode(code, info)
    flags = orc.gen_erasures_iid(code.n, 99, e - b, P=12, frame0=b)   # global frame index -> sharding independent
    cw[flags == 1] = 0
    res = orc.decode(code, cw, flags, max_iter=50)
    stats = sharding.reduce_stats({"frames": e - b, "ldpc_errors": int(res["fail_sys"].sum())})
    slowest = sharding.reduce_max(10.0 + rank)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), out=res["out"], fail=res["fail_sys"], frames=stats["frames"],
             errors=stats["ldpc_errors"], slowest=slowest)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_single_process(tmp_path):
    from oracle import oracle as orc
    total, world = 48, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    code = orc.Code.builtin(1)
    rng = np.random.default_rng(1234)
    info = rng.integers(0, 256, (total, code.k, 16), dtype=np.uint8)
    cw = orc.encode(code, info)
    flags = orc.gen_erasures_iid(code.n, 99, total, P=12)
    cw[flags == 1] = 0
    ref = orc.decode(code, cw, flags, max_iter=50)
    assert np.array_equal(np.concatenate([p["out"] for p in parts]), ref["out"])
    assert np.array_equal(np.concatenate([p["fail"] for p in parts]), ref["fail_sys"])
    for p in parts:
        assert int(p["frames"]) == total and int(p["errors"]) == int(ref["fail_sys"].sum())
        assert float(p["slowest"]) == 11.0


# ---- the same sharding, run by the LIBRARY on GPUs (one process per rank, rank r on GPU r % ngpu) -----------------
def _gpu_worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ldpc_erasure_codes_b200.codec import LdpcCodec, fill_random
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    codec = LdpcCodec(code=1, symbol_bytes=16, device=dev, max_batch=64)
    b, e = sharding.shard_range(total, rank, world)
    info = torch.empty((e - b, codec.k, 16), dtype=torch.uint8, device=f"cuda:{dev}")
    fill_random(info, seed=1234, block0=b * codec.k)                       # 16-byte blocks: one per symbol, global index
    cw = codec.encode(info)
    mask = codec.gen_erasures(e - b, 99, P=12, frame0=b, payload=cw)       # global frame index -> sharding independent
    out, fail = codec.decode(cw, mask, max_iter=50)
    stats = sharding.reduce_stats(codec.stats())
    np.savez(os.path.join(out_dir, f"g{rank}.npz"), info=info.cpu().numpy(), out=out.cpu().numpy(), fail=fail.cpu().numpy(),
             frames=stats["frames"], errors=stats["ldpc_errors"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_rank_gpu_shards_equal_oracle(tmp_path):
    from oracle import oracle as orc
    total, world = 150, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_gpu_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"g{r}.npz") for r in range(world)]
    code = orc.Code.builtin(1)
    info = np.concatenate([p["info"] for p in parts])
    cw = orc.encode(code, info)
    flags = orc.gen_erasures_iid(code.n, 99, total, P=12)
    cw[flags == 1] = 0
    ref = orc.decode(code, cw, flags, max_iter=50)
    assert np.array_equal(np.concatenate([p["out"] for p in parts]), ref["out"])
    assert np.array_equal(np.concatenate([p["fail"] for p in parts]), ref["fail_sys"])
    for p in parts:
        assert int(p["frames"]) == total and int(p["errors"]) == int(ref["fail_sys"].sum())
    # the shards' payload generator is sharding independent too: one rank would have produced the same info
    import torch as _t
    from ldpc_erasure_codes_b200.codec import fill_random
    whole = _t.empty((total, code.k, 16), dtype=_t.uint8, device="cuda")
    fill_random(whole, seed=1234, block0=0)
    assert np.array_equal(whole.cpu().numpy(), info)
