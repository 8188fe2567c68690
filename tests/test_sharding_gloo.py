"""N>1 host logic on CPU: two gloo ranks shard frames round-robin, gather boxes to rank 0, reduce the timing."""
import importlib
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    sys.path.insert(0, ROOT)
    sh = importlib.import_module("dsvt-ai-trt_b200.sharding")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sh.frames_for_rank(n_frames, rank, world)
    K = 5
    boxes = torch.stack([torch.full((K, 9), float(f)) for f in mine])         # frame f's boxes are all == f
    valid = torch.tensor(mine, dtype=torch.int32)
    for i, f in enumerate(mine):
        assert sh.global_frame_id(rank, world, i) == f
    res = sh.gather_results(boxes, valid, dst=0)
    t = sh.reduce_max(10.0 + rank)
    # BASELINE.json configs[3]'s batched gather: one packed row per frame (boxes + count)
    packed = torch.stack([torch.cat([torch.full((K * 9,), float(f)), torch.tensor([f], dtype=torch.int32).view(torch.float32)])
                          for f in mine])
    gp = sh.gather_packed(packed, dst=0)
    if rank == 0:
        assert gp.shape == (n_frames, K * 9 + 1)
        assert gp[:, 0].tolist() == [float(f) for f in range(n_frames)]
        assert gp[:, K * 9].contiguous().view(torch.int32).tolist() == list(range(n_frames))
        q.put((res[0][:, 0, 0].tolist(), res[1].tolist(), t))
    else:
        assert res is None and gp is None
        q.put(("t", t))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_gather_and_timing():
    world, n_frames = 2, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = [o for o in out if o[0] != "t"][0]
    assert full[0] == [float(f) for f in range(n_frames)]      # global frame order restored
    assert full[1] == list(range(n_frames))
    assert all(abs(o[-1] - 11.0) < 1e-9 for o in out)          # max over ranks


def test_sharding_is_a_partition():
    sh = importlib.import_module("dsvt-ai-trt_b200.sharding")
    for world in (1, 2, 4, 8):
        seen = sorted(f for r in range(world) for f in sh.frames_for_rank(64, r, world))
        assert seen == list(range(64))
        assert all(len(sh.frames_for_rank(64, r, world)) == 64 // world for r in range(world))
