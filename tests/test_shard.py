"""CPU: ray sharding index arithmetic and the world-size-2 frame all-gather over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_index_partitions_the_frame(lib):
    from ngf_b200.render import shard_index, shard_count
    for n, block, world in ((640000, 2048, 8), (1000, 64, 3), (5, 8, 2), (4096, 32, 4), (0, 16, 2)):
        seen = torch.zeros(n, dtype=torch.int32)
        for r in range(world):
            idx = shard_index(n, block, r, world)
            assert idx.numel() == shard_count(n, block, r, world)
            seen[idx] += 1
        assert bool((seen == 1).all())


def test_unshard_inverts_shard_on_cpu(lib):
    from ngf_b200.render import shard_rays, unshard_frame, shard_count
    n, block, world = 1000, 64, 3
    frame = torch.arange(n * 4, dtype=torch.float32).view(n, 4)
    max_shard = max(shard_count(n, block, r, world) for r in range(world))
    g = torch.zeros((world, max_shard, 4))
    for r in range(world):
        s = shard_rays(frame, block, r, world)
        g[r, :s.shape[0]] = s
    assert torch.equal(unshard_frame(g, n, block, world), frame)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, block, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from ngf_b200.render import shard_rays, frame_allgather
        frame = torch.arange(n * 4, dtype=torch.float32).view(n, 4) * 0.5
        local = shard_rays(frame, block, rank, world) + 1.0        # stands in for "render my rays"
        out = frame_allgather(local, n, block)
        q.put((rank, bool(torch.equal(out, frame + 1.0))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,block", [(4096, 256), (1000, 64)])
def test_frame_allgather_world2_gloo(lib, n, block):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, block, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]
