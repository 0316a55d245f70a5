"""CPU (gloo, world size 2): the data-parallel gradient synchronisation protocol of ctunet_b200.parallel and the
collective-free sharding of independent units.  The reference's counterpart is single-process nn.DataParallel
(ctunet/pytorch/Model.py:481-486): averaged gradients over equal per-replica batches."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _Toy(nn.Module):
    """Same naming pattern as the generic UNet: `cblock.*` never receives a gradient (models.py:241)."""

    def __init__(self):
        super().__init__()
        self.d_blocks = nn.ModuleList([nn.Conv3d(2, 4, 3, bias=False), nn.Conv3d(4, 4, 3, bias=False)])
        self.cblock = nn.Conv3d(4, 8, 3, bias=False)
        self.last_conv = nn.Conv3d(4, 3, 1)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ctunet_b200.parallel import GradSync, shard_range
        torch.manual_seed(0)
        net = _Toy()
        sync = GradSync(net, n_buckets=2)
        assert sync.world == world
        live = [p for n, p in net.named_parameters() if not n.startswith("cblock")]
        assert sync.buffer_for(net.cblock.weight) is None
        results = []
        for step in range(2):                                  # the protocol must be re-armable
            expect = {}
            for i, p in enumerate(reversed(live)):             # production order of a backward pass
                buf = sync.buffer_for(p)
                assert buf is not None and buf.shape == p.shape
                g = torch.full_like(p, float(rank + 1) * (i + 1) + step)
                buf.copy_(g)
                sync.delivered(p)
                expect[id(p)] = sum(float(r + 1) * (i + 1) + step for r in range(world)) / world
            sync.finish()
            ok = all(torch.allclose(p.grad, torch.full_like(p, expect[id(p)])) for p in live)
            ok = ok and net.cblock.weight.grad is None
            ok = ok and all(p.grad.data_ptr() == sync.buffer_for(p).data_ptr() for p in live)   # views of the flat buffer
            results.append(ok)
        # a gradient that never arrives must be reported, not silently skipped
        sync.buffer_for(live[0]).zero_()
        sync.delivered(live[-1])
        try:
            sync.finish()
            results.append(False)
        except RuntimeError:
            results.append(True)
        # deferred mode (the CUDA-graph step): nothing is reduced while gradients arrive, finish() reduces once
        dsync = GradSync(net, n_buckets=2, deferred=True)
        for i, p in enumerate(reversed(live)):
            dsync.buffer_for(p).copy_(torch.full_like(p, float(10 * rank + i)))
            dsync.delivered(p)
        assert not dsync._handles
        dsync.finish()
        results.append(all(torch.allclose(p.grad, torch.full_like(p, sum(10.0 * r + i for r in range(world)) / world))
                           for i, p in enumerate(reversed(live))))
        lo, hi = shard_range(7, rank, world)
        q.put((rank, results, (lo, hi)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gradsync_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    out.sort()
    for rank, results, _ in out:
        assert all(results), (rank, results)
    assert [o[2] for o in out] == [(0, 4), (4, 7)]            # contiguous, balanced, collective-free partition


def test_shard_range_partitions_exactly():
    from ctunet_b200.parallel import shard_range
    for n_items in (0, 1, 5, 32, 33):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n_items, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n_items
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
