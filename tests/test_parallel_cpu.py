"""CPU (`gloo`, world_size 2): the N>1 host logic — bucket planning over the flat g_hat buffer, asynchronous averaging,
the trailing small-gradient message and batch sharding — without any GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tinyedm_b200.parallel import Bucket, GradReducer, plan_buckets, shard_slice


def test_plan_buckets_tiles_the_buffer_from_the_end():
    # 5 units completed back to front over a 1000-element buffer
    units = [(900, 1000), (600, 900), (590, 600), (200, 590), (0, 200)]
    b = plan_buckets(units, bucket_elems=300)
    assert [(x.start, x.end, x.ready_after) for x in b] == [(600, 1000, 1), (200, 600, 3), (0, 200, 4)]
    covered = sorted((x.start, x.end) for x in b)
    assert covered[0][0] == 0 and covered[-1][1] == 1000 and all(a[1] == c[0] for a, c in zip(covered, covered[1:]))
    assert [x.ready_after for x in plan_buckets(units, 1)] == [0, 1, 2, 3, 4]       # one bucket per unit
    assert len(plan_buckets(units, 10 ** 9)) == 1                                    # everything in the trailing bucket
    with pytest.raises(ValueError):
        plan_buckets([(900, 1000), (500, 800)], 10)                                  # gap


def test_plan_buckets_lone_tail_keeps_the_last_message_small():
    units = [(900, 1000), (600, 900), (590, 600), (200, 590), (0, 200)]
    b = plan_buckets(units, 300, lone_tail=True)
    assert [(x.start, x.end, x.ready_after) for x in b] == [(600, 1000, 1), (200, 600, 3), (0, 200, 4)]
    b = plan_buckets(units, 10 ** 9, lone_tail=True)      # even with one huge bucket the last unit travels alone
    assert [(x.start, x.end, x.ready_after) for x in b] == [(200, 1000, 3), (0, 200, 4)]
    assert [(x.start, x.end) for x in plan_buckets(units[:1], 10, lone_tail=True)] == [(900, 1000)]


def test_cifar_engine_units_tile_the_ghat_buffer():
    import tinyedm_b200 as T
    from oracle import edm2_oracle as O
    from tests.helpers import spec_kwargs
    den = T.Denoiser(**spec_kwargs(O.CIFAR10["denoiser"]))
    eng = den.engine
    eng.bank.materialise(torch.device("cpu"))         # buffer bookkeeping only; no kernel is launched
    units = eng.ghat_units_backward_order()
    assert len(units) == 22 and units[-1][0] == 0
    for a, b in zip(units, units[1:]):
        assert b[1] == a[0]
    buckets = plan_buckets(units, int(25 * (1 << 20) / 4))
    assert 4 <= len(buckets) <= 8                      # 142.5 MB of fp32 g_hat in ~25 MiB buckets
    assert sum(b.end - b.start for b in buckets) == eng.bank._ghat_flat.numel()


@pytest.mark.parametrize("n,world", [(128, 8), (50000, 8), (7, 4), (3, 8), (128, 1)])
def test_shard_slice_partitions_exactly(n, world):
    seen = []
    for r in range(world):
        s = shard_slice(n, r, world)
        seen += list(range(n))[s]
    assert seen == list(range(n))
    sizes = [len(range(n)[shard_slice(n, r, world)]) for r in range(world)]
    assert max(sizes) - min(sizes) <= 1


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)
        n = 1000
        flat = torch.randn(n)
        mine = flat.clone()
        units = [(900, 1000), (600, 900), (590, 600), (200, 590), (0, 200)]
        plan = plan_buckets(units, 300)
        red = GradReducer()
        nxt = 0
        for i in range(len(units)):                 # what DistributedEDM.unit_done does
            while nxt < len(plan) and plan[nxt].ready_after <= i:
                red.launch(flat[plan[nxt].start:plan[nxt].end])
                nxt += 1
        red.wait()
        gathered = [torch.empty(n) for _ in range(world)]
        dist.all_gather(gathered, mine)
        expect = torch.stack(gathered).mean(0)
        ok = torch.allclose(flat, expect, atol=1e-6)
        # every rank must hold the same reduced buffer
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        ok = ok and torch.equal(ref, flat)
        # sharded sampling bookkeeping: slices of all ranks re-assemble the batch
        part = torch.arange(11)[shard_slice(11, rank, world)]
        parts = [None] * world
        dist.all_gather_object(parts, part.tolist())
        ok = ok and sum(parts, []) == list(range(11))
        out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_bucketed_average_world_size_2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    results = dict(out.get(timeout=5) for _ in range(2))
    assert results == {0: True, 1: True}
