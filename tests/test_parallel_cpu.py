"""world_size-2 gloo test of the N>1 host logic (scene sharding, max-over-ranks timing, observation
gather); the step itself has no collective."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from img_env_b200.parallel import ObservationGatherer, gather_observations, owner_of, reduce_max, shard_scenes


def test_shards_partition_all_scenes():
    for total in (1, 7, 16, 4096, 4097):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                s, n = shard_scenes(total, world, r)
                seen += list(range(s, s + n))
                for sc in range(s, s + n):
                    assert owner_of(sc, total, world) == r
            assert seen == list(range(total))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, n = shard_scenes(6, world, rank)
    out = {"lasers": torch.full((n, 2, 5), float(rank)), "is_collisions": torch.arange(start, start + n, dtype=torch.int8)[:, None].repeat(1, 2)}
    g = gather_observations(out, dst=0)
    # persistent gatherer, unequal shards (7 scenes -> 4 + 3), delivered to rank 1, called twice
    s7, n7 = shard_scenes(7, world, rank)
    o7 = {"step_ds": torch.arange(s7, s7 + n7, dtype=torch.float32)[:, None].repeat(1, 3)}
    gat = ObservationGatherer(o7, dst=1)
    for rep in range(2):
        o7["step_ds"] += 100.0
        g7 = gat(o7)
        if rank == 1:
            assert g7["step_ds"][:, 0].tolist() == [100.0 * (rep + 1) + i for i in range(7)]
            assert gat.bytes_to_learner == 4 * 3 * 4
        else:
            assert g7 is None
    m = reduce_max(10.0 + rank)
    if rank == 0:
        q.put((g["lasers"][:, 0, 0].tolist(), g["is_collisions"][:, 0].tolist(), m))
    else:
        assert g is None
        q.put(m)
    dist.destroy_process_group()


def test_gloo_world2_gather_and_max():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    full = [r for r in res if isinstance(r, tuple)][0]
    assert full[0] == [0.0, 0.0, 0.0, 1.0, 1.0, 1.0]
    assert full[1] == [0, 1, 2, 3, 4, 5]
    assert full[2] == 11.0
    assert [r for r in res if not isinstance(r, tuple)][0] == 11.0
