"""CPU, world_size 2 over gloo: the batch-sharding host logic used at N>1 GPUs (no data-path collective;
only the id all-gather and the max-reduce of the early-exit step count)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, total, L, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from parseq_b200.parallel import shard_rows, gather_ids, global_steps
    full = torch.arange(total * L, dtype=torch.int32).reshape(total, L)
    a, b = shard_rows(total, world, rank)
    out = gather_ids(full[a:b].clone(), total)
    ok = torch.equal(out, full)
    s = global_steps(5 + 3 * rank)
    ret[rank] = (ok, s, (a, b))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_rows_partition():
    from parseq_b200.parallel import shard_rows
    for total in (0, 1, 7, 512, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_rows(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_gather_ids_world2_gloo():
    world, total, L = 2, 9, 26          # ragged: 5 + 4 rows
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, total, L, ret), nprocs=world, join=True)
    assert ret[0][0] and ret[1][0]
    assert ret[0][1] == ret[1][1] == 8
    assert ret[0][2] == (0, 5) and ret[1][2] == (5, 9)
