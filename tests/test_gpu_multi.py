"""2 GPUs, one process per GPU over NCCL: the batch is sharded by contiguous row blocks, every rank decodes its block,
the decoded token ids are all-gathered (the only collective of the path) and must equal the single-GPU run
of the whole batch.  Skipped on a one-GPU box."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, total, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from parseq_b200.config import make_config
    from parseq_b200.factory import create_model
    from parseq_b200.parallel import shard_rows, gather_ids, global_steps
    from parseq_b200.weights import init_state_dict, synth_images
    cfg = make_config("parseq")
    m = create_model("parseq", decode_ar=True, refine_iters=1)
    m.model.load_state_dict(init_state_dict(cfg, 0))
    m = m.eval().to(f"cuda:{rank}")
    x = synth_images(cfg, total, 4321)
    a, b = shard_rows(total, world, rank)
    with torch.inference_mode():
        _, ids = m.model.forward(m.tokenizer, x[a:b].cuda(), None, return_ids=True)
        gathered = gather_ids(ids, total)
        S = global_steps(26, ids.device)
        ok = True
        if rank == 0:
            _, full = m.model.forward(m.tokenizer, x.cuda(), None, return_ids=True)
            # rows are batch-invariant within a kernel regime; both runs are small batches (unfused LayerNorm path)
            ok = bool(torch.equal(gathered, full))
    ret[rank] = (ok, tuple(gathered.shape), S, (a, b))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_decode_ids_all_gather_nccl():
    import torch.multiprocessing as mp
    world, total = 2, 75          # ragged: 38 + 37 rows
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29700 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(world, port, total, ret), nprocs=world, join=True)
    assert ret[0][0], "all-gathered ids of the sharded run differ from the single-GPU run"
    assert ret[0][1] == ret[1][1] == (total, 26)
    assert ret[0][2] == ret[1][2] == 26
    assert ret[0][3] == (0, 38) and ret[1][3] == (38, 75)
