"""2-GPU NCCL test of the row-sharded table + sharded head with the real
CUDA kernels: lookup bit-exact against the replicated table, loss against the
single-process oracle on the global batch.  Skipped with fewer than 2 GPUs."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

V, De, B, L = 30011, 64, 16, 20


def _table():
    g = torch.Generator().manual_seed(0)
    return torch.randn((V, De), generator=g) * 0.1


def _ids(rank):
    g = torch.Generator().manual_seed(10 + rank)
    ids = torch.randint(1, V, (B, L), generator=g)
    ids[:, -3:] = 0
    return ids


def _labels(rank):
    g = torch.Generator().manual_seed(20 + rank)
    T = 40 + 17 * rank
    return torch.randn((T, De), generator=g), torch.randint(1, V, (T,), generator=g)


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from transformers4rec_b200 import distributed as D
    from transformers4rec_b200 import ops
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        table = _table()
        lo, hi = D.shard_bounds(V, rank, world)
        local = table[lo:hi].contiguous().cuda()
        ids = _ids(rank)
        rows, planes = D.sharded_embedding_lookup(local, ids.cuda(), V)
        ref = table[ids.reshape(-1)]
        ok_lookup = torch.equal(rows.cpu(), ref)
        rec = (planes[0].float() + planes[1].float()).cpu()
        ok_planes = (rec[:, :De] - ref).abs().max().item() < 1e-4
        xt, y = _labels(rank)
        w_planes = ops.split_planes(local)
        row_loss, loss, T_total = D.sharded_softmax_ce(xt.cuda(), y.cuda(), local, V, w_planes=w_planes)
        xs, ys = zip(*[_labels(r) for r in range(world)])
        xg, yg = torch.cat(xs), torch.cat(ys)
        ref_rows = torch.nn.functional.cross_entropy(xg @ table.t(), yg, reduction="none")
        start = sum(x.shape[0] for x in xs[:rank])
        err_rows = (row_loss.cpu() - ref_rows[start:start + xt.shape[0]]).abs().max().item()
        err_loss = abs(loss.item() - ref_rows.mean().item())
        q.put((rank, ok_lookup, ok_planes, err_rows, err_loss, T_total == xg.shape[0]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_lookup_and_head_nccl():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_lookup, ok_planes, err_rows, err_loss, ok_T in res:
        assert ok_lookup and ok_planes and ok_T, f"rank {rank}: lookup/planes/T mismatch"
        assert err_rows < 1e-3 and err_loss < 1e-4, f"rank {rank}: loss rows {err_rows} mean {err_loss}"
