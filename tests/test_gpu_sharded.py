"""Multi-process tests of the row-sharded table + sharded head with the real CUDA kernels: lookup bit-exact
against the replicated table, loss / Recall@k against the single-process oracle on the GLOBAL batch.  Both formulations
run: ``peer`` = NVLink peer memory (csrc/t4r_peer.cu, the default on GPUs) and ``nccl`` = all-gather + all-to-all
(``T4R_PEER=0``).  World size 2 needs 2 GPUs (skipped otherwise); world size 1 runs the same code path -- process group,
IPC export, windows, device-side counts -- on a single GPU, so the driver's one-GPU box covers it too."""
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


def _worlds():
    return [pytest.param(1, id="world1"),
            pytest.param(2, id="world2", marks=pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs"))]


def _run(target, world, port, extra=(), timeout=600):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=target, args=(r, world, port, q) + tuple(extra)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=timeout) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


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
        # the same lookup over peer memory (the module's default on GPUs): bit-exact too, incl. a device-side count
        emb = D.ShardedEmbedding.from_full(table.cuda())
        assert emb.peer_view() is not None, "peer memory should be available between the GPUs of one box"
        prow, pplanes = emb.lookup(ids.cuda())
        ok_lookup = ok_lookup and torch.equal(prow.cpu(), ref) and torch.equal(pplanes.cpu(), planes.cpu())
        cnt = torch.tensor([77], dtype=torch.int32, device="cuda")
        crow, _ = emb.lookup(ids.cuda(), count=cnt)
        ok_lookup = ok_lookup and torch.equal(crow[:77].cpu(), ref[:77]) and not crow[77:256].any()
        q.put((rank, ok_lookup, ok_planes, err_rows, err_loss, T_total == xg.shape[0]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", _worlds())
def test_sharded_lookup_and_head_nccl(world):
    res = _run(_worker, world, 29600 + (os.getpid() % 1000) + 7 * world, timeout=300)
    for rank, ok_lookup, ok_planes, err_rows, err_loss, ok_T in res:
        assert ok_lookup and ok_planes and ok_T, f"rank {rank}: lookup/planes/T mismatch"
        assert err_rows < 1e-3 and err_loss < 1e-4, f"rank {rank}: loss rows {err_rows} mean {err_loss}"


# --------------------------------------------------------------------------- #
# model level: TabularSequenceFeatures + XLNet + NextItemPredictionTask over a row-sharded item table
# (BASELINE configs 4-5 at test size), each rank with its own sessions; reference = the single-process
# CPU oracle on the GLOBAL batch ("multi-GPU parity is defined against the single-process
# oracle on the global batch")
# --------------------------------------------------------------------------- #
def _model_worker(rank, world, port, q, sampled, peer):
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    os.environ["T4R_PEER"] = "1" if peer else "0"
    import torch.distributed as dist

    import t4r_oracle as O
    from _util import make_pair, mlm_draws, synth_batch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cards, dims = {"item_id/list": 12001}, {"item_id/list": 64}
        B, L, S = 24, 20, 300
        oracle, model = make_pair(cards, dims, "item_id/list", (), 64, 4, 1, L, sampled=sampled, max_n_samples=S,
                                  weight_scale=0.08, device=f"cuda:{rank}")
        inputs = model.heads[0].body[0]
        inputs.categorical_module.shard_item_table()
        task = model.heads[0].prediction_task_dict["next-item"]
        batches = [synth_batch(B, L, cards, seed=100 + r) for r in range(world)]
        us = [mlm_draws(B, L, seed=200 + r) for r in range(world)]
        gb = {k: torch.cat([b[k] for b in batches]) for k in batches[0]}
        gdraws = {k: torch.cat([u[1][k] for u in us]) for k in us[0][1]}
        res = {}
        with torch.no_grad():
            inputs.masking.set_draws(us[rank][0].cuda())
            if sampled:
                torch.manual_seed(3)
                raw = torch.multinomial(oracle.dist, 2 * S, replacement=True)
                task.set_negative_draws(raw.cuda())
                ref = oracle(gb, training=True, draws=gdraws, neg_samples=O.negatives_from_draws(raw, S))
                out = model({k: v.cuda() for k, v in batches[rank].items()}, training=True)
                res["train"] = abs(out["loss"].item() - ref["loss"].item())
            else:
                ref = oracle(gb, training=True, draws=gdraws)
                out = model({k: v.cuda() for k, v in batches[rank].items()}, training=True)
                res["train"] = abs(out["loss"].item() - ref["loss"].item())
                # evaluation: loss over the global batch, Recall@k of THIS rank's sessions from cross-shard ranks
                ref_e = oracle(gb, training=False, testing=True)
                out_e = model({k: v.cuda() for k, v in batches[rank].items()}, training=False, testing=True)
                res["eval"] = abs(out_e["loss"].item() - ref_e["loss"].item())
                mine = slice(rank * B, (rank + 1) * B)  # eval: one label (the last item) per session
                ref_rec = O.recall_at_mean([1, 5, 20], ref_e["predictions"][mine], ref_e["labels"][mine])
                import transformers4rec_b200.torch as tr
                m = tr.RecallAt(top_ks=[1, 5, 20], labels_onehot=True)
                m.update_from_ranks(out_e.row_rank, out_e.count)
                res["recall"] = (m.metric_mean[-1].cpu() - ref_rec).abs().max().item()
        res["peer"] = inputs.categorical_module.embedding_tables["item_id/list"].peer_view() is not None
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", _worlds())
@pytest.mark.parametrize("peer", [True, False], ids=["peer", "nccl"])
@pytest.mark.parametrize("sampled", [False, True], ids=["full", "sampled"])
def test_model_over_sharded_item_table_nccl(sampled, peer, world):
    port = 29700 + (os.getpid() % 1000) + (50 if sampled else 0) + (11 if peer else 0) + 3 * world
    res = _run(_model_worker, world, port, extra=(sampled, peer))
    for rank, r in res:
        assert r["peer"] == peer, (rank, r)
        assert r["train"] < 1e-3, (rank, r)
        if not sampled:
            assert r["eval"] < 1e-3 and r["recall"] < 1e-6, (rank, r)


# --------------------------------------------------------------------------- #
# serving over the sharded table: Model.top_k through per-shard top-k + one candidate exchange
# --------------------------------------------------------------------------- #
def _topk_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import torch.distributed as dist

    import t4r_oracle as O
    from _util import make_pair, synth_batch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cards, dims = {"item_id/list": 12001}, {"item_id/list": 64}
        B, L, K = 24, 20, 10
        oracle, model = make_pair(cards, dims, "item_id/list", (), 64, 4, 1, L, weight_scale=0.08, device=f"cuda:{rank}")
        model.heads[0].body[0].categorical_module.shard_item_table()
        batch = synth_batch(B, L - 1, cards, seed=300 + rank)  # room for the extra [MASK] position
        batch = {k: torch.nn.functional.pad(v, (0, 1)) for k, v in batch.items()}
        with torch.no_grad():
            x, _, _ = oracle.input_block(batch, False, False)
            h = O.hf_encoder_forward(oracle.transformer, x)
            last = (batch["item_id/list"] != 0).sum(1)
            ref_scores = h[torch.arange(B), last] @ oracle.item_table().t()
            model.top_k = K
            s, i = model({k: v.cuda() for k, v in batch.items()}, training=False, testing=False)
        rs, ri = torch.topk(ref_scores, K)
        q.put((rank, (s.cpu() - rs).abs().max().item(), (i.cpu() == ri).float().mean().item(),
               bool((s[:, :-1] >= s[:, 1:]).all())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", _worlds())
def test_topk_over_sharded_item_table_nccl(world):
    res = _run(_topk_worker, world, 29800 + (os.getpid() % 1000) + 5 * world)
    for rank, err, agree, sorted_desc in res:
        assert err < 1e-3 and agree > 0.98 and sorted_desc, (rank, err, agree, sorted_desc)
