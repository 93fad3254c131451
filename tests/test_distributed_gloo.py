"""world_size-2 gloo tests (CPU) of the sharded lookup / sharded head choreography
(SURVEY §8e): routing plan, all-to-all split sizes, un-permutation and the
cross-shard log-sum-exp combination, checked against a single-process computation.
The local compute steps are test-side stand-ins (the product's are CUDA kernels)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from transformers4rec_b200 import distributed as D

V, De, B, L = 1001, 16, 6, 5


def _full_table():
    g = torch.Generator().manual_seed(0)
    return torch.randn((V, De), generator=g)


def _ids(rank):
    g = torch.Generator().manual_seed(10 + rank)
    ids = torch.randint(1, V, (B, L), generator=g)
    ids[:, -1] = 0  # padding id lives in shard 0
    return ids


def _labels(rank):
    g = torch.Generator().manual_seed(20 + rank)
    T = 4 + 3 * rank  # ragged label counts per rank
    return torch.randn((T, De), generator=g), torch.randint(1, V, (T,), generator=g)


def _head_rows(xg, yg, local_table, w_planes, lo, inv_tau, rank_tgt=None):
    logits = (xg @ local_table.t()) * inv_tau
    lse = torch.logsumexp(logits, dim=1)
    loc = yg - lo
    mine = (loc >= 0) & (loc < local_table.shape[0])
    tgt = torch.where(mine, logits.gather(1, loc.clamp(0, local_table.shape[0] - 1).unsqueeze(1)).squeeze(1),
                      torch.zeros_like(lse))
    part = torch.stack([lse, tgt], dim=1)
    if rank_tgt is None:
        return part
    # classes of THIS shard scoring above the label's global logit (ties: lower id first), the label itself excluded
    col = torch.arange(local_table.shape[0]).unsqueeze(0) + lo
    above = (logits > rank_tgt.unsqueeze(1)) | ((logits == rank_tgt.unsqueeze(1)) & (col < yg.unsqueeze(1)))
    above &= col != yg.unsqueeze(1)
    return part, above.sum(dim=1).to(torch.int32)


def _label_logit(xg, yg, local_table, lo, inv_tau):
    loc = yg - lo
    mine = (loc >= 0) & (loc < local_table.shape[0])
    rows = local_table[loc.clamp(0, local_table.shape[0] - 1)]
    return torch.where(mine, (xg * rows).sum(dim=1) * inv_tau, torch.zeros(xg.shape[0]))


def _combine(parts):
    lse = torch.logsumexp(parts[:, :, 0], dim=0)
    tgt = parts[:, :, 1].sum(dim=0)
    row = lse - tgt
    return row, row.mean().reshape(1)


def _stable_topk(scores, k):
    """(value desc, index asc) -- the order of t4r_topk; torch.topk leaves ties unspecified."""
    order = torch.sort(-scores, dim=1, stable=True).indices[:, :k]
    return scores.gather(1, order), order


def _sessions(rank):
    g = torch.Generator().manual_seed(30 + rank)
    return torch.randn((5, De), generator=g)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        table = _full_table()
        lo, hi = D.shard_bounds(V, rank, world)
        local = table[lo:hi].contiguous()
        ids = _ids(rank)
        rows, _ = D.sharded_embedding_lookup(local, ids, V, gather_rows=lambda t, i: t[i],
                                             place=lambda recv, unp: (recv[unp.long()], None))
        ok_lookup = torch.equal(rows, table[ids.reshape(-1)])
        xt, y = _labels(rank)
        row_loss, loss, T_total = D.sharded_softmax_ce(xt, y, local, V, head_rows=_head_rows, combine=_combine)
        # single-process reference over the global batch
        xs, ys = zip(*[_labels(r) for r in range(world)])
        xg, yg = torch.cat(xs), torch.cat(ys)
        ref_rows = torch.nn.functional.cross_entropy(xg @ table.t(), yg, reduction="none")
        start = sum(x.shape[0] for x in xs[:rank])
        ok_head = (torch.allclose(row_loss, ref_rows[start:start + xt.shape[0]], atol=1e-5)
                   and abs(loss.item() - ref_rows.mean().item()) < 1e-5 and T_total == xg.shape[0])
        # evaluation: cross-shard label ranks (two more all-reduces) against a single-process count
        _, loss_e, _, ranks = D.sharded_softmax_ce(xt, y, local, V, head_rows=_head_rows, combine=_combine,
                                                   want_rank=True, label_logit=_label_logit)
        full = xt @ table.t()
        tgt = full.gather(1, y.unsqueeze(1))
        ids = torch.arange(V).unsqueeze(0)
        ref_rank = (((full > tgt) | ((full == tgt) & (ids < y.unsqueeze(1)))) & (ids != y.unsqueeze(1))).sum(dim=1)
        ok_head = ok_head and torch.equal(ranks.long(), ref_rank) and abs(loss_e.item() - ref_rows.mean().item()) < 1e-5
        # module form: ShardedEmbedding with ragged (per-rank different) id counts, as the sampled-softmax head uses it
        emb = D.ShardedEmbedding.from_full(table)
        emb.gather_rows, emb.place = (lambda t, i: t[i]), (lambda recv, unp: (recv[unp.long()], None))
        some = y[: 3 + 2 * rank]
        got, _ = emb.lookup(some, ragged=True)
        ok_lookup = ok_lookup and torch.equal(got, table[some]) and emb.weight.shape[0] == hi - lo
        # serving: top-k over the sharded table == top-k over the replicated one, ties (two identical rows living in
        # different shards) resolved towards the lower item id
        table_t = table.clone()
        table_t[700] = table_t[3]
        local_t = table_t[lo:hi].contiguous()
        xs = _sessions(rank)
        xs[0] = 10.0 * table_t[3]  # rows 3 (shard 0) and 700 (shard 1) tie for the best score of session 0
        for k in (1, 7, 20):
            sc, ids = D.sharded_topk(xs, local_t, V, k, inv_tau=0.5,
                                     local_topk=lambda xa, t, wp, tau, kk: _stable_topk((xa @ t.t()) * tau, kk),
                                     merge=_stable_topk)
            ref_sc, ref_ids = _stable_topk((xs @ table_t.t()) * 0.5, k)
            ok_head = ok_head and torch.equal(ids, ref_ids) and torch.allclose(sc, ref_sc, atol=1e-6)
            ok_head = ok_head and ids[0, 0].item() == 3 and (k == 1 or ids[0, 1].item() == 700)
        q.put((rank, ok_lookup, ok_head))
    finally:
        dist.destroy_process_group()


def test_sharded_lookup_and_head_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_lookup, ok_head in res:
        assert ok_lookup, f"rank {rank}: sharded lookup differs from the replicated table"
        assert ok_head, f"rank {rank}: sharded head differs from the single-process loss"


def test_shard_bounds_and_plan():
    assert D.shard_bounds(10, 0, 4) == (0, 3) and D.shard_bounds(10, 3, 4) == (9, 10)
    assert D.shard_bounds(1_000_001, 7, 8) == (875_007, 1_000_001)
    ids_all = torch.tensor([[0, 5, 9, 3], [7, 7, 1, 2]])
    plan = D.LookupPlan(ids_all, rank=1, world=2, V=10)  # rank 1 owns rows [5, 10)
    assert plan.send_counts == [2, 2] and plan.send_local_idx.tolist() == [0, 4, 2, 2]
    assert plan.recv_counts == [2, 2]          # my ids [7,7,1,2]: two owned by rank 0, two by rank 1
    assert plan.unpermute.tolist() == [2, 3, 0, 1]
    assert torch.equal(D.owner_of(torch.tensor([0, 4, 5, 9]), 10, 2), torch.tensor([0, 0, 1, 1]))


# --------------------------------------------------------------------------- #
# model level over gloo, kernels replaced by the test doubles (tests/_ops_double.py): TabularSequenceFeatures +
# XLNet + NextItemPredictionTask over a row-sharded item table -- training loss with the 3-product and the 2-unit
# head, evaluation loss + cross-shard Recall, top-k serving -- against the single-process oracle on the GLOBAL batch.
# --------------------------------------------------------------------------- #
def _model_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import _ops_double
    import t4r_oracle as O
    from _util import make_pair, mlm_draws, synth_batch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _ops_double.install_plain()
        cards, dims = {"item_id/list": 2001}, {"item_id/list": 32}
        Bm, Lm, K = 6, 10, 5
        oracle, model = make_pair(cards, dims, "item_id/list", (), 32, 2, 1, Lm, weight_scale=0.08, device="cpu")
        inputs = model.heads[0].body[0]
        inputs.categorical_module.shard_item_table()
        task = model.heads[0].prediction_task_dict["next-item"]
        batches = [synth_batch(Bm, Lm, cards, seed=100 + r) for r in range(world)]
        us = [mlm_draws(Bm, Lm, seed=200 + r) for r in range(world)]
        gb = {k: torch.cat([b[k] for b in batches]) for k in batches[0]}
        gdraws = {k: torch.cat([u[1][k] for u in us]) for k in us[0][1]}
        res = {}
        with torch.no_grad():
            inputs.masking.set_draws(us[rank][0])
            ref = oracle(gb, training=True, draws=gdraws)
            res["train3"] = abs(model(batches[rank], training=True)["loss"].item() - ref["loss"].item())
            task.nprod = 2   # sharded full softmax through the 2-unit product (emulated from the packed operands)
            res["train2"] = abs(model(batches[rank], training=True)["loss"].item() - ref["loss"].item())
            ref_e = oracle(gb, training=False, testing=True)
            out_e = model(batches[rank], training=False, testing=True)
            res["eval"] = abs(out_e["loss"].item() - ref_e["loss"].item())
            mine = slice(rank * Bm, (rank + 1) * Bm)
            ref_rec = O.recall_at_mean([1, 5, 20], ref_e["predictions"][mine], ref_e["labels"][mine])
            import transformers4rec_b200.torch as tr
            m = tr.RecallAt(top_ks=[1, 5, 20], labels_onehot=True)
            m.update_from_ranks(out_e.row_rank, None)
            res["recall"] = (m.metric_mean[-1] - ref_rec).abs().max().item()
            # serving
            short = {k: torch.nn.functional.pad(v[:, :-1], (0, 1)) for k, v in batches[rank].items()}
            x, _, _ = oracle.input_block(short, False, False)
            h = O.hf_encoder_forward(oracle.transformer, x)
            last = (short["item_id/list"] != 0).sum(1)
            ref_scores = h[torch.arange(Bm), last] @ oracle.item_table().t()
            model.top_k = K
            s, i = model(short, training=False, testing=False)
            rs, ri = torch.topk(ref_scores, K)
            res["topk_scores"] = (s - rs).abs().max().item()
            res["topk_ids"] = (i == ri).float().mean().item()
            try:
                model.top_k = None
                model(short, training=False, testing=False)
                res["scores_error"] = "no error"
            except NotImplementedError as exc:
                res["scores_error"] = str(exc)
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_model_over_sharded_item_table_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_model_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, r in res:
        assert r["train3"] < 1e-4 and r["train2"] < 1e-3 and r["eval"] < 1e-4 and r["recall"] < 1e-6, (rank, r)
        assert r["topk_scores"] < 1e-4 and r["topk_ids"] == 1.0, (rank, r)
        assert "top_k" in r["scores_error"], (rank, r)


# --------------------------------------------------------------------------- #
# training step over the row-sharded table (N3: "sharded-table grads via the transposed all-to-all"): every rank
# back-propagates its own sessions; gradients against torch autograd of the single-process oracle on the GLOBAL batch
# --------------------------------------------------------------------------- #
def _train_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import _ops_double
    from _util import make_pair, mlm_draws, synth_batch
    from test_host_training_cpu import _pairs
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _ops_double.install_plain()
        from transformers4rec_b200.training import FusedTrainingStep
        cards, dims = {"item_id/list": 1201, "category/list": 23}, {"item_id/list": 32, "category/list": 32}
        Bm, Lm = 5, 8
        oracle, model = make_pair(cards, dims, "item_id/list", (), 32, 2, 1, Lm, weight_scale=0.08, device="cpu")
        oracle.train(False)
        inputs = model.heads[0].body[0]
        full_item = inputs.categorical_module.embedding_tables["item_id/list"].weight.detach().clone()
        inputs.categorical_module.shard_item_table()
        table = inputs.categorical_module.embedding_tables["item_id/list"]
        batches = [synth_batch(Bm, Lm, cards, seed=100 + r) for r in range(world)]
        us = [mlm_draws(Bm, Lm, seed=200 + r) for r in range(world)]
        gb = {k: torch.cat([b[k] for b in batches]) for k in batches[0]}
        gdraws = {k: torch.cat([u[1][k] for u in us]) for k in us[0][1]}
        ref = oracle(gb, training=True, draws=gdraws)
        ref["loss"].backward()
        inputs.masking.set_draws(us[rank][0])
        step = FusedTrainingStep(model, head_chunk=256)
        for p in model.parameters():
            p.grad = None
        loss = step.forward(batches[rank])
        step.backward()
        res = {"loss": abs(loss.item() - ref["loss"].item()), "worst": 0.0, "n": 0}
        og = oracle.tables["item_id__list"].weight.grad
        res["table"] = (table.weight.grad - og[table.lo:table.hi]).abs().max().item() / max(1.0, og.abs().max().item())
        res["table_rows"] = tuple(table.weight.grad.shape) == (table.hi - table.lo, 32)
        for name, po, pm in _pairs(oracle, model):
            if "item_id" in name or (po.grad is None and pm.grad is None):
                continue
            err = (pm.grad - po.grad.reshape(pm.grad.shape)).abs().max().item() / max(1.0, po.grad.abs().max().item())
            res["worst"] = max(res["worst"], err)
            res["n"] += 1
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_training_step_over_sharded_item_table_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, r in res:
        assert r["loss"] < 1e-4 and r["table"] < 3e-4 and r["table_rows"] and r["worst"] < 3e-4 and r["n"] >= 15, (rank, r)


def _ddp_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import _ops_double
    from _util import make_pair, mlm_draws, synth_batch
    from test_host_training_cpu import _pairs
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _ops_double.install_plain()
        from torch.nn.parallel import DistributedDataParallel as DDP
        cards, dims = {"item_id/list": 401, "category/list": 23}, {"item_id/list": 32, "category/list": 32}
        Bm, Lm = 5, 8
        oracle, model = make_pair(cards, dims, "item_id/list", (), 32, 2, 1, Lm, weight_scale=0.08, device="cpu")
        oracle.train(False)
        batches = [synth_batch(Bm, Lm, cards, seed=100 + r) for r in range(world)]
        us = [mlm_draws(Bm, Lm, seed=200 + r) for r in range(world)]
        # what DDP computes: the mean over the ranks of every rank's own mean-loss gradient
        want = {}
        for r in range(world):
            for p in oracle.parameters():
                p.grad = None
            oracle(batches[r], training=True, draws=us[r][1])["loss"].backward()
            for n, p in oracle.named_parameters():
                if p.grad is not None:
                    want[n] = want.get(n, 0) + p.grad.detach().clone() / world
        names = {id(p): n for n, p in oracle.named_parameters()}
        model.heads[0].body[0].masking.set_draws(us[rank][0])
        model.enable_fused_training(head_chunk=256)
        ddp = DDP(model, find_unused_parameters=True)   # as the reference's multi-GPU recipe needs (docs/source/multi_gpu_train.md)
        res = {"worst": 0.0, "n": 0, "steps": 0}
        for it in range(2):                              # a second step proves the reducer finished the first one cleanly
            for p in model.parameters():
                p.grad = None
            out = ddp(batches[rank], training=True)
            out["loss"].backward()
            res["steps"] += 1
        for name, po, pm in _pairs(oracle, model):
            key = names[id(po)]
            if key not in want and pm.grad is None:
                continue
            err = (pm.grad - want[key].reshape(pm.grad.shape)).abs().max().item() / max(1.0, want[key].abs().max().item())
            res["worst"] = max(res["worst"], err)
            res["n"] += 1
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_fused_training_under_distributed_data_parallel_world2():
    """The reference's multi-GPU training mode is torch DDP around the model (docs/source/multi_gpu_train.md).
    With ``enable_fused_training`` the same wrapper works: the fused step's autograd node feeds DDP's gradient hooks, and
    every parameter ends with the mean over the ranks of the oracle's per-rank gradients."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, r in res:
        assert r["steps"] == 2 and r["worst"] < 3e-4 and r["n"] >= 15, (rank, r)


def test_sharded_model_and_training_step_world3():
    """The same workers at world size 3 (shards of uneven size: 1201 = 401 + 400 + 400 rows): model forward / eval /
    top-k over the sharded table and the sharded training step."""
    ctx = mp.get_context("spawn")
    for worker, base in ((_model_worker, 37500), (_train_worker, 39500)):
        q = ctx.Queue()
        port = base + (os.getpid() % 2000)
        procs = [ctx.Process(target=worker, args=(r, 3, port, q)) for r in range(3)]
        for p in procs:
            p.start()
        res = [q.get(timeout=300) for _ in procs]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        for rank, r in res:
            if worker is _model_worker:
                assert r["train3"] < 1e-4 and r["train2"] < 1e-3 and r["eval"] < 1e-4 and r["recall"] < 1e-6, (rank, r)
                assert r["topk_scores"] < 1e-4 and r["topk_ids"] == 1.0, (rank, r)
            else:
                assert r["loss"] < 1e-4 and r["table"] < 3e-4 and r["table_rows"] and r["worst"] < 3e-4, (rank, r)
