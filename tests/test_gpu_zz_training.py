"""GPU tests of the training step (transformers4rec_b200/training.py, csrc/t4r_train.cu).  CPU-side
evidence: the composition matches torch autograd of the oracle graph with kernel doubles AND with the kernels' real
per-item code on their host twins (tests/test_host_training_cpu.py, tests/test_abi_and_host.py); first run on a B200 in
round 2."""
import os

import pytest
import torch

import _ops_double as DD
from _util import make_pair, mlm_draws, synth_batch
from test_host_training_cpu import _oracle_grads, _pairs

pytestmark = [pytest.mark.gpu]


def test_training_primitives_device_vs_host_twin():
    from transformers4rec_b200 import _lib, ops
    g = torch.Generator().manual_seed(41)
    H = ops.host_twin
    close = lambda a, b, tol=1e-5: (a.cpu() - b).abs().max().item() <= tol * max(1.0, b.abs().max().item())
    x, dy = torch.randn(300, 96, generator=g), torch.randn(300, 96, generator=g)
    assert torch.equal(ops.transpose(x.cuda()).cpu(), H("transpose")(x))
    for kind in (_lib.ACT_GELU, _lib.ACT_RELU):
        assert close(ops.act_fwd(kind, x.cuda()), H("act_fwd")(kind, x))
        assert close(ops.act_bwd(kind, x.cuda(), dy.cuda()), H("act_bwd")(kind, x, dy))
    assert close(ops.col_sum(x.cuda()), H("col_sum")(x), 1e-4)
    gam, bet = torch.rand(96, generator=g) + 0.5, torch.randn(96, generator=g)
    assert close(ops.layer_norm_fwd(x.cuda(), gam.cuda(), bet.cuda(), 0.03), H("layer_norm_fwd")(x, gam, bet, 0.03))
    got = ops.layer_norm_bwd(x.cuda(), gam.cuda(), 0.03, dy.cuda(), add=x.cuda())
    ref = H("layer_norm_bwd")(x, gam, 0.03, dy, add=x)
    assert all(close(a, b, 1e-4) for a, b in zip(got, ref))
    B, L, Hh, dh = 5, 20, 4, 16
    d = Hh * dh
    qkv = torch.randn(B * L, 3 * d, generator=g)
    R, rw, rr = torch.randn(2 * L, d, generator=g), torch.randn(d, generator=g), torch.randn(d, generator=g)
    dout = torch.randn(B * L, d, generator=g)
    got = ops.xlnet_attn_bwd(qkv.cuda(), R.cuda(), rw.cuda(), rr.cuda(), dout.cuda(), B, L, Hh)
    ref = H("xlnet_attn_bwd")(qkv, R, rw, rr, dout, B, L, Hh)
    assert all(close(a, b, 1e-4) for a, b in zip(got, ref))
    assert close(ops.causal_attn_bwd(qkv.cuda(), dout.cuda(), B, L, Hh), H("causal_attn_bwd")(qkv, dout, B, L, Hh), 1e-4)
    # the forward attention wrappers (inference kernels on fp32 q|k|v) against the doubles
    assert close(ops.xlnet_attn_fwd(qkv.cuda(), R.cuda(), rw.cuda(), rr.cuda(), B, L, Hh), DD.xlnet_attn_fwd(qkv, R, rw, rr, B, L, Hh), 1e-4)
    assert close(ops.causal_attn_fwd(qkv.cuda(), B, L, Hh), DD.causal_attn_fwd(qkv, B, L, Hh), 1e-4)


def test_dropout_kernels_device_vs_host_twin():
    """Dropout of the training step on the device: the element-wise mask is bit-identical to its host twin (same Philox
    stream), the attention forward / backward with dropped probabilities agree with theirs."""
    from transformers4rec_b200 import ops
    g = torch.Generator().manual_seed(43)
    H = ops.host_twin
    close = lambda a, b, tol=1e-5: (a.cpu() - b).abs().max().item() <= tol * max(1.0, b.abs().max().item())
    x = torch.randn(100_003, generator=g)
    assert torch.equal(ops.dropout(x.cuda(), 0.3, 2**40 + 17, 35).cpu(), H("dropout")(x, 0.3, 2**40 + 17, 35))
    B, L, Hh, dh = 5, 20, 4, 16
    d = Hh * dh
    drop = (0.3, 977, 18)
    qkv, dout = torch.randn(B * L, 3 * d, generator=g), torch.randn(B * L, d, generator=g)
    R, rw, rr = torch.randn(2 * L, d, generator=g), torch.randn(d, generator=g), torch.randn(d, generator=g)
    c = lambda *ts: [t.cuda() for t in ts]
    assert close(ops.attn_drop_fwd(*c(qkv, R, rw, rr), B, L, Hh, drop), H("attn_drop_fwd")(qkv, R, rw, rr, B, L, Hh, drop), 1e-4)
    assert close(ops.attn_drop_fwd(qkv.cuda(), None, None, None, B, L, Hh, drop), H("attn_drop_fwd")(qkv, None, None, None, B, L, Hh, drop), 1e-4)
    got = ops.xlnet_attn_bwd(*c(qkv, R, rw, rr, dout), B, L, Hh, drop=drop)
    ref = H("xlnet_attn_bwd")(qkv, R, rw, rr, dout, B, L, Hh, drop=drop)
    assert all(close(a, b, 1e-4) for a, b in zip(got, ref))
    assert close(ops.causal_attn_bwd(qkv.cuda(), dout.cuda(), B, L, Hh, drop=drop), H("causal_attn_bwd")(qkv, dout, B, L, Hh, drop=drop), 1e-4)
    pm = torch.rand(B, L, L, generator=g) < 0.4
    qkv2, dout2 = torch.cat([qkv, qkv.flip(0)]), torch.cat([dout, dout.flip(0)])
    assert close(ops.attn_drop_fwd(*c(qkv2, R, rw, rr), B, L, Hh, drop, plm_mask=pm.cuda()),
                 H("attn_drop_fwd")(qkv2, R, rw, rr, B, L, Hh, drop, plm_mask=pm), 1e-4)
    got = ops.xlnet_attn_bwd(*c(qkv2, R, rw, rr, dout2), B, L, Hh, plm_mask=pm.cuda(), drop=drop)
    ref = H("xlnet_attn_bwd")(qkv2, R, rw, rr, dout2, B, L, Hh, plm_mask=pm, drop=drop)
    assert all(close(a, b, 1e-4) for a, b in zip(got, ref))


@pytest.mark.parametrize("arch,masking", [("xlnet", "mlm"), ("gpt2", "clm")])
def test_training_step_with_dropout_on_gpu(monkeypatch, arch, masking):
    """Train mode on the device: dropout at HF's sites with masks regenerated in the backward -- loss and gradients
    against autograd of the oracle graph whose encoder is the restated forward carrying the SAME masks (host twin of
    the mask kernel)."""
    import t4r_oracle as O
    from transformers4rec_b200 import ops
    from transformers4rec_b200.training import FusedTrainingStep
    cards, dims = {"item_id/list": 3001, "category/list": 37}, {"item_id/list": 64, "category/list": 64}
    NL, Hh, p, seed = 2, 4, 0.3, 31337
    oracle, model = make_pair(cards, dims, "item_id/list", (), 64, Hh, NL, 20, arch=arch, masking=masking, weight_scale=0.08)
    oracle.train(False)
    B, L = 16, 20
    batch = synth_batch(B, L, cards, seed=5)
    u, draws = mlm_draws(B, L)
    model.heads[0].body[0].masking.set_draws(u.cuda())
    twin = ops.host_twin("dropout")

    def drop(site, t):
        return t * twin(torch.ones(t.numel()), p, seed, site).reshape(t.shape)

    def encoder_with_masks(hf, x):
        sd = dict(hf.named_parameters())
        if arch == "xlnet":
            return O.xlnet_forward_restated(x, sd, NL, Hh, drop=drop)
        return O.gpt2_forward_restated(x, sd, NL, Hh, eps=hf.config.layer_norm_epsilon, drop=drop)
    monkeypatch.setattr(O, "hf_encoder_forward", encoder_with_masks)
    ref_loss = _oracle_grads(oracle, batch, draws)
    step = FusedTrainingStep(model, head_chunk=1024).set_dropout_seed(seed)
    model.heads[0].body[1].transformer.train()
    for prm in model.parameters():
        prm.grad = None
    loss = step.forward({k: v.cuda() for k, v in batch.items()})
    step.backward()
    assert abs(loss.item() - ref_loss) < 1e-3
    for name, po, pm in _pairs(oracle, model):
        if po.grad is None and pm.grad is None:
            continue
        err = (pm.grad.cpu() - po.grad.reshape(pm.grad.shape)).abs().max().item()
        assert err < 1e-3 * max(1.0, po.grad.abs().max().item()), (name, err)


@pytest.mark.parametrize("arch,masking", [("xlnet", "mlm"), ("gpt2", "clm")])
def test_training_step_gradients_on_gpu(arch, masking):
    from transformers4rec_b200.training import FusedTrainingStep, training_loss
    cards = {"item_id/list": 3001, "category/list": 37}
    dims = {"item_id/list": 64, "category/list": 64}
    oracle, model = make_pair(cards, dims, "item_id/list", (), 64, 4, 2, 20, arch=arch, masking=masking, weight_scale=0.08)
    oracle.train(False)
    B, L = 32, 20
    batch = synth_batch(B, L, cards, seed=3)
    u, draws = mlm_draws(B, L)
    model.heads[0].body[0].masking.set_draws(u.cuda())
    ref_loss = _oracle_grads(oracle, batch, draws)
    step = FusedTrainingStep(model, head_chunk=1024)
    dev = {k: v.cuda() for k, v in batch.items()}
    loss = training_loss(model, dev, step)
    loss.backward()
    assert abs(loss.item() - ref_loss) < 1e-3
    for name, po, pm in _pairs(oracle, model):
        if po.grad is None and pm.grad is None:
            continue
        err = (pm.grad.cpu() - po.grad.reshape(pm.grad.shape)).abs().max().item()
        assert err < 1e-3 * max(1.0, po.grad.abs().max().item()), (name, err)
    # a few optimizer steps (torch.optim is plumbing) must lower the loss on the same batch
    opt = torch.optim.SGD(model.parameters(), lr=0.5)
    first = loss.item()
    for _ in range(5):
        opt.zero_grad()
        l = training_loss(model, dev, step)
        l.backward()
        opt.step()
    assert l.item() < first


def test_training_step_sampled_softmax_on_gpu():
    import t4r_oracle as O
    from transformers4rec_b200.training import FusedTrainingStep
    cards, dims, S = {"item_id/list": 3001}, {"item_id/list": 64}, 300
    oracle, model = make_pair(cards, dims, "item_id/list", (), 64, 4, 1, 20, weight_scale=0.08, sampled=True, max_n_samples=S)
    oracle.train(False)
    B, L = 32, 20
    batch = synth_batch(B, L, cards, seed=3)
    u, draws = mlm_draws(B, L)
    model.heads[0].body[0].masking.set_draws(u.cuda())
    task = model.heads[0].prediction_task_dict["next-item"]
    torch.manual_seed(4)
    raw = torch.multinomial(oracle.dist, 2 * S, replacement=True)
    task.set_negative_draws(raw.cuda())
    for p in oracle.parameters():
        p.grad = None
    ref = oracle(batch, training=True, draws=draws, neg_samples=O.negatives_from_draws(raw, S))
    ref["loss"].backward()
    step = FusedTrainingStep(model)
    loss = step.forward({k: v.cuda() for k, v in batch.items()})
    step.backward()
    assert abs(loss.item() - ref["loss"].item()) < 1e-3
    for name, po, pm in _pairs(oracle, model):
        if po.grad is None and pm.grad is None:
            continue
        err = (pm.grad.cpu() - po.grad.reshape(pm.grad.shape)).abs().max().item()
        assert err < 1e-3 * max(1.0, po.grad.abs().max().item()), (name, err)


def test_training_step_plm_on_gpu():
    """PLM: two-stream forward (MASKED tensor-path attention) and backward on the GPU vs autograd of HF's two-stream model."""
    from transformers4rec_b200.training import FusedTrainingStep
    oracle, model = make_pair({"item_id/list": 3001}, {"item_id/list": 64}, "item_id/list", (), 64, 4, 2, 20,
                              masking="plm", weight_scale=0.08)
    oracle.train(False)
    enc = model.heads[0].body[1].transformer
    with torch.no_grad():
        enc.mask_emb.normal_(0.0, 0.5)
        oracle.transformer.mask_emb.copy_(enc.mask_emb.cpu())
    B, L = 24, 20
    batch = synth_batch(B, L, {"item_id/list": 3001}, seed=8)
    g = torch.Generator().manual_seed(3)
    draws = {"u_span": torch.rand((B, L), generator=g), "u_start": torch.rand((B, L), generator=g),
             "u_force": torch.rand((B,), generator=g), "u_unmask": torch.rand((B,), generator=g),
             "perm": torch.stack([torch.randperm(L, generator=g) for _ in range(B)])}
    model.heads[0].body[0].masking.set_draws({k: v.cuda() for k, v in draws.items()})
    for p in oracle.parameters():
        p.grad = None
    ref = oracle(batch, training=True, draws=draws)
    ref["loss"].backward()
    step = FusedTrainingStep(model, head_chunk=1024)
    loss = step.forward({k: v.cuda() for k, v in batch.items()})
    step.backward()
    assert abs(loss.item() - ref["loss"].item()) < 1e-3
    for name, po, pm in list(_pairs(oracle, model)) + [("mask_emb", oracle.transformer.mask_emb, enc.mask_emb)]:
        if po.grad is None and pm.grad is None:
            continue
        err = (pm.grad.cpu() - po.grad.reshape(pm.grad.shape)).abs().max().item()
        assert err < 1e-3 * max(1.0, po.grad.abs().max().item()), (name, err)


def test_fused_adamw_on_gpu():
    """The AdamW kernel on the device against torch.optim.AdamW (fp32, same hyper-parameters) over several steps."""
    from transformers4rec_b200.training import FusedAdamW
    g = torch.Generator().manual_seed(5)
    shapes = [(1000, 64), (257,), (3, 5, 7)]
    pa = [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa = torch.optim.AdamW(pa, lr=1e-2, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.05)
    ob = FusedAdamW(pb, lr=1e-2, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.05)
    for _ in range(5):
        for x, y in zip(pa, pb):
            gr = torch.randn(x.shape, generator=g).cuda()
            x.grad, y.grad = gr.clone(), gr.clone()
        oa.step(); ob.step()
    for x, y in zip(pa, pb):
        assert (x - y).abs().max().item() < 1e-5


def test_model_fit_and_evaluate_on_gpu():
    """Model.fit (fused step + FusedAdamW by default) lowers the loss over a few epochs; Model.evaluate returns the
    streaming metrics and is deterministic."""
    cards = {"item_id/list": 3001, "category/list": 37}
    dims = {"item_id/list": 64, "category/list": 64}
    _, model = make_pair(cards, dims, "item_id/list", (), 64, 4, 2, 20, weight_scale=0.08)
    B, L = 32, 20
    u, _ = mlm_draws(B, L)
    model.heads[0].body[0].masking.set_draws(u.cuda())
    batches = [({k: v.cuda() for k, v in synth_batch(B, L, cards, seed=s).items()}, None) for s in (3, 4)]
    losses = model.fit(batches, num_epochs=6, verbose=False)
    assert losses.shape == (6,) and losses[-1] < losses[0]
    m1 = model.evaluate(batches, verbose=False)
    m2 = model.evaluate(batches, verbose=False)
    assert m1 and m1.keys() == m2.keys()
    assert all(torch.equal(torch.as_tensor(m1[k]), torch.as_tensor(m2[k])) for k in m1)


def test_widened_input_block_training_on_gpu():
    """Soft-embedding / element-wise kernels on the device vs their host twins, then the training step over the widened
    input block (soft embeddings + per-feature LayerNorm, every aggregation) against autograd of the oracle graph."""
    from test_host_training_cpu import _pairs_no_proj, _wide_batch, _wide_pair
    from transformers4rec_b200 import ops
    from transformers4rec_b200.training import FusedTrainingStep
    g = torch.Generator().manual_seed(2)
    M, n, dim = 500, 10, 64
    x, w, b = torch.randn(M, generator=g), torch.randn(n, generator=g), torch.randn(n, generator=g)
    tab, dout = torch.randn(n, dim, generator=g), torch.randn(M, dim, generator=g)
    H = ops.host_twin
    out, p = ops.soft_emb_fwd(x.cuda(), w.cuda(), b.cuda(), tab.cuda())
    out_h, p_h = H("soft_emb_fwd")(x, w, b, tab)
    assert (out.cpu() - out_h).abs().max().item() < 1e-5 and (p.cpu() - p_h).abs().max().item() < 1e-6
    dl, dlx = ops.soft_emb_bwd(x.cuda(), tab.cuda(), p, dout.cuda())
    dl_h, dlx_h = H("soft_emb_bwd")(x, tab, p_h, dout)
    assert (dl.cpu() - dl_h).abs().max().item() < 1e-4 and (dlx.cpu() - dlx_h).abs().max().item() < 1e-4
    assert torch.equal(ops.ew_mul(out, dout.cuda()).cpu(), out.cpu() * dout)
    for aggregation in ("concat", "element-wise-sum", "element-wise-sum-item-multi"):
        oracle, model = _wide_pair(aggregation, True, None, 32)
        model = model.cuda()
        B, L = 6, 8
        batch = _wide_batch(B, L)
        u, draws = mlm_draws(B, L)
        model.heads[0].body[0].masking.set_draws(u.cuda())
        ref_loss = _oracle_grads(oracle, batch, draws)
        step = FusedTrainingStep(model, head_chunk=128)
        for prm in model.parameters():
            prm.grad = None
        loss = step.forward({k: v.cuda() for k, v in batch.items()})
        step.backward()
        assert abs(loss.item() - ref_loss) < 1e-3
        for name, po, pm in list(oracle.pairs) + list(_pairs_no_proj(oracle, model)):
            if po.grad is None and pm.grad is None:
                continue
            err = (pm.grad.cpu() - po.grad.reshape(pm.grad.shape)).abs().max().item()
            assert err < 1e-3 * max(1.0, po.grad.abs().max().item()), (aggregation, name, err)


def _ddp_nccl_worker(rank, world, port, q):
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cards, dims = {"item_id/list": 3001, "category/list": 37}, {"item_id/list": 64, "category/list": 64}
        B, L = 16, 20
        oracle, model = make_pair(cards, dims, "item_id/list", (), 64, 4, 2, L, weight_scale=0.08, device=f"cuda:{rank}")
        oracle.train(False)
        batches = [synth_batch(B, L, cards, seed=100 + r) for r in range(world)]
        us = [mlm_draws(B, L, seed=200 + r) for r in range(world)]
        want = {}
        for r in range(world):
            for p in oracle.parameters():
                p.grad = None
            oracle(batches[r], training=True, draws=us[r][1])["loss"].backward()
            for n, p in oracle.named_parameters():
                if p.grad is not None:
                    want[n] = want.get(n, 0) + p.grad.detach().clone() / world
        names = {id(p): n for n, p in oracle.named_parameters()}
        model.heads[0].body[0].masking.set_draws(us[rank][0].cuda())
        model.enable_fused_training(head_chunk=1024)
        ddp = DDP(model, device_ids=[rank], find_unused_parameters=True)
        dev = {k: v.cuda() for k, v in batches[rank].items()}
        for _ in range(2):
            for p in model.parameters():
                p.grad = None
            ddp(dev, training=True)["loss"].backward()
        worst, n = 0.0, 0
        for name, po, pm in _pairs(oracle, model):
            key = names[id(po)]
            if key not in want and pm.grad is None:
                continue
            err = (pm.grad.cpu() - want[key].reshape(pm.grad.shape)).abs().max().item() / max(1.0, want[key].abs().max().item())
            worst, n = max(worst, err), n + 1
        q.put((rank, worst, n))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_training_under_ddp_nccl():
    """DistributedDataParallel (NCCL, 2 GPUs) around a model with the fused training step: every parameter's gradient is
    the mean over the ranks of the oracle's per-rank gradients."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_ddp_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, worst, n in res:
        assert worst < 1e-3 and n >= 15, (rank, worst, n)
