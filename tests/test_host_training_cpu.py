"""N3 on the CPU: the COMPOSITION of the fused training step (transformers4rec_b200/training.py) -- every backward
formula, transposition, residual branch, mask / code / tied-weight bookkeeping -- against torch autograd of the oracle
graph, with the kernels replaced by the test doubles (tests/_ops_double.py: each primitive is a few lines of torch, the
backward ones obtained from autograd of the forward one).  What runs on the GPU instead of the doubles is covered by the
host twins (tests/test_abi_and_host.py) and the gated GPU tests."""
import numpy as np
import pytest
import torch

import _ops_double as D
import t4r_oracle as O
from _util import make_pair, mlm_draws, synth_batch

CARDS = {"item_id/list": 1501, "category/list": 37}
CONT = ("cont0/list",)


def _oracle_grads(oracle, batch, draws):
    for p in oracle.parameters():
        p.grad = None
    out = oracle(batch, training=True, draws=draws)
    out["loss"].backward()
    return out["loss"].item()


def _pairs(oracle, model):
    """(name, oracle parameter, product parameter) for every trainable tensor of the path."""
    head = model.heads[0]
    inputs, tblock = head.body[0], head.body[1]
    yield "masked_item_embedding", oracle.masked_item_embedding, inputs.masking.masked_item_embedding
    for name in oracle.table_names:
        yield f"table[{name}]", oracle.tables[name.replace("/", "__")].weight, inputs.categorical_module.embedding_tables[name].weight
    lin = inputs.projection_module[0][0]
    yield "proj.weight", oracle.proj.weight, lin.weight
    yield "proj.bias", oracle.proj.bias, lin.bias
    od, md = dict(oracle.transformer.named_parameters()), dict(tblock.transformer.named_parameters())
    for k, v in md.items():
        if k in od and k not in ("word_embedding.weight", "mask_emb", "wte.weight"):
            yield "transformer." + k, od[k], v
    if oracle.task_block is not None:
        tl = head.prediction_task_dict["next-item"].task_block[0][0]
        yield "task_block.weight", oracle.task_block.weight, tl.weight
        yield "task_block.bias", oracle.task_block.bias, tl.bias


@pytest.mark.parametrize("arch,masking,dims", [("xlnet", "mlm", {"item_id/list": 32, "category/list": 32}),
                                               ("xlnet", "clm", {"item_id/list": 32, "category/list": 16}),
                                               ("gpt2", "clm", {"item_id/list": 16, "category/list": 32})])
def test_training_step_gradients_match_autograd(monkeypatch, arch, masking, dims):
    from transformers4rec_b200.training import FusedTrainingStep, training_loss
    D.install(monkeypatch)
    oracle, model = make_pair(CARDS, dims, "item_id/list", CONT, 32, 2, 2, 10, arch=arch, masking=masking, device="cpu",
                              weight_scale=0.08)
    oracle.train(False)   # dropout off, gradients on
    B, L = 7, 10
    batch = synth_batch(B, L, CARDS, CONT, seed=3)
    u, draws = mlm_draws(B, L)
    model.heads[0].body[0].masking.set_draws(u)
    ref_loss = _oracle_grads(oracle, batch, draws)
    step = FusedTrainingStep(model, head_chunk=400)     # several column chunks of the 1501-row table
    for p in model.parameters():
        p.grad = None
    loss = step.forward(batch)
    step.backward()
    assert abs(loss.item() - ref_loss) < 1e-4
    checked = 0
    for name, po, pm in _pairs(oracle, model):
        if po.grad is None and pm.grad is None:
            continue            # parameters the path never touches (XLNet segment embeddings)
        assert pm.grad is not None, name
        go = po.grad if po.grad is not None else torch.zeros_like(po)
        err = (pm.grad - go.reshape(pm.grad.shape)).abs().max().item()
        assert err < 2e-4 * max(1.0, go.abs().max().item()), (name, err, go.abs().max().item())
        checked += 1
    assert checked >= 20
    # the autograd bridge: loss.backward() delivers the same gradients (scaled by the upstream gradient)
    saved = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    for p in model.parameters():
        p.grad = None
    out = training_loss(model, batch, step) * 2.0
    out.backward()
    for n, p in model.named_parameters():
        if n in saved:
            assert torch.allclose(p.grad, 2.0 * saved[n], atol=1e-6), n


def test_training_step_with_the_real_per_item_code(monkeypatch):
    """Same gradient check, but every element / row primitive of t4r_train.cu runs its REAL per-item code through the
    host twins (only the GEMMs, the gathers and the forward attention / head kernels stay doubles)."""
    from transformers4rec_b200 import ops
    from transformers4rec_b200.training import FusedTrainingStep
    names = ("transpose", "act_fwd", "act_bwd", "add_positions", "sum_over_sessions", "apply_row_codes", "row_codes_bwd",
             "gather_rows", "scatter_rows", "softmax_ce_bwd", "index_add_rows", "col_sum", "layer_norm_fwd",
             "layer_norm_bwd", "xlnet_attn_bwd", "causal_attn_bwd")
    twins = {name: ops.host_twin(name) for name in names}   # bound to the real entry points before the doubles go in
    D.install(monkeypatch)
    for name in names:
        monkeypatch.setattr(ops, name, twins[name])
    for arch, masking in (("xlnet", "mlm"), ("gpt2", "clm")):
        oracle, model = make_pair(CARDS, {"item_id/list": 16, "category/list": 32}, "item_id/list", CONT, 32, 2, 1, 8,
                                  arch=arch, masking=masking, device="cpu", weight_scale=0.08)
        oracle.train(False)
        B, L = 5, 8
        batch = synth_batch(B, L, CARDS, CONT, seed=4)
        u, draws = mlm_draws(B, L)
        model.heads[0].body[0].masking.set_draws(u)
        ref_loss = _oracle_grads(oracle, batch, draws)
        step = FusedTrainingStep(model, head_chunk=512)
        for p in model.parameters():
            p.grad = None
        loss = step.forward(batch)
        step.backward()
        assert abs(loss.item() - ref_loss) < 1e-4
        for name, po, pm in _pairs(oracle, model):
            if po.grad is None and pm.grad is None:
                continue
            err = (pm.grad - po.grad.reshape(pm.grad.shape)).abs().max().item()
            assert err < 3e-4 * max(1.0, po.grad.abs().max().item()), (arch, name, err)


def test_sgd_trajectories_coincide(monkeypatch):
    """Five optimizer steps through ``training_loss(...).backward()`` and through torch autograd of the oracle graph,
    same data and draws: losses fall and the two parameter trajectories stay together (tied item table included)."""
    from transformers4rec_b200.training import FusedTrainingStep, training_loss
    D.install(monkeypatch)
    oracle, model = make_pair(CARDS, {"item_id/list": 32, "category/list": 32}, "item_id/list", CONT, 32, 2, 1, 8,
                              device="cpu", weight_scale=0.08)
    oracle.train(False)
    B, L = 6, 8
    batch = synth_batch(B, L, CARDS, CONT, seed=9)
    u, draws = mlm_draws(B, L)
    model.heads[0].body[0].masking.set_draws(u)
    step = FusedTrainingStep(model, head_chunk=512)
    opt_o = torch.optim.SGD(oracle.parameters(), lr=0.3)
    opt_m = torch.optim.SGD(model.parameters(), lr=0.3)
    losses = []
    for _ in range(5):
        opt_o.zero_grad()
        lo = oracle(batch, training=True, draws=draws)["loss"]
        lo.backward()
        opt_o.step()
        opt_m.zero_grad()
        lm = training_loss(model, batch, step)
        lm.backward()
        opt_m.step()
        losses.append((lo.item(), lm.item()))
    assert all(abs(a - b) < 1e-3 for a, b in losses) and losses[-1][1] < losses[0][1]
    for name, po, pm in _pairs(oracle, model):
        assert (pm.detach() - po.detach().reshape(pm.shape)).abs().max().item() < 1e-3, name


def test_training_step_sampled_softmax(monkeypatch):
    """Sampled softmax (config 5's head): gradients against autograd of the oracle's sampled head with the same negatives,
    with the real per-item code of the new kernel on its host twin."""
    from transformers4rec_b200 import ops
    from transformers4rec_b200.training import FusedTrainingStep
    twin = ops.host_twin("sampled_ce_bwd")
    D.install(monkeypatch)
    monkeypatch.setattr(ops, "sampled_ce_bwd", twin)
    S = 150
    oracle, model = make_pair(CARDS, {"item_id/list": 32, "category/list": 32}, "item_id/list", CONT, 32, 2, 1, 8,
                              device="cpu", weight_scale=0.08, sampled=True, max_n_samples=S)
    oracle.train(False)
    B, L = 6, 8
    batch = synth_batch(B, L, CARDS, CONT, seed=5)
    u, draws = mlm_draws(B, L)
    model.heads[0].body[0].masking.set_draws(u)
    task = model.heads[0].prediction_task_dict["next-item"]
    torch.manual_seed(4)
    raw = torch.multinomial(oracle.dist, 2 * S, replacement=True)
    labels_all = batch["item_id/list"][batch["item_id/list"] != 0]
    raw[:3] = labels_all[:3]     # make accidental hits likely
    task.set_negative_draws(raw)
    for p in oracle.parameters():
        p.grad = None
    ref = oracle(batch, training=True, draws=draws, neg_samples=O.negatives_from_draws(raw, S))
    ref["loss"].backward()
    step = FusedTrainingStep(model)
    for p in model.parameters():
        p.grad = None
    loss = step.forward(batch)
    step.backward()
    assert abs(loss.item() - ref["loss"].item()) < 1e-4
    for name, po, pm in _pairs(oracle, model):
        if po.grad is None and pm.grad is None:
            continue
        err = (pm.grad - po.grad.reshape(pm.grad.shape)).abs().max().item()
        assert err < 3e-4 * max(1.0, po.grad.abs().max().item()), (name, err)


def test_training_step_label_smoothing(monkeypatch):
    """nn.CrossEntropyLoss(label_smoothing=0.1) (transformers4rec/torch/losses.py:4-20): the smoothed target
    distribution in the head's backward, with the kernel's real per-item code."""
    from transformers4rec_b200 import ops
    from transformers4rec_b200.training import FusedTrainingStep
    twin = ops.host_twin("softmax_ce_bwd")
    D.install(monkeypatch)
    monkeypatch.setattr(ops, "softmax_ce_bwd", twin)
    oracle, model = make_pair(CARDS, {"item_id/list": 32, "category/list": 32}, "item_id/list", CONT, 32, 2, 1, 8,
                              device="cpu", weight_scale=0.08)
    oracle.train(False)
    B, L = 6, 8
    batch = synth_batch(B, L, CARDS, CONT, seed=6)
    u, draws = mlm_draws(B, L)
    model.heads[0].body[0].masking.set_draws(u)
    model.heads[0].prediction_task_dict["next-item"].label_smoothing = 0.1
    for p in oracle.parameters():
        p.grad = None
    out = oracle(batch, training=True, draws=draws)
    ref = torch.nn.functional.cross_entropy(out["predictions"], out["labels"], label_smoothing=0.1)
    ref.backward()
    step = FusedTrainingStep(model, head_chunk=500)
    for p in model.parameters():
        p.grad = None
    loss = step.forward(batch)
    step.backward()
    assert abs(loss.item() - ref.item()) < 1e-4
    for name, po, pm in _pairs(oracle, model):
        if po.grad is None and pm.grad is None:
            continue
        err = (pm.grad - po.grad.reshape(pm.grad.shape)).abs().max().item()
        assert err < 3e-4 * max(1.0, po.grad.abs().max().item()), (name, err)


def test_model_forward_routes_to_the_training_step_when_enabled(monkeypatch):
    """``model.enable_fused_training()``: the reference's training loop shape -- forward with training=True, loss.backward(),
    optimizer.step() -- runs on the fused step; under no_grad (or with the switch off) nothing changes."""
    D.install(monkeypatch)
    oracle, model = make_pair(CARDS, {"item_id/list": 32, "category/list": 32}, "item_id/list", CONT, 32, 2, 1, 8,
                              device="cpu", weight_scale=0.08)
    batch = synth_batch(6, 8, CARDS, CONT, seed=7)
    u, draws = mlm_draws(6, 8)
    model.heads[0].body[0].masking.set_draws(u)
    with torch.no_grad():
        plain = model(batch, training=True)["loss"].item()
    model.enable_fused_training(head_chunk=512)
    with torch.no_grad():
        assert model(batch, training=True)["loss"].item() == plain          # no autograd -> the forward-only path
    opt = torch.optim.SGD(model.parameters(), lr=0.2)
    first = None
    for _ in range(4):
        opt.zero_grad()
        out = model(batch, training=True)
        out["loss"].backward()
        opt.step()
        first = first if first is not None else out["loss"].item()
    assert abs(first - plain) < 1e-4 and out["loss"].item() < first
    assert out["labels"].numel() > 0 and all(p.grad is not None for n, p in model.named_parameters()
                                              if "seg_embed" not in n and "r_s_bias" not in n and "word_embedding" not in n
                                              and "mask_emb" not in n)
    model.enable_fused_training(False)
    assert getattr(model, "_fused_step") is None


def test_training_step_permutation_language_modeling(monkeypatch):
    """PLM: XLNet's two-stream forward and its backward (incl. mask_emb) against autograd of HF's own two-stream model,
    with the attention backward's real per-item code (mask / stream handling) on its host twin."""
    from transformers4rec_b200 import ops
    from transformers4rec_b200.training import FusedTrainingStep
    twin = ops.host_twin("xlnet_attn_bwd")
    D.install(monkeypatch)
    monkeypatch.setattr(ops, "xlnet_attn_bwd", twin)
    oracle, model = make_pair({"item_id/list": 801}, {"item_id/list": 32}, "item_id/list", (), 32, 2, 2, 10,
                              masking="plm", device="cpu", weight_scale=0.08)
    oracle.train(False)
    enc = model.heads[0].body[1].transformer
    with torch.no_grad():
        enc.mask_emb.normal_(0.0, 0.5)
        oracle.transformer.mask_emb.copy_(enc.mask_emb)
    B, L = 7, 10
    batch = synth_batch(B, L, {"item_id/list": 801}, seed=8)
    g = torch.Generator().manual_seed(3)
    draws = {"u_span": torch.rand((B, L), generator=g), "u_start": torch.rand((B, L), generator=g),
             "u_force": torch.rand((B,), generator=g), "u_unmask": torch.rand((B,), generator=g),
             "perm": torch.stack([torch.randperm(L, generator=g) for _ in range(B)])}
    model.heads[0].body[0].masking.set_draws(draws)
    for p in oracle.parameters():
        p.grad = None
    ref = oracle(batch, training=True, draws=draws)
    ref["loss"].backward()
    step = FusedTrainingStep(model, head_chunk=300)
    for p in model.parameters():
        p.grad = None
    loss = step.forward(batch)
    step.backward()
    assert abs(loss.item() - ref["loss"].item()) < 1e-4
    pairs = list(_pairs(oracle, model)) + [("mask_emb", oracle.transformer.mask_emb, enc.mask_emb)]
    for name, po, pm in pairs:
        if po.grad is None and pm.grad is None:
            continue
        err = (pm.grad - po.grad.reshape(pm.grad.shape)).abs().max().item()
        assert err < 3e-4 * max(1.0, po.grad.abs().max().item()), (name, err)
    assert enc.mask_emb.grad is not None and enc.mask_emb.grad.abs().max().item() > 0


def test_fused_adamw_follows_torch_adamw(monkeypatch):
    """FusedAdamW (kernel's real per-element code on its host twin) against torch.optim.AdamW over several steps, with
    weight decay, on parameters of different shapes; and through a short training run of the fused step."""
    from transformers4rec_b200 import ops
    from transformers4rec_b200.training import FusedAdamW, FusedTrainingStep, training_loss
    twin = ops.host_twin("adamw_step")
    D.install(monkeypatch)
    monkeypatch.setattr(ops, "adamw_step", twin)
    g = torch.Generator().manual_seed(1)
    shapes = [(7, 5), (13,), (3, 4, 2)]
    pa = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa = torch.optim.AdamW(pa, lr=3e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    ob = FusedAdamW(pb, lr=3e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    for it in range(6):
        for x, y in zip(pa, pb):
            gr = torch.randn(x.shape, generator=g)
            x.grad, y.grad = gr.clone(), gr.clone()
        oa.step(); ob.step()
    for x, y in zip(pa, pb):
        assert (x - y).abs().max().item() < 1e-6
    assert ob.state[pb[0]]["step"] == 6 and set(ob.state[pb[0]]) == {"step", "exp_avg", "exp_avg_sq"}
    # a short run of the whole step with it
    oracle, model = make_pair(CARDS, {"item_id/list": 32, "category/list": 32}, "item_id/list", CONT, 32, 2, 1, 8,
                              device="cpu", weight_scale=0.08)
    batch = synth_batch(6, 8, CARDS, CONT, seed=9)
    u, _ = mlm_draws(6, 8)
    model.heads[0].body[0].masking.set_draws(u)
    step = FusedTrainingStep(model, head_chunk=512)
    opt = FusedAdamW(model.parameters(), lr=1e-2)
    losses = []
    for _ in range(4):
        opt.zero_grad()
        l = training_loss(model, batch, step)
        l.backward()
        opt.step()
        losses.append(l.item())
    assert losses[-1] < losses[0]


def test_model_fit_and_evaluate(monkeypatch, capsys):
    """Model.fit / Model.evaluate (model/base.py:669-738; the reference's tests call them as
    ``losses = model.fit(dataset, num_epochs=5); metrics = model.evaluate(dataset)``, test_model.py:228-229): fit's loss
    curve equals a hand-written loop with torch.optim.Adam over the same fused step, evaluate returns the streaming
    metrics of the evaluation forward."""
    from transformers4rec_b200 import ops
    from transformers4rec_b200.training import FusedTrainingStep, training_loss
    twin = ops.host_twin("adamw_step")
    D.install(monkeypatch)
    monkeypatch.setattr(ops, "adamw_step", twin)
    dims = {"item_id/list": 32, "category/list": 32}

    def fresh():
        _, m = make_pair(CARDS, dims, "item_id/list", CONT, 32, 2, 1, 8, device="cpu", weight_scale=0.08)
        u, _ = mlm_draws(6, 8)
        m.heads[0].body[0].masking.set_draws(u)
        m.heads[0].body[1].transformer.config.dropout = 0.0   # fit() trains in train() mode: keep the curves comparable
        return m
    batches = [(synth_batch(6, 8, CARDS, CONT, seed=s), None) for s in (3, 4)]
    model = fresh()
    losses = model.fit(batches, num_epochs=5, verbose=False)
    assert losses.shape == (5,) and np.isfinite(losses).all() and losses[-1] < losses[0]
    # the same run by hand: Adam (the reference's default optimizer) on the fused step
    other = fresh()
    step = FusedTrainingStep(other, head_chunk=512)
    opt = torch.optim.Adam(other.parameters())
    want = []
    for _ in range(5):
        ep = []
        for x, _y in batches:
            loss = training_loss(other, x, step)
            opt.zero_grad(); loss.backward(); opt.step()
            ep.append(float(loss.detach()))
        want.append(np.mean(ep))
    assert np.allclose(losses, np.array(want), rtol=2e-5, atol=2e-6)
    # an optimizer CLASS is instantiated on the model's parameters, as in the reference
    third = fresh()
    l3 = third.fit(batches, optimizer=torch.optim.Adam, num_epochs=5, verbose=False)
    assert np.allclose(l3, losses, rtol=2e-5, atol=2e-6)
    # evaluate: streaming metrics of the evaluation forward; verbose fit prints them per epoch
    metrics = model.evaluate(batches, verbose=False)
    assert metrics and all(torch.isfinite(torch.as_tensor(v)).all() for v in metrics.values())
    assert any("recall_at" in k for k in metrics)
    again = model.evaluate(batches, verbose=False)
    assert metrics.keys() == again.keys() and all(torch.equal(torch.as_tensor(metrics[k]), torch.as_tensor(again[k]))
                                                  for k in metrics)
    model.fit(batches, eval_dataloader=batches, num_epochs=1, verbose=True)
    assert "recall_at" in capsys.readouterr().out
    # train-mode metrics come from the label ranks among the training logits: one epoch at lr 0 leaves the weights
    # alone, so they must equal the metrics of the forward-only training pass (materialised predictions)
    model.reset_metrics()
    model.fit(batches, optimizer=torch.optim.SGD(model.parameters(), lr=0.0), num_epochs=1, verbose=False)
    from_ranks = model.compute_metrics(mode="train")
    model.reset_metrics()
    model.fit(batches, num_epochs=1, train=False, verbose=False)
    from_logits = model.compute_metrics(mode="train")
    assert from_ranks.keys() == from_logits.keys() and len(from_ranks) >= 2
    assert any(float(torch.as_tensor(v).max()) > 0 for v in from_ranks.values())
    for k in from_ranks:
        assert torch.allclose(torch.as_tensor(from_ranks[k]), torch.as_tensor(from_logits[k]), atol=1e-6), k
    # train=False: forward-only pass with train-mode metrics, no parameter moves
    before = [p.detach().clone() for p in model.parameters()]
    l0 = model.fit(batches, num_epochs=1, train=False, verbose=False)
    assert np.isfinite(l0).all() and all(torch.equal(a, b) for a, b in zip(before, model.parameters()))


def test_fit_evaluate_fit_evaluate_sees_fresh_weights(monkeypatch):
    """FusedAdamW writes the parameters through a raw pointer; the forward-only path caches split planes of the GEMM
    weights keyed on ``(data_ptr, _version)``.  After fit -> evaluate (caches filled) -> fit, the second evaluate must
    run on the NEW weights: it has to equal an evaluate with every plane cache dropped."""
    from transformers4rec_b200 import ops
    twin = ops.host_twin("adamw_step")
    D.install(monkeypatch)
    monkeypatch.setattr(ops, "adamw_step", twin)
    _, model = make_pair(CARDS, {"item_id/list": 32, "category/list": 32}, "item_id/list", CONT, 32, 2, 1, 8,
                         device="cpu", weight_scale=0.08)
    u, _ = mlm_draws(6, 8)
    model.heads[0].body[0].masking.set_draws(u)
    batches = [(synth_batch(6, 8, CARDS, CONT, seed=s), None) for s in (3, 4)]

    def eval_loss():
        with torch.no_grad():
            return float(model(batches[0][0], training=False, testing=True)["loss"])
    model.fit(batches, num_epochs=1, verbose=False)
    first = eval_loss()                                   # fills the plane caches
    versions = [p._version for p in model.parameters()]
    model.fit(batches, num_epochs=10, verbose=False)
    assert all(p._version > v for p, v in zip(model.parameters(), versions) if p.grad is not None)
    cached = eval_loss()
    for m in model.modules():
        pc = getattr(m, "_planes", None)
        if isinstance(pc, ops.PlaneCache):
            pc.clear()
    fresh = eval_loss()
    assert abs(first - fresh) > 1e-3, "the second fit did not move the evaluation loss: the test would prove nothing"
    assert cached == fresh, (cached, fresh)


# ---------------------------------------------------------------------------------------------------------------
# the widened input block in training (N3 x N4): soft embeddings, per-feature LayerNorm, continuous projection,
# element-wise aggregations
# ---------------------------------------------------------------------------------------------------------------
class _WideOracle(O.OracleSessionModel):
    """The oracle graph with the input block restated, with the oracle's functions, from clones of the product input
    block's own parameters (embedding rows -> LayerNorm, soft embeddings -> LayerNorm, continuous projection MLP,
    aggregation, projection, MLM mask)."""

    def attach(self, inputs):
        import torch.nn.functional as F  # noqa: F401
        self.aggregation = inputs.aggregation or "concat"
        self.extra = torch.nn.ParameterList()
        self.pairs = []

        def clone(p, label):
            q = torch.nn.Parameter(p.detach().clone())
            self.extra.append(q)
            self.pairs.append((label, q, p))
            return q
        cm, cont = inputs.categorical_module, inputs.continuous_module
        self.cat_ln = {}
        if getattr(cm, "post", None) is not None:
            for n, m in cm.post.feature_layer_norm.items():
                self.cat_ln[n] = (clone(m.weight, f"cat_ln[{n}].weight"), clone(m.bias, f"cat_ln[{n}].bias"))
        self.soft, self.mlp, self.plain_cont = {}, None, ()
        if hasattr(cont, "embedding_tables"):
            for n, m in cont.embedding_tables.items():
                ln = cont.post.feature_layer_norm[n] if cont.post is not None and n in cont.post.feature_layer_norm else None
                self.soft[n] = (clone(m.projection_layer.weight, f"soft[{n}].w"), clone(m.projection_layer.bias, f"soft[{n}].b"),
                                clone(m.embedding_table.weight, f"soft[{n}].table"),
                                (clone(ln.weight, f"soft_ln[{n}].weight"), clone(ln.bias, f"soft_ln[{n}].bias")) if ln else None)
        elif hasattr(cont, "mlp"):
            self.mlp = ([(clone(b[0].weight, f"mlp{i}.w"), clone(b[0].bias, f"mlp{i}.b")) for i, b in enumerate(cont.mlp)],
                        list(cont.features))
        elif cont is not None:
            self.plain_cont = tuple(cont.features)
        lin = inputs._projection_linear()
        if lin is None:
            self.proj = None
        else:
            self.proj = torch.nn.Linear(lin.in_features, lin.out_features)
            with torch.no_grad():
                self.proj.weight.copy_(lin.weight); self.proj.bias.copy_(lin.bias)
            self.pairs += [("proj.weight", self.proj.weight, lin.weight), ("proj.bias", self.proj.bias, lin.bias)]
        return self

    def input_block(self, inputs, training, testing, draws=None):
        import torch.nn.functional as F
        feats = {}
        for n in self.table_names:
            x = F.embedding(inputs[n], self.tables[n.replace("/", "__")].weight, padding_idx=0)
            feats[n] = O.tabular_layer_norm(x, *self.cat_ln[n]) if n in self.cat_ln else x
        for n, (w, b, table, ln) in self.soft.items():
            y = O.soft_embedding(inputs[n], w, b, table)
            feats[n] = O.tabular_layer_norm(y, *ln) if ln is not None else y
        if self.mlp is not None:
            layers, names = self.mlp
            c = torch.cat([inputs[n].unsqueeze(-1) for n in names], dim=-1)
            for w, b in layers:
                c = O.project_relu(c, w, b)
            feats["continuous_projection"] = c
        for n in self.plain_cont:
            feats[n] = inputs[n].float().unsqueeze(-1)
        x = O.aggregate(feats, self.aggregation, self.item_id)
        if self.proj is not None:
            x = O.project_relu(x, self.proj.weight, self.proj.bias)
        d = draws or {}
        mask, labels = O.mlm_compute_masked_targets(inputs[self.item_id], training, testing, u_bern=d.get("u_bern"),
                                                    u_force=d.get("u_force"), u_unmask=d.get("u_unmask"))
        return O.mlm_apply_mask_to_inputs(x, mask, self.masked_item_embedding, training, testing), mask, labels


def _wide_pair(aggregation, soft, proj_cont, d_output, D=16, d_model=32, L=8):
    import transformers4rec_b200.torch as tr
    torch.manual_seed(31)
    schema = tr.Schema([tr.ColumnSchema.create_categorical("item_id/list", 300, tags=[tr.Tags.ITEM_ID]),
                        tr.ColumnSchema.create_categorical("category/list", 40),
                        tr.ColumnSchema.create_categorical("user_country", 30, is_list=False),
                        tr.ColumnSchema.create_continuous("price/list"),
                        tr.ColumnSchema.create_continuous("rel_time/list")])
    kw = dict(continuous_soft_embeddings=True, soft_embedding_dim_default=D) if soft else {}
    if proj_cont:
        kw["continuous_projection"] = proj_cont
    inputs = tr.TabularSequenceFeatures.from_schema(schema, max_sequence_length=L, aggregation=aggregation,
                                                    embedding_dim_default=D, post="layer-norm", d_output=d_output,
                                                    masking="mlm", **kw)
    model = tr.XLNetConfig.build(d_model, 2, 1, L).to_torch_model(inputs, tr.NextItemPredictionTask(weight_tying=True))
    model = model.eval()
    inputs, tblock = model.heads[0].body[0], model.heads[0].body[1]
    with torch.no_grad():
        for n, p in tblock.transformer.named_parameters():
            if "layer_norm" not in n and (p.ndim >= 2 or "bias" in n):
                p.normal_(0.0, 0.08)
        inputs.masking.masked_item_embedding.normal_(0.0, 0.5)
        posts = [m for m in (getattr(inputs.categorical_module, "post", None), getattr(inputs.continuous_module, "post", None)) if m]
        for post in posts:           # non-trivial LayerNorm parameters
            for mod in post.feature_layer_norm.values():
                mod.weight.uniform_(0.5, 1.5); mod.bias.normal_(0.0, 0.3)
    cards = {"item_id/list": 301, "category/list": 41, "user_country": 31}
    oracle = _WideOracle(cardinalities=cards, embedding_dims={k: D for k in cards}, item_id="item_id/list", continuous=(),
                         d_model=d_model, n_head=2, n_layer=1, max_seq_len=L, arch="xlnet", masking="mlm",
                         project=(d_model != D) or True).eval()
    with torch.no_grad():
        for name in oracle.table_names:
            oracle.tables[name.replace("/", "__")].weight.copy_(inputs.categorical_module.embedding_tables[name].weight)
        oracle.masked_item_embedding.copy_(inputs.masking.masked_item_embedding)
        oracle.transformer.load_state_dict(tblock.transformer.state_dict(), strict=False)
        task = model.heads[0].prediction_task_dict["next-item"]
        if oracle.task_block is not None:
            tl = task.task_block[0][0]
            oracle.task_block.weight.copy_(tl.weight); oracle.task_block.bias.copy_(tl.bias)
    oracle.attach(inputs)
    return oracle, model


def _wide_batch(B, L, seed=4):
    batch = synth_batch(B, L, {"item_id/list": 301, "category/list": 41}, ("price/list", "rel_time/list"), seed=seed)
    batch["user_country"] = torch.randint(1, 31, (B,), generator=torch.Generator().manual_seed(seed))
    return batch


@pytest.mark.parametrize("aggregation,soft,proj_cont,d_output", [
    ("concat", True, None, 32),                           # the paper recipes: soft embeddings + layer-norm + MLP merge
    ("element-wise-sum", True, None, 32),
    ("element-wise-sum-item-multi", True, None, 32),
    ("element-wise-sum", True, None, None),               # no projection: the aggregate enters the encoder directly
    ("concat", False, [24, 16], 32),                      # continuous projection MLP as one feature
    ("concat", False, None, 32),                          # plain scalars next to LayerNorm'd embeddings
])
def test_training_step_widened_input_block(monkeypatch, aggregation, soft, proj_cont, d_output):
    """Gradients of every parameter of the widened input block (and everything above it) against torch autograd of the
    oracle composition; the new element / row kernels run their real per-item code through the host twins."""
    from transformers4rec_b200 import ops
    from transformers4rec_b200.training import FusedTrainingStep
    names = ("soft_emb_fwd", "soft_emb_bwd", "ew_add", "ew_mul", "gather_rows", "layer_norm_fwd", "layer_norm_bwd",
             "index_add_rows", "col_sum", "act_fwd", "act_bwd", "transpose")
    twins = {name: ops.host_twin(name) for name in names}
    D.install(monkeypatch)
    for name in names:
        monkeypatch.setattr(ops, name, twins[name])
    d_model = 32 if d_output else 16
    oracle, model = _wide_pair(aggregation, soft, proj_cont, d_output, D=16, d_model=d_model)
    B, L = 6, 8
    batch = _wide_batch(B, L)
    u, draws = mlm_draws(B, L)
    inputs = model.heads[0].body[0]
    inputs.masking.set_draws(u)
    # forward agreement of the two product paths first: training composition vs the one-kernel inference input block
    ref_loss = _oracle_grads(oracle, batch, draws)
    step = FusedTrainingStep(model, head_chunk=128)
    assert step.wide is not None
    for p in model.parameters():
        p.grad = None
    loss = step.forward(batch)
    step.backward()
    assert abs(loss.item() - ref_loss) < 1e-4
    with torch.no_grad():
        assert abs(model(batch, training=True)["loss"].item() - loss.item()) < 1e-4
    pairs = list(oracle.pairs) + [(n, a, b) for n, a, b in _pairs_no_proj(oracle, model)]
    seen = 0
    for name, po, pm in pairs:
        if po.grad is None and pm.grad is None:
            continue
        assert po.grad is not None and pm.grad is not None, name
        err = (pm.grad - po.grad.reshape(pm.grad.shape)).abs().max().item()
        assert err < 3e-4 * max(1.0, po.grad.abs().max().item()), (name, err)
        seen += 1
    assert seen >= 10 + 2 * (3 if soft else 0)


def _pairs_no_proj(oracle, model):
    head = model.heads[0]
    inputs, tblock = head.body[0], head.body[1]
    yield "masked_item_embedding", oracle.masked_item_embedding, inputs.masking.masked_item_embedding
    for name in oracle.table_names:
        yield f"table[{name}]", oracle.tables[name.replace("/", "__")].weight, inputs.categorical_module.embedding_tables[name].weight
    od, md = dict(oracle.transformer.named_parameters()), dict(tblock.transformer.named_parameters())
    for k, v in md.items():
        if k in od and k not in ("word_embedding.weight", "mask_emb", "wte.weight"):
            yield "transformer." + k, od[k], v
    if oracle.task_block is not None:
        tl = head.prediction_task_dict["next-item"].task_block[0][0]
        yield "task_block.weight", oracle.task_block.weight, tl.weight
        yield "task_block.bias", oracle.task_block.bias, tl.bias


def test_soft_embedding_and_binary_twins_vs_torch():
    """Real per-item code of the soft-embedding forward / backward and the element-wise kernels against torch autograd
    of the oracle's soft_embedding (features/embedding.py:517-556)."""
    from transformers4rec_b200 import ops
    g = torch.Generator().manual_seed(0)
    M, n, dim = 37, 10, 16
    x, w, b = torch.randn(M, generator=g), torch.randn(n, 1, generator=g), torch.randn(n, generator=g)
    tab = torch.randn(n, dim, generator=g)
    leaves = [t.clone().requires_grad_() for t in (w, b, tab)]
    ref = O.soft_embedding(x, *leaves)
    dout = torch.randn(M, dim, generator=g)
    ref.backward(dout)
    H = ops.host_twin
    out, p = H("soft_emb_fwd")(x, w, b, tab)
    assert (out - ref.detach()).abs().max().item() < 1e-5 and (p.sum(1) - 1).abs().max().item() < 1e-5
    dl, dlx = H("soft_emb_bwd")(x, tab, p, dout)
    assert (dlx.sum(0) - leaves[0].grad.reshape(-1)).abs().max().item() < 1e-4
    assert (dl.sum(0) - leaves[1].grad).abs().max().item() < 1e-4
    assert (p.t() @ dout - leaves[2].grad).abs().max().item() < 1e-4
    assert torch.equal(H("ew_add")(out, dout), out + dout) and torch.equal(H("ew_mul")(out, dout), out * dout)


def test_training_step_with_stochastic_swap_noise(monkeypatch):
    """StochasticSwapNoise as the input block's ``pre`` (tabular/transformations.py:29-92; the recipes'
    --stochastic_shared_embeddings_replacement_prob): the step trains on the noised inputs -- loss and gradients equal
    autograd of the oracle graph fed with the oracle's own noised batch (same draws)."""
    import transformers4rec_b200.torch as tr
    from transformers4rec_b200.training import FusedTrainingStep
    D.install(monkeypatch)
    oracle, model = make_pair(CARDS, {"item_id/list": 32, "category/list": 32}, "item_id/list", CONT, 32, 2, 1, 8,
                              device="cpu", weight_scale=0.08)
    oracle.train(False)
    inputs = model.heads[0].body[0]
    ssn = tr.StochasticSwapNoise(schema=inputs.schema, pad_token=0, replacement_prob=0.3)
    inputs.pre = ssn
    model.train()
    model.heads[0].body[1].transformer.eval()   # the noise is what is under test here; dropout has its own tests below
    B, L = 6, 8
    batch = synth_batch(B, L, CARDS, CONT, seed=11)
    mask = batch["item_id/list"] != 0
    g = torch.Generator().manual_seed(1)
    draws_ssn, noisy = {}, {}
    for k, v in batch.items():
        u, perm = torch.rand(v.shape, generator=g), torch.randperm(int(mask.sum()), generator=g)
        draws_ssn[k] = (u, perm)
        noisy[k] = O.stochastic_swap_noise(v, mask, u, perm, 0.3)
    assert any(not torch.equal(noisy[k], batch[k]) for k in batch)
    ssn.set_draws(draws_ssn)
    u, draws = mlm_draws(B, L)
    inputs.masking.set_draws(u)
    ref_loss = _oracle_grads(oracle, noisy, draws)
    step = FusedTrainingStep(model, head_chunk=512)
    for p in model.parameters():
        p.grad = None
    loss = step.forward(batch)
    step.backward()
    assert abs(loss.item() - ref_loss) < 1e-4
    for name, po, pm in _pairs(oracle, model):
        if po.grad is None and pm.grad is None:
            continue
        err = (pm.grad - po.grad.reshape(pm.grad.shape)).abs().max().item()
        assert err < 3e-4 * max(1.0, po.grad.abs().max().item()), (name, err)


def _kernel_mask(shape, p, seed, site):
    """The dropout mask the kernels use, as a tensor of ``shape`` in the kernels' flat index order."""
    return D._REAL_DROPOUT(torch.ones(int(torch.Size(shape).numel())), p, seed, site, _on_host=True).reshape(shape)


def test_dropout_mask_is_counter_based():
    """t4r_train_dropout: keep probability 1 - p, scale 1 / (1 - p), a pure function of (seed, site, index)."""
    from transformers4rec_b200 import ops
    tw = ops.host_twin("dropout")
    x = torch.randn(50_000)
    y = tw(x, 0.3, 99, 17)
    kept = y != 0
    assert abs(kept.float().mean().item() - 0.7) < 0.01
    assert torch.allclose(y[kept], x[kept] / 0.7)
    assert torch.equal(y, tw(x, 0.3, 99, 17)) and not torch.equal(y, tw(x, 0.3, 99, 18)) and not torch.equal(y, tw(x, 0.3, 100, 17))
    assert tw(x, 0.0, 1, 1) is x
    # a longer tensor under the same key extends the same stream (the backward of a slice sees the forward's mask)
    assert torch.equal(tw(torch.ones(100), 0.5, 5, 3), tw(torch.ones(1000), 0.5, 5, 3)[:100])


@pytest.mark.parametrize("arch", ["xlnet", "gpt2"])
def test_restated_dropout_sites_are_huggingfaces(arch):
    """The oracle's restated encoders mark every place HF calls dropout (``drop(site, tensor)``); held to HF's own
    forward in train() mode with ``torch.nn.functional.dropout`` replaced by the same masks, identified by call order."""
    import torch.nn.functional as F
    torch.manual_seed(3)
    B, L, d, H, NL, p = 3, 6, 16, 2, 2, 0.3      # the rate the reference builds HF with (config/transformer.py:217-260, :432-482)
    hf = (O.build_hf_xlnet(d, H, NL) if arch == "xlnet" else O.build_hf_gpt2(d, H, NL, L)).train()
    if arch == "gpt2":
        hf.config._attn_implementation = "eager"     # the eager path applies F.dropout to the probabilities (HF:gpt2:66)
    with torch.no_grad():
        for prm in hf.parameters():
            prm.add_(torch.randn_like(prm) * 0.1)
    x = torch.randn(B, L, d)
    if arch == "xlnet":   # call order of XLNetModel.forward / XLNetLayer: input, pos_emb, then per layer prob, o-proj, act, ff
        order = [0, None] + [s for l in range(NL) for s in (16 * (l + 1) + 1, 16 * (l + 1) + 2, 16 * (l + 1) + 3, 16 * (l + 1) + 4)] + [5]
    else:
        order = [0] + [s for l in range(NL) for s in (16 * (l + 1) + 1, 16 * (l + 1) + 2, 16 * (l + 1) + 4)]
    calls = []

    def to_kernel_layout(t, site):   # HF's XLNet works on [L, B, d]; the probabilities are [B, H, L, L] in both
        return t.transpose(0, 1) if (arch == "xlnet" and t.dim() == 3) else t

    def fake_dropout(inp, p=0.5, training=True, inplace=False):
        site = order[len(calls)]
        calls.append(tuple(inp.shape))
        if site is None or not training:
            return inp
        k = to_kernel_layout(inp, site)
        out = k * _kernel_mask(k.shape, p, 7, site)
        return to_kernel_layout(out, site)
    orig = F.dropout
    F.dropout = fake_dropout
    try:
        with torch.no_grad():
            got = hf(inputs_embeds=x)[0]
    finally:
        F.dropout = orig
    assert len(calls) == len(order), (len(calls), len(order))

    def drop(site, t):
        return t if site % 16 == 0 and site > 0 else t * _kernel_mask(t.shape, p, 7, site)   # R site: identity (HF drops pos_emb)
    sd = dict(hf.named_parameters())
    with torch.no_grad():
        want = (O.xlnet_forward_restated(x, sd, NL, H, drop=drop) if arch == "xlnet" else
                O.gpt2_forward_restated(x, sd, NL, H, eps=hf.config.layer_norm_epsilon, drop=drop))
    assert (got - want).abs().max().item() < 2e-5


def test_attention_kernels_with_dropped_probabilities():
    """The per-(session, head) attention forward / backward with dropout of the probabilities (host twins = the
    kernels' per-item code) against autograd of the plain formula with the same mask: XLNet relative, XLNet two-stream
    (PLM) and GPT-2 causal forms."""
    from transformers4rec_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, L, Hh, d = 3, 7, 2, 16
    drop = (0.3, 4242, 33)
    qkv, dout = torch.randn(B * L, 3 * d, generator=g), torch.randn(B * L, d, generator=g)
    R, rw, rr = torch.randn(2 * L, d, generator=g), torch.randn(d, generator=g), torch.randn(d, generator=g)
    fwd, bwd, cbwd = ops.host_twin("attn_drop_fwd"), ops.host_twin("xlnet_attn_bwd"), ops.host_twin("causal_attn_bwd")
    assert (fwd(qkv, R, rw, rr, B, L, Hh, drop) - D.xlnet_attn_fwd(qkv, R, rw, rr, B, L, Hh, drop=drop)).abs().max() < 1e-5
    assert (fwd(qkv, R, rw, rr, B, L, Hh, drop) - D.xlnet_attn_fwd(qkv, R, rw, rr, B, L, Hh)).abs().max() > 1e-2
    for got, want in zip(bwd(qkv, R, rw, rr, dout, B, L, Hh, drop=drop), D.xlnet_attn_bwd(qkv, R, rw, rr, dout, B, L, Hh, drop=drop)):
        assert (got - want).abs().max().item() < 1e-4
    assert (fwd(qkv, None, None, None, B, L, Hh, drop) - D.causal_attn_fwd(qkv, B, L, Hh, drop=drop)).abs().max() < 1e-5
    assert (cbwd(qkv, dout, B, L, Hh, drop=drop) - D.causal_attn_bwd(qkv, dout, B, L, Hh, drop=drop)).abs().max() < 1e-4
    qkv2, dout2 = torch.randn(2 * B * L, 3 * d, generator=g), torch.randn(2 * B * L, d, generator=g)
    pm = (torch.rand(B, L, L, generator=g) < 0.4)
    assert (fwd(qkv2, R, rw, rr, B, L, Hh, drop, plm_mask=pm) -
            D.xlnet_attn_plm_fwd(qkv2, R, rw, rr, B, L, Hh, pm, drop=drop)).abs().max() < 1e-5
    for got, want in zip(bwd(qkv2, R, rw, rr, dout2, B, L, Hh, plm_mask=pm, drop=drop),
                         D.xlnet_attn_bwd(qkv2, R, rw, rr, dout2, B, L, Hh, plm_mask=pm, drop=drop)):
        assert (got - want).abs().max().item() < 1e-4


@pytest.mark.parametrize("arch,masking", [("xlnet", "mlm"), ("gpt2", "clm")])
def test_training_step_applies_dropout_in_train_mode(monkeypatch, arch, masking):
    """With the encoder in train() mode the fused step applies dropout at HF's sites (masks regenerated in the
    backward): loss and every gradient equal autograd of the oracle graph whose encoder is the restated forward with
    the SAME masks; in eval() mode (and at rate 0) nothing is dropped."""
    from transformers4rec_b200.training import FusedTrainingStep
    D.install(monkeypatch)
    dims = {"item_id/list": 32, "category/list": 32}
    oracle, model = make_pair(CARDS, dims, "item_id/list", CONT, 32, 2, 2, 8, arch=arch, masking=masking, device="cpu",
                              weight_scale=0.08)
    B, L, NL, Hh, p, seed = 6, 8, 2, 2, 0.3, 1234
    batch = synth_batch(B, L, CARDS, CONT, seed=9)
    u, draws = mlm_draws(B, L)
    model.heads[0].body[0].masking.set_draws(u)
    enc = model.heads[0].body[1].transformer
    assert float(enc.config.dropout if arch == "xlnet" else enc.config.resid_pdrop) == 0.3   # the reference's default

    def drop(site, t):
        return t * _kernel_mask(t.shape, p, seed, site)

    def encoder_with_masks(hf, x):
        sd = dict(hf.named_parameters())
        if arch == "xlnet":
            return O.xlnet_forward_restated(x, sd, NL, Hh, drop=drop)
        return O.gpt2_forward_restated(x, sd, NL, Hh, eps=hf.config.layer_norm_epsilon, drop=drop)
    monkeypatch.setattr(O, "hf_encoder_forward", encoder_with_masks)
    ref_loss = _oracle_grads(oracle, batch, draws)
    step = FusedTrainingStep(model, head_chunk=512).set_dropout_seed(seed)
    enc.train()
    for prm in model.parameters():
        prm.grad = None
    loss = step.forward(batch)
    step.backward()
    assert abs(loss.item() - ref_loss) < 1e-4
    n = 0
    for name, po, pm in _pairs(oracle, model):
        if po.grad is None and pm.grad is None:
            continue
        err = (pm.grad - po.grad.reshape(pm.grad.shape)).abs().max().item()
        assert err < 3e-4 * max(1.0, po.grad.abs().max().item()), (name, err)
        n += 1
    assert n > 10
    # the next step draws new masks; eval() mode drops nothing
    loss2 = step.forward(batch)
    assert abs(loss2.item() - loss.item()) > 1e-6
    enc.eval()
    monkeypatch.setattr(O, "hf_encoder_forward", lambda hf, x: (
        O.xlnet_forward_restated(x, dict(hf.named_parameters()), NL, Hh) if arch == "xlnet" else
        O.gpt2_forward_restated(x, dict(hf.named_parameters()), NL, Hh, eps=hf.config.layer_norm_epsilon)))
    with torch.no_grad():
        plain = float(oracle(batch, training=True, draws=draws)["loss"])
    assert abs(step.forward(batch).item() - plain) < 1e-4


def test_plm_attention_backward_on_fully_masked_rows():
    """A session whose every item is a target (upstream allows it, masking.py:646) leaves the first target of the
    permutation with nothing to attend to in the query stream: HF's ``score - 1e30 * mask`` softmaxes that row to the
    uniform distribution and autograd passes gradient through the masked scores (d/d score = 1).  The kernel's real
    per-item code follows that: gradients equal autograd of the formula, fully masked rows included (also L = 1)."""
    import random
    from transformers4rec_b200 import ops
    twin = ops.host_twin("xlnet_attn_bwd")
    g = torch.Generator().manual_seed(0)
    random.seed(1)
    for trial in range(12):
        B, L, Hh, dh = random.choice([1, 2, 5]), random.choice([1, 2, 3, 7, 12]), random.choice([1, 2]), 8
        d = Hh * dh
        R = torch.randn(2 * L, d, generator=g)
        rw, rr = torch.randn(d, generator=g), torch.randn(d, generator=g)
        pm = torch.rand(B, L, L, generator=g) < 0.4
        pm[0, 0, :] = True                               # row 0 of session 0 sees nobody (g stream; h keeps its diagonal)
        qkv = torch.randn(2 * B * L, 3 * d, generator=g)
        dout = torch.randn(2 * B * L, d, generator=g)
        want = D.xlnet_attn_bwd(qkv, R, rw, rr, dout, B, L, Hh, plm_mask=pm)      # autograd of the formula
        got = twin(qkv, R, rw, rr, dout, B, L, Hh, plm_mask=pm)
        for a, b in zip(want, got):
            assert torch.isfinite(b).all()
            assert (a - b).abs().max().item() < 1e-5 * max(1.0, a.abs().max().item())
