"""GPU parity tests for the widened input block and the remaining ranking metrics
(N2): CUDA (through the C ABI) vs vectors produced by the upstream code
(tests/golden/reference_vectors_n4.pt) and vs the oracle on larger seeded inputs.  Integer / copy
work bit-exact; fp32 results within 1e-5 (pure fp32 CUDA-core arithmetic here, no tensor cores)."""
import os

import pytest
import torch

import t4r_oracle as O
from _util import synth_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors_n4.pt"), weights_only=False)


@pytest.fixture(scope="module")
def ops():
    from transformers4rec_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def lib():
    from transformers4rec_b200 import _lib
    return _lib


# --------------------------------------------------------------------------- #
# kernels against the upstream vectors
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("mode", ["concat", "element-wise-sum", "element-wise-sum-item-multi"])
def test_aggregations_match_upstream(ops, lib, gold, mode):
    a = gold["aggregation"]
    names = sorted(a["features"].keys())
    B, L, D = a["features"]["item"].shape
    agg = {"concat": lib.AGG_CONCAT, "element-wise-sum": lib.AGG_SUM, "element-wise-sum-item-multi": lib.AGG_SUM_ITEM_MULTI}[mode]
    feats, col = [], 0
    for n in names:
        v = a["features"][n].cuda()
        feats.append(dict(kind=lib.FEAT_DENSE, dim=D, col=col if agg == lib.AGG_CONCAT else 0, input=v,
                          per_session=(v.dim() == 2)))
        col += D
    C = col if agg == lib.AGG_CONCAT else D
    y, yp, _ = ops.input_block(feats, B * L, L, C, agg=agg, item_feature=names.index("item"), want_planes=True)
    want = a[mode].reshape(B * L, C)
    if mode == "concat":
        assert torch.equal(y.cpu(), want)  # pure copy
    else:
        assert (y.cpu() - want).abs().max().item() < 1e-5
    assert ((yp[0].float() + yp[1].float())[:, :C].cpu() - y.cpu()).abs().max().item() < 1e-4
    assert (yp[:, :, C:] == 0).all()


def test_layer_norm_and_soft_embedding_match_upstream(ops, lib, gold):
    import transformers4rec_b200.torch as tr
    ln = gold["layer_norm"]
    mod = tr.TabularLayerNorm({k: v.shape[-1] for k, v in ln["inputs"].items() if k in ln["params"]}).cuda()
    with torch.no_grad():
        for k, (g, b) in ln["params"].items():
            mod.feature_layer_norm[k].weight.copy_(g)
            mod.feature_layer_norm[k].bias.copy_(b)
        got = mod({k: v.cuda() for k, v in ln["inputs"].items()})
    for k, want in ln["outputs"].items():
        assert (got[k].cpu() - want).abs().max().item() < 1e-5, k
    se = gold["soft_embedding"]
    m = tr.SoftEmbedding(se["table"].shape[0], se["table"].shape[1]).cuda()
    with torch.no_grad():
        m.embedding_table.weight.copy_(se["table"])
        m.projection_layer.weight.copy_(se["proj_weight"])
        m.projection_layer.bias.copy_(se["proj_bias"])
        got = m(se["x"].cuda())
    assert got.shape == se["out"].shape
    assert (got.cpu() - se["out"]).abs().max().item() < 1e-5


def test_swap_noise_bit_exact_vs_upstream(ops, gold):
    sn = gold["swap_noise"]
    for name, c in sn["cases"].items():
        mask = None if c["mask"] is None else c["mask"].cuda()
        got = ops.swap_noise(c["values"].cuda(), mask, c["u"].cuda(), c["perm"].cuda(), sn["replacement_prob"])
        assert torch.equal(got.cpu(), c["out"]), name


def test_swap_noise_large_vs_oracle(ops):
    g = torch.Generator().manual_seed(5)
    B, L = 2048, 20  # 40 iterations of the single-block scan
    ids = synth_batch(B, L, {"i": 100000}, seed=3)["i"]
    mask = ids != 0
    for vals in (ids, torch.rand((B, L), generator=g)):
        u = torch.rand((B, L), generator=g)
        perm = torch.randperm(int(mask.sum()), generator=g)
        want = O.stochastic_swap_noise(vals, mask, u, perm, 0.2)
        got = ops.swap_noise(vals.cuda(), mask.cuda(), u.cuda(), perm.cuda(), 0.2)
        assert torch.equal(got.cpu(), want)
        assert torch.equal(got.cpu()[~mask], vals[~mask])


def test_ranking_metrics_from_ranks_match_upstream(gold):
    import transformers4rec_b200.torch as tr
    m = gold["metrics"]
    for name, want in m["results"].items():
        metric = tr.ranking_metrics_registry[name](top_ks=m["ks"], labels_onehot=True)
        got = metric(m["scores"].cuda(), m["labels"].cuda())
        assert (got.cpu() - want).abs().max().item() < 1e-6, name
    # known answers of tests/unit/torch/test_ranking_metrics.py:49-85 (one-hot labels given as a matrix)
    sc = m["known_answer_scores"].float().cuda()
    oh = m["known_answer_onehot"].cuda()
    mrr = tr.MeanReciprocalRankAt(top_ks=[1, 2, 3, 4], labels_onehot=False)(sc, oh)
    assert (mrr.cpu() - m["known_answer_mrr"]).abs().max().item() < 1e-3


# --------------------------------------------------------------------------- #
# module level: the input block configured like the paper recipes, vs the oracle composition
# --------------------------------------------------------------------------- #
def _schema(tr):
    return tr.Schema([tr.ColumnSchema.create_categorical("item_id/list", 800, tags=[tr.Tags.ITEM_ID]),
                      tr.ColumnSchema.create_categorical("category/list", 40),
                      tr.ColumnSchema.create_categorical("user_country", 30, is_list=False),
                      tr.ColumnSchema.create_continuous("price/list"),
                      tr.ColumnSchema.create_continuous("rel_time/list")])


def _batch(B, L, seed=4):
    batch = synth_batch(B, L, {"item_id/list": 801, "category/list": 41}, continuous=("price/list", "rel_time/list"), seed=seed)
    batch["user_country"] = torch.randint(1, 31, (B,), generator=torch.Generator().manual_seed(seed))
    return batch


def _oracle_features(inputs, batch, B, L):
    """per-feature tensors the way the reference's sub-modules produce them (CPU, from the module's weights)."""
    cm, cont = inputs.categorical_module, inputs.continuous_module
    feats = {}
    for name in cm.feature_config:
        x = torch.nn.functional.embedding(batch[name], cm.embedding_tables[name].weight.detach().cpu(), padding_idx=0)
        ln = cm.post.params(name) if cm.post is not None else None
        feats[name] = O.tabular_layer_norm(x, ln[0].detach().cpu(), ln[1].detach().cpu()) if ln is not None else x
    return feats, cont


@pytest.mark.parametrize("aggregation", ["concat", "element-wise-sum", "element-wise-sum-item-multi"])
def test_soft_embeddings_layer_norm_and_aggregations(aggregation):
    import transformers4rec_b200.torch as tr
    torch.manual_seed(21)
    B, L, D = 13, 10, 16
    inputs = tr.TabularSequenceFeatures.from_schema(
        _schema(tr), max_sequence_length=L, aggregation=aggregation, continuous_soft_embeddings=True,
        soft_embedding_dim_default=D, embedding_dim_default=D, post="layer-norm", d_output=32, masking="mlm").cuda()
    with torch.no_grad():  # non-trivial LayerNorm parameters
        for mod in list(inputs.categorical_module.post.feature_layer_norm.values()) + list(
                inputs.continuous_module.post.feature_layer_norm.values()):
            mod.weight.uniform_(0.5, 1.5)
            mod.bias.normal_(0.0, 0.3)
    batch = _batch(B, L)
    u = torch.rand((B, L + 2), generator=torch.Generator().manual_seed(9))
    inputs.masking.set_draws(u.cuda())
    with torch.no_grad():
        x = inputs({k: v.cuda() for k, v in batch.items()}, training=True).cpu()
        feats, cont = _oracle_features(inputs, batch, B, L)
        for name in cont.features:
            se = cont.embedding_tables[name]
            y = O.soft_embedding(batch[name], se.projection_layer.weight.cpu(), se.projection_layer.bias.cpu(),
                                 se.embedding_table.weight.cpu())
            g, b = cont.post.params(name)
            feats[name] = O.tabular_layer_norm(y, g.cpu(), b.cpu())
        ref = O.aggregate(feats, aggregation, "item_id/list")
        assert ref.shape[-1] == inputs._layout()[1]
        lin = inputs.projection_module[0][0]
        ref = O.project_relu(ref, lin.weight.cpu(), lin.bias.cpu())
        mask, _ = O.mlm_compute_masked_targets(batch["item_id/list"], True, False, u_bern=u[:, :L], u_force=u[:, L],
                                               u_unmask=u[:, L + 1])
        ref = O.mlm_apply_mask_to_inputs(ref, mask, inputs.masking.masked_item_embedding.detach().cpu(), True, False)
    assert (x - ref).abs().max().item() < 1e-4


def test_continuous_projection_enters_as_one_feature():
    import transformers4rec_b200.torch as tr
    torch.manual_seed(22)
    B, L = 11, 9
    inputs = tr.TabularSequenceFeatures.from_schema(_schema(tr), max_sequence_length=L, continuous_projection=[24, 16],
                                                    aggregation="concat", embedding_dim_default=16).cuda()
    batch = _batch(B, L, seed=6)
    with torch.no_grad():
        x = inputs({k: v.cuda() for k, v in batch.items()}).cpu()
        feats, cont = _oracle_features(inputs, batch, B, L)
        c = torch.cat([batch[n].unsqueeze(-1) for n in sorted(cont.features)], dim=-1)
        for blk in cont.mlp:
            c = O.project_relu(c, blk[0].weight.cpu(), blk[0].bias.cpu())
        feats["continuous_projection"] = c
        ref = O.aggregate(feats, "concat")
    assert x.shape == (B, L, 16 * 3 + 16)
    assert (x - ref).abs().max().item() < 1e-4


def test_stochastic_swap_noise_as_pre_transform():
    import transformers4rec_b200.torch as tr
    torch.manual_seed(23)
    B, L = 12, 8
    schema = _schema(tr)
    ssn = tr.StochasticSwapNoise(schema=schema, pad_token=0, replacement_prob=0.3)
    inputs = tr.TabularSequenceFeatures.from_schema(schema, max_sequence_length=L, aggregation="concat",
                                                    embedding_dim_default=8, pre=ssn).cuda().train()
    batch = _batch(B, L, seed=7)
    mask = batch["item_id/list"] != 0
    g = torch.Generator().manual_seed(1)
    draws, noisy = {}, {}
    for k, v in batch.items():
        eff = mask[:, 0] if v.dim() == 1 else mask
        u, perm = torch.rand(v.shape, generator=g), torch.randperm(int(eff.sum()), generator=g)
        draws[k] = (u.cuda(), perm.cuda())
        noisy[k] = O.stochastic_swap_noise(v, mask, u, perm, 0.3)
    ssn.set_draws(draws)
    with torch.no_grad():
        x = inputs({k: v.cuda() for k, v in batch.items()}).cpu()
        feats, cont = _oracle_features(inputs, noisy, B, L)
        for n in cont.features:
            feats[n] = noisy[n].unsqueeze(-1)
        ref = O.aggregate(feats, "concat")
    assert torch.equal(x, ref)  # gather + copies only
    inputs.eval()  # transformations.py:58-59: identity outside training
    with torch.no_grad():
        x_eval = inputs({k: v.cuda() for k, v in batch.items()}).cpu()
        feats, _ = _oracle_features(inputs, batch, B, L)
        for n in cont.features:
            feats[n] = batch[n].unsqueeze(-1)
    assert torch.equal(x_eval, O.aggregate(feats, "concat"))


def test_reference_feature_test_shape_203():
    """tests/unit/torch/features/test_sequential.py:217-223 on the testing schema: [100, 20, 203]."""
    import transformers4rec_b200.torch as tr
    from test_abi_and_host import _testing_schema
    schema = _testing_schema(tr)
    tab = tr.TabularSequenceFeatures.from_schema(schema, aggregation="concat").cuda()
    g = torch.Generator().manual_seed(0)
    batch = {}
    for col in schema:
        shape = (100, 20) if col.is_list else (100,)
        batch[col.name] = (torch.randint(1, col.int_max + 1, shape, generator=g) if col.int_max
                           else torch.rand(shape, generator=g)).cuda()
    out = tab(batch)
    assert list(out.shape) == [100, 20, 203]


def test_out_of_range_ids_surface_as_the_reference_error(monkeypatch):
    """nn.Embedding raises IndexError for an id outside the table; the fused gather flags it (and reads row 0)."""
    import transformers4rec_b200.torch as tr
    schema = tr.Schema([tr.ColumnSchema.create_categorical("item_id/list", 100, tags=[tr.Tags.ITEM_ID]),
                        tr.ColumnSchema.create_categorical("category/list", 10)])
    for kwargs in ({}, {"post": "layer-norm"}):  # specialised gather and general input-block kernel
        tab = tr.TabularSequenceFeatures.from_schema(schema, max_sequence_length=6, aggregation="concat", **kwargs).cuda()
        batch = synth_batch(4, 6, {"item_id/list": 101, "category/list": 11}, seed=1)
        tab({k: v.cuda() for k, v in batch.items()})
        tab.check_ids()  # in range: no error
        bad = {k: v.clone() for k, v in batch.items()}
        bad["category/list"][2, 1] = 11  # one past the last row of the 11-row table
        tab({k: v.cuda() for k, v in bad.items()})
        with pytest.raises(IndexError, match="index out of range in self"):
            tab.check_ids()
        monkeypatch.setenv("T4R_CHECK_IDS", "1")
        with pytest.raises(IndexError):
            tab({k: v.cuda() for k, v in bad.items()})
        monkeypatch.delenv("T4R_CHECK_IDS")
