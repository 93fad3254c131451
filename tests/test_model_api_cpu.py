"""The next-item model through the reference's public API -- the calls of the reference's own model tests
(tests/unit/torch/model/test_model.py:296-407) on its testing schema -- for the maskings on this path.  CPU, kernels
replaced by the test doubles: what is exercised is construction, the module protocol and the outputs' shapes / types."""
import pytest
import torch

import _ops_double as D
import transformers4rec_b200.torch as tr
from test_abi_and_host import _testing_schema


def _batch(schema, B=16, L=20, seed=0):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(2, L + 1, (B,), generator=g)
    valid = torch.arange(L)[None] < lens[:, None]
    out = {}
    for col in schema:
        shape = (B, L) if col.is_list else (B,)
        v = torch.randint(1, col.int_max + 1, shape, generator=g) if col.int_max else torch.rand(shape, generator=g)
        out[col.name] = torch.where(valid, v, torch.zeros_like(v)) if col.is_list else v
    return out


@pytest.mark.parametrize("masking", ["causal", "mlm", "plm"])
def test_eval_metrics_with_masking(monkeypatch, masking):
    """test_model.py:342-359: default task (own output layer), continuous_projection, training forward, metrics from the
    materialised predictions / labels."""
    D.install(monkeypatch)
    schema = _testing_schema(tr)
    inputs = tr.TabularSequenceFeatures.from_schema(schema, max_sequence_length=20, continuous_projection=64, d_output=64,
                                                    masking=masking)
    model = tr.XLNetConfig.build(64, 4, 2, 20).to_torch_model(inputs, tr.NextItemPredictionTask())
    with torch.no_grad():
        out = model(_batch(schema), training=True)
        result = model.calculate_metrics(predictions=out["predictions"], targets=out["labels"])
    assert result is not None and len(result) >= 1
    assert out["predictions"].shape == (out["labels"].numel(), 51997)


@pytest.mark.parametrize("d_model", [32, 64, 128])
def test_with_d_model_different_from_item_dim(monkeypatch, d_model):
    """test_model.py:362-374"""
    D.install(monkeypatch)
    schema = _testing_schema(tr)
    inputs = tr.TabularSequenceFeatures.from_schema(schema, max_sequence_length=20, continuous_projection=64,
                                                    d_output=d_model, masking="mlm")
    model = tr.XLNetConfig.build(d_model, 4, 2, 20).to_torch_model(inputs, tr.NextItemPredictionTask(weight_tying=True))
    with torch.no_grad():
        out = model(_batch(schema), training=True)
    assert torch.isfinite(out["loss"])


def test_with_next_item_pred_sampled_softmax(monkeypatch):
    """test_model.py:377-389"""
    D.install(monkeypatch)
    schema = _testing_schema(tr)
    inputs = tr.TabularSequenceFeatures.from_schema(schema, max_sequence_length=20, continuous_projection=64, d_output=32,
                                                    masking="mlm")
    task = tr.NextItemPredictionTask(weight_tying=True, sampled_softmax=True, max_n_samples=1000)
    model = tr.XLNetConfig.build(32, 4, 2, 20).to_torch_model(inputs, task)
    with torch.no_grad():
        out = model(_batch(schema), training=True)
    assert torch.isfinite(out["loss"]) and out["predictions"].shape[1] <= 1001
    # model/prediction_task.py:693-696: the positive is column 0 of the sampled logits, the returned targets are zeros
    assert out["labels"].shape == (out["predictions"].shape[0],) and not out["labels"].any()


@pytest.mark.parametrize("masking", ["causal", "mlm", "plm"])
def test_output_shape_mode_eval(monkeypatch, masking):
    """test_model.py:392-407: one prediction row per session in evaluation"""
    D.install(monkeypatch)
    schema = _testing_schema(tr)
    inputs = tr.TabularSequenceFeatures.from_schema(schema, max_sequence_length=20, d_output=64, masking=masking)
    model = tr.XLNetConfig.build(d_model=64, n_head=8, n_layer=2, total_seq_length=20).to_torch_model(
        inputs, tr.NextItemPredictionTask(weight_tying=True))
    batch = _batch(schema)
    with torch.no_grad():
        out = model(batch, training=False, testing=True)
    assert out["predictions"].shape[0] == batch["item_id/list"].size(0)


@pytest.mark.parametrize("arch", ["xlnet", "gpt2"])
def test_item_prediction_model_from_config_inference(monkeypatch, arch):
    """test_model.py:296-316: ``model(batch)`` (inference) returns [B, target_dim] scores"""
    D.install(monkeypatch)
    schema = _testing_schema(tr)
    inputs = tr.TabularSequenceFeatures.from_schema(schema, max_sequence_length=20, continuous_projection=64, d_output=128,
                                                    masking="causal")
    task = tr.NextItemPredictionTask()
    cfg = (tr.XLNetConfig if arch == "xlnet" else tr.GPT2Config).build(128, 4, 2, 20)
    model = cfg.to_torch_model(inputs, task)
    with torch.no_grad():
        out = model(_batch(schema))
    assert out.dim() == 2 and out.size(1) == task.target_dim == 51997


def test_state_dict_keys_follow_the_reference_module_tree_and_save_load_round_trip(monkeypatch, tmp_path):
    """Checkpoint surface ("state-dict key names ... are the compatibility surface"; model/base.py:839-922).
    The key names below are what the reference's module tree produces for the same construction: ``to_merge`` /
    ``embedding_tables`` (features/sequence.py, features/embedding.py), ``projection_module.0.0`` (MLPBlock ->
    DenseBlock -> Linear), ``_masking.masked_item_embedding`` (masking.py:103-108), HF's XLNet names under
    ``body.1.transformer``, and the task's aliases set in NextItemPredictionTask.build (prediction_task.py:380-417:
    ``embeddings``, ``item_embedding_table``, ``masking``, ``task_block``)."""
    D.install(monkeypatch)
    schema = tr.Schema([tr.ColumnSchema.create_categorical("item_id/list", 300, tags=[tr.Tags.ITEM_ID]),
                        tr.ColumnSchema.create_categorical("category/list", 40),
                        tr.ColumnSchema.create_continuous("price/list")])

    def build():
        inputs = tr.TabularSequenceFeatures.from_schema(schema, max_sequence_length=8, d_output=32, masking="mlm")
        return tr.XLNetConfig.build(32, 2, 1, 8).to_torch_model(inputs, tr.NextItemPredictionTask(weight_tying=True))
    model = build()
    keys = set(model.state_dict().keys())
    expected = {
        "heads.0.body.0.to_merge.categorical_module.embedding_tables.item_id/list.weight",
        "heads.0.body.0.to_merge.categorical_module.embedding_tables.category/list.weight",
        "heads.0.body.0.projection_module.0.0.weight", "heads.0.body.0.projection_module.0.0.bias",
        "heads.0.body.0._masking.masked_item_embedding",
        "heads.0.body.1.masking.masked_item_embedding",
        "heads.0.body.1.transformer.mask_emb", "heads.0.body.1.transformer.word_embedding.weight",
        "heads.0.prediction_task_dict.next-item.embeddings.embedding_tables.item_id/list.weight",
        "heads.0.prediction_task_dict.next-item.item_embedding_table.weight",
        "heads.0.prediction_task_dict.next-item.masking.masked_item_embedding",
        "heads.0.prediction_task_dict.next-item.task_block.0.0.weight",
        "heads.0.prediction_task_dict.next-item.task_block.0.0.bias",
    }
    for leaf in ("q", "k", "v", "o", "r", "r_r_bias", "r_s_bias", "r_w_bias", "seg_embed", "layer_norm.weight",
                 "layer_norm.bias"):
        expected.add("heads.0.body.1.transformer.layer.0.rel_attn." + leaf)
    for leaf in ("layer_norm.weight", "layer_norm.bias", "layer_1.weight", "layer_1.bias", "layer_2.weight", "layer_2.bias"):
        expected.add("heads.0.body.1.transformer.layer.0.ff." + leaf)
    assert expected <= keys, sorted(expected - keys)
    # save -> torch.load -> Model.load over freshly built heads: same weights, same loss
    with torch.no_grad():
        for p in model.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    model.save(tmp_path, model_name="m")
    sd = torch.load(tmp_path / "m.pt")
    other = tr.Model.load(sd, build().heads[0], max_sequence_length=8)
    assert all(torch.equal(a, b) for a, b in zip(model.state_dict().values(), other.state_dict().values()))
    batch = {"item_id/list": torch.randint(1, 301, (4, 8)), "category/list": torch.randint(1, 41, (4, 8)),
             "price/list": torch.rand(4, 8)}
    u = torch.rand(4, 10, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        for m in (model, other):
            m.heads[0].body[0].masking.set_draws(u)
        a, b = model(dict(batch), training=True)["loss"], other(dict(batch), training=True)["loss"]
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        tr.Model.load([1, 2], build().heads[0])
    with pytest.raises(RuntimeError):      # strict by default, as in the reference
        tr.Model.load({k: v for k, v in sd.items() if "projection_module" not in k}, build().heads[0])
