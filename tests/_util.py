"""Shared helpers for the parity tests: synthetic yoochoose-shaped batches
(SURVEY §8d), oracle <-> product weight transfer."""
import torch

import t4r_oracle as O


def synth_batch(B, L, cardinalities, continuous=(), seed=0, min_len=2, device="cpu"):
    """Right-padded sessions, len ~ U{min_len..L}, ids uniform in [1, card) (utils/schema_utils.py:78-96)."""
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(min_len, L + 1, (B,), generator=g)
    pos = torch.arange(L).unsqueeze(0)
    valid = pos < lens.unsqueeze(1)
    batch = {}
    for name, card in cardinalities.items():
        ids = torch.randint(1, card, (B, L), generator=g)
        batch[name] = torch.where(valid, ids, torch.zeros_like(ids))
    for name in continuous:
        v = torch.rand((B, L), generator=g)
        batch[name] = torch.where(valid, v, torch.zeros_like(v))
    return {k: v.to(device) for k, v in batch.items()}


def mlm_draws(B, L, seed=2):
    g = torch.Generator().manual_seed(seed)
    u = torch.rand((B, L + 2), generator=g)
    return u, {"u_bern": u[:, :L], "u_force": u[:, L], "u_unmask": u[:, L + 1]}


def make_pair(cardinalities, embedding_dims, item_id, continuous, d_model, n_head, n_layer, L, arch="xlnet",
              masking="mlm", sampled=False, max_n_samples=100, seed=1, device="cuda", weight_scale=None):
    """Build the oracle graph and the product model with identical weights."""
    import transformers4rec_b200.torch as tr

    torch.manual_seed(seed)
    oracle = O.OracleSessionModel(cardinalities=cardinalities, embedding_dims=embedding_dims, item_id=item_id,
                                  continuous=continuous, d_model=d_model, n_head=n_head, n_layer=n_layer,
                                  max_seq_len=L, arch=arch, masking=masking, sampled_softmax=sampled,
                                  max_n_samples=max_n_samples).eval()
    if weight_scale is not None:
        # "trained-like" weights: HF initialises with std 0.01, which makes the encoder nearly linear
        with torch.no_grad():
            for n, p in oracle.transformer.named_parameters():
                if p.ndim >= 2 and "layer_norm" not in n and "ln_" not in n:
                    p.normal_(0.0, weight_scale)
                elif "bias" in n and "layer_norm" not in n and "ln_" not in n:
                    p.normal_(0.0, weight_scale)
            oracle.masked_item_embedding.normal_(0.0, 0.5)
    cols = [tr.ColumnSchema.create_categorical(n, c - 1, tags=[tr.Tags.ITEM_ID] if n == item_id else None)
            for n, c in cardinalities.items()]
    cols += [tr.ColumnSchema.create_continuous(n) for n in continuous]
    schema = tr.Schema(cols)
    inputs = tr.TabularSequenceFeatures.from_schema(schema, max_sequence_length=L, d_output=d_model,
                                                    masking=masking, embedding_dims=embedding_dims)
    cfg_cls = tr.XLNetConfig if arch == "xlnet" else tr.GPT2Config
    cfg = cfg_cls.build(d_model=d_model, n_head=n_head, n_layer=n_layer, total_seq_length=L)
    task = tr.NextItemPredictionTask(weight_tying=True, sampled_softmax=sampled, max_n_samples=max_n_samples)
    model = cfg.to_torch_model(inputs, task)
    copy_weights(oracle, model)
    return oracle, model.to(device).eval()


def copy_weights(oracle, model):
    head = model.heads[0]
    inputs, tblock = head.body[0], head.body[1]
    with torch.no_grad():
        for name in oracle.table_names:
            inputs.categorical_module.embedding_tables[name].weight.copy_(oracle.tables[name.replace("/", "__")].weight)
        lin = inputs.projection_module[0][0]
        lin.weight.copy_(oracle.proj.weight)
        lin.bias.copy_(oracle.proj.bias)
        inputs.masking.masked_item_embedding.copy_(oracle.masked_item_embedding)
        missing, unexpected = tblock.transformer.load_state_dict(oracle.transformer.state_dict(), strict=False)
        assert not missing, missing
        task = head.prediction_task_dict["next-item"]
        if oracle.task_block is not None:
            tl = task.task_block[0][0]
            tl.weight.copy_(oracle.task_block.weight)
            tl.bias.copy_(oracle.task_block.bias)
