"""GPU tests at BASELINE.json's full sizes, through size-independent properties:
sessions are independent (a slice of a big batch must match the CPU oracle run on
that slice alone), the fused head must agree with materialised fp32 logits computed
by the plain SIMT reference GEMM, ranks must agree with a direct count on the
materialised logits, and probabilities must sum to one."""
import math

import pytest
import torch

import t4r_oracle as O
from _util import mlm_draws, synth_batch

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _build(cards, dims, item, cont, d, H, NL, L, arch, masking, seed=1):
    import transformers4rec_b200.torch as tr
    torch.manual_seed(seed)
    cols = [tr.ColumnSchema.create_categorical(n, c - 1, tags=[tr.Tags.ITEM_ID] if n == item else None)
            for n, c in cards.items()]
    cols += [tr.ColumnSchema.create_continuous(n) for n in cont]
    inputs = tr.TabularSequenceFeatures.from_schema(tr.Schema(cols), max_sequence_length=L, d_output=d,
                                                    masking=masking, embedding_dims=dims)
    cfg = (tr.XLNetConfig if arch == "xlnet" else tr.GPT2Config).build(d_model=d, n_head=H, n_layer=NL,
                                                                       total_seq_length=L)
    return cfg.to_torch_model(inputs, tr.NextItemPredictionTask(weight_tying=True))


def test_config2_full_size_head_and_hidden():
    """configs[1]: 1M items, L=20, XLNet d=256 x 4 layers, MLM, B=2048."""
    from transformers4rec_b200 import ops
    cards, dims = {"item_id/list": 1_000_001}, {"item_id/list": 256}
    B, L, d = 2048, 20, 256
    model = _build(cards, dims, "item_id/list", (), d, 8, 4, L, "xlnet", "mlm").cuda().eval()
    with torch.no_grad():  # "trained-like" encoder weights
        for n, p in model.heads[0].body[1].transformer.named_parameters():
            if p.ndim >= 2 and "layer_norm" not in n:
                p.normal_(0.0, 0.05)
    batch = synth_batch(B, L, cards, seed=0)
    u, draws = mlm_draws(B, L)
    body = model.heads[0].body
    body[0].masking.set_draws(u.cuda())
    dev = {k: v.cuda() for k, v in batch.items()}
    with torch.no_grad():
        out = model(dev, training=True)
        hid = body(dev, training=True)
    assert math.isfinite(out["loss"].item())
    # (1) session independence: the first 24 sessions alone, on the CPU oracle
    nb = 24
    table = body[0].item_embedding_table.weight.detach()
    small_ids = batch["item_id/list"][:nb]
    lin = body[0].projection_module[0][0]
    with torch.no_grad():
        x = torch.nn.functional.embedding(small_ids, table.cpu())
        x = O.project_relu(x, lin.weight.cpu(), lin.bias.cpu())
        mask, labels = O.mlm_compute_masked_targets(small_ids, True, False, u_bern=draws["u_bern"][:nb],
                                                    u_force=draws["u_force"][:nb], u_unmask=draws["u_unmask"][:nb])
        x = O.mlm_apply_mask_to_inputs(x, mask, body[0].masking.masked_item_embedding.detach().cpu(), True, False)
        hf = O.build_hf_xlnet(d, 8, 4).eval()
        hf.load_state_dict({k: v.cpu() for k, v in body[1].transformer.state_dict().items()}, strict=False)
        ref_hid = O.hf_encoder_forward(hf, x)
    assert torch.equal(body[0].masking.masked_targets[:nb].cpu(), labels)
    assert (hid[:nb].cpu() - ref_hid).abs().max().item() < TOL
    # (2) fused head vs materialised fp32 logits from the SIMT reference GEMM (1M classes, 256 rows)
    task = model.heads[0].prediction_task_dict["next-item"]
    T = int(task._last["count"].item())
    assert T == out["labels"].numel() and T > B
    flat_rows = body[0].masking.masked_targets.flatten().nonzero().squeeze(1)
    xt = hid.reshape(-1, d)[flat_rows[:256]]
    logits_ref = ops.debug_sgemm_nt(xt, table)
    lse_ref = torch.logsumexp(logits_ref.double(), dim=1).float()
    assert (task._last["row_lse"][:256] - lse_ref).abs().max().item() < 2e-4
    y = task._last["labels"][:256]
    row_loss_ref = lse_ref - logits_ref.gather(1, y.unsqueeze(1)).squeeze(1)
    assert (task._last["row_loss"][:256] - row_loss_ref).abs().max().item() < 2e-4
    # (3) probabilities sum to one
    p_sum = torch.exp(logits_ref - task._last["row_lse"][:256].unsqueeze(1)).sum(1)
    assert (p_sum - 1).abs().max().item() < 1e-3
    # (4) evaluation at full size: ranks vs a direct count on materialised logits for a slice
    with torch.no_grad():
        out_e = model(dev, training=False, testing=True)
        hs = body(dev, training=False, testing=True).reshape(-1, d)
    ranks = out_e.row_rank[:128].cpu().long()
    rows_e = body[0].masking.masked_targets.flatten().nonzero().squeeze(1)[:128]
    lg = ops.debug_sgemm_nt(hs[rows_e], table)
    ye = task._last["labels"][:128]
    rank_ref = (lg > lg.gather(1, ye.unsqueeze(1))).sum(1).cpu()
    # with 1M classes there are ~2e5 logits per unit of score around a random label, so the ~1e-5
    # difference between the two summation orders moves a rank of ~5e5 by a few places at most
    assert (ranks - rank_ref).abs().max().item() <= 8
    assert ((ranks - rank_ref).abs().float() / rank_ref.clamp(min=1).float()).max().item() < 1e-3


def test_config3_full_size_gpt2_side_features():
    """configs[2]: 1M items + 6 categorical side features, concat, GPT-2 CLM, B=4096,
    task_block Linear(256 -> 64) in front of the tied 64-d table."""
    cards = {"item_id/list": 1_000_001, "category/list": 337, "brand/list": 1000, "shop/list": 10000,
             "price_bin/list": 100, "weekday/list": 32, "hour_bin/list": 7}
    dims = {k: 64 for k in cards}
    B, L, d = 4096, 20, 256
    model = _build(cards, dims, "item_id/list", (), d, 8, 4, L, "gpt2", "clm").cuda().eval()
    task = model.heads[0].prediction_task_dict["next-item"]
    assert task.task_block is not None and model.heads[0].body[0]._layout()[1] == 448
    batch = synth_batch(B, L, cards, seed=3)
    dev = {k: v.cuda() for k, v in batch.items()}
    with torch.no_grad():
        out = model(dev, training=True)
    loss = out["loss"].item()
    T = int(task._last["count"].item())
    assert T == int((batch["item_id/list"][:, 1:] != 0).sum())  # CLM: one label per next item
    assert abs(loss - math.log(1_000_001)) < 0.5
    # session independence against the CPU oracle on the first 16 sessions
    nb = 16
    body = model.heads[0].body
    lin = body[0].projection_module[0][0]
    with torch.no_grad():
        hid = body(dev, training=True)[:nb].cpu()
        tables = {n: body[0].categorical_module.embedding_tables[n].weight.detach().cpu() for n in cards}
        small = {n: batch[n][:nb] for n in cards}
        x = O.embed_concat(tables, small)
        x = O.project_relu(x, lin.weight.cpu(), lin.bias.cpu())
        mask, labels = O.clm_compute_masked_targets(small["item_id/list"], True, False)
        x = O.clm_apply_mask_to_inputs(x, mask, body[0].masking.masked_item_embedding.detach().cpu(), True, False)
        hf = O.build_hf_gpt2(d, 8, 4, L).eval()
        hf.load_state_dict({k: v.cpu() for k, v in body[1].transformer.state_dict().items()}, strict=False)
        ref_hid = O.hf_encoder_forward(hf, x)
    assert (hid - ref_hid).abs().max().item() < TOL


def test_config5_shape_sampled_softmax_L50():
    """configs[4] at one rank's shard size: 6.25M-row table (50M / 8), L=50, sampled softmax
    with 50K negatives, XLNet d=256; row losses must match a direct evaluation of the
    sampled logits for a slice of the label rows."""
    import transformers4rec_b200.torch as tr
    V = 6_250_001
    dims = {"item_id/list": 256}
    B, L, d, S = 512, 50, 256, 50_000
    torch.manual_seed(5)
    inputs = tr.TabularSequenceFeatures.from_schema(
        tr.Schema([tr.ColumnSchema.create_categorical("item_id/list", V - 1, tags=[tr.Tags.ITEM_ID])]),
        max_sequence_length=L, d_output=d, masking="mlm", embedding_dims=dims)
    cfg = tr.XLNetConfig.build(d_model=d, n_head=8, n_layer=2, total_seq_length=L)
    task = tr.NextItemPredictionTask(weight_tying=True, sampled_softmax=True, max_n_samples=S)
    model = cfg.to_torch_model(inputs, task).cuda().eval()
    batch = synth_batch(B, L, {"item_id/list": V}, seed=4)
    dev = {k: v.cuda() for k, v in batch.items()}
    with torch.no_grad():
        out = model(dev, training=True)
    st = task._last
    T = int(st["count"].item())
    neg = st["neg"]
    assert 0.7 * S <= neg.numel() <= S and torch.equal(neg, neg.unique())  # sorted, unique, truncated
    table = inputs.item_embedding_table.weight.detach()
    n = 64
    xt = st["xt_planes"][0, :n].float() + st["xt_planes"][1, :n].float()
    y = st["labels"][:n]
    nlq = task.sampler.neg_log_q
    pos = (xt * table[y]).sum(1) + nlq[y]
    negs = xt @ table[neg].t() + nlq[neg].unsqueeze(0)
    negs[y.unsqueeze(1) == neg.unsqueeze(0)] = torch.finfo(torch.float16).min / 100.0
    ref_rows = torch.nn.functional.cross_entropy(torch.cat([pos.unsqueeze(1), negs], 1),
                                                 torch.zeros(n, dtype=torch.long, device="cuda"), reduction="none")
    assert (st["row_loss"][:n] - ref_rows).abs().max().item() < 1e-3
    assert math.isfinite(out["loss"].item()) and T > B
