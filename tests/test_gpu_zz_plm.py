"""GPU tests of XLNet Permutation Language Modeling: the mask kernel, the two-stream (MASKED)
instantiations of the tensor-path attention and the stacked h/g encoder forward.  CPU-side evidence: mask code bit-exact on its host twin
against the upstream vectors, attention index algebra emulated (tools/emu_attn_mma.py), host flow against HF's
two-stream forward with kernel doubles; validated on a B200 in round 2."""
import os

import pytest
import torch

import t4r_oracle as O
from _util import make_pair, synth_batch

pytestmark = [pytest.mark.gpu]
TOL = 1e-3


def _draws(B, L, seed):
    g = torch.Generator().manual_seed(seed)
    return {"u_span": torch.rand((B, L), generator=g), "u_start": torch.rand((B, L), generator=g),
            "u_force": torch.rand((B,), generator=g), "u_unmask": torch.rand((B,), generator=g),
            "perm": torch.stack([torch.randperm(L, generator=g) for _ in range(B)])}


@pytest.mark.parametrize("span,prob", [(5, 1 / 6), (3, 0.5)])
def test_mask_kernel_matches_host_twin(span, prob):
    from transformers4rec_b200 import _lib, ops
    B, L = 1000, 20
    ids = synth_batch(B, L, {"i": 5000}, seed=3, min_len=1)["i"]
    ids[7] = 0
    d = _draws(B, L, 4)
    for mode in (_lib.PLM_TRAIN, _lib.PLM_EVAL_LAST, _lib.PLM_EVAL_ALL):
        h = ops.mask_plm_host(ids, mode, 0, span, prob, d)
        g = ops.mask_plm(ids.cuda(), mode, 0, span, prob, {k: v.cuda() for k, v in d.items()})
        for a, b in zip(h, g):
            assert torch.equal(a, b.cpu()), mode


@pytest.mark.parametrize("d,H,NL,B,L", [(64, 4, 2, 9, 20), (256, 8, 1, 5, 30), (128, 2, 1, 3, 7), (64, 4, 1, 33, 12)])
def test_two_stream_encoder_matches_hf(d, H, NL, B, L):
    import transformers4rec_b200.torch as tr
    torch.manual_seed(10)
    hf = O.build_hf_xlnet(d, H, NL).eval()
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if "layer_norm" in n:
                p.add_(torch.randn_like(p) * 0.1)
            else:
                p.normal_(0.0, 0.08)
    blk = tr.TransformerBlock(hf).cuda()
    x = torch.randn(B, L, d)
    ids = synth_batch(B, L, {"i": 500}, seed=1)["i"]
    _, _, tm, pm, _ = O.plm_compute_masked_targets(ids, True, draws=_draws(B, L, 2))
    pm[0, 3, :] = 1.0   # a query that sees nothing: uniform attention, HF's -1e30 rule
    with torch.no_grad():
        ref = O.hf_encoder_forward_plm(hf, x, pm, tm)
        got = blk.transformer(inputs_embeds=x.cuda(), perm_mask=pm.to(torch.uint8).cuda())[0].cpu()
    assert (got - ref).abs().max().item() < TOL


def test_model_plm_training_and_eval_loss():
    oracle, model = make_pair({"item_id/list": 3001}, {"item_id/list": 64}, "item_id/list", (), 64, 4, 2, 20,
                              masking="plm", weight_scale=0.08)
    with torch.no_grad():
        model.heads[0].body[1].transformer.mask_emb.normal_(0.0, 0.5)
        oracle.transformer.mask_emb.copy_(model.heads[0].body[1].transformer.mask_emb.cpu())
    B, L = 48, 20
    batch = synth_batch(B, L, {"item_id/list": 3001}, seed=5)
    d = _draws(B, L, 6)
    inputs = model.heads[0].body[0]
    inputs.masking.set_draws({k: v.cuda() for k, v in d.items()})
    dev = {k: v.cuda() for k, v in batch.items()}
    with torch.no_grad():
        ref = oracle(batch, training=True, draws=d)
        out = model(dev, training=True)
        ref_e = oracle(batch, training=False, testing=True)
        out_e = model(dev, training=False, testing=True)
    assert torch.equal(inputs.masking.masked_targets.cpu(), ref_e["masked_targets"])
    assert abs(out["loss"].item() - ref["loss"].item()) < TOL
    assert abs(out_e["loss"].item() - ref_e["loss"].item()) < TOL
