"""GPU tests of the 2-unit product head (``nprod=2``: fp16 x fp16 + two e4m3 cross terms,
csrc/t4r_mixed_pack.cuh).

The operand packing is pinned on the CPU (tests/test_abi_and_host.py, host twin of the kernel); the tcgen05 side
(single-CTA, CTA-pair and resident-A kernels) was validated on a B200 in round 2.
"""
import math
import os

import pytest
import torch

import t4r_oracle as O

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def ops():
    from transformers4rec_b200 import ops as _ops
    return _ops


def test_device_packing_matches_host_twin_bit_exactly(ops):
    g = torch.Generator().manual_seed(21)
    x = torch.randn(300, 200, generator=g)
    x[3] = 0.0
    x[5] *= 1e20
    x[6, :] = torch.exp2(torch.randint(-20, 4, (200,), generator=g).float())
    ph, ih = ops.split_planes_mixed_host(x)
    pd, idv = ops.split_planes_mixed(x.cuda())
    assert torch.equal(idv.cpu(), ih)
    assert torch.equal(pd.cpu(), ph)


@pytest.mark.parametrize("two_cta", ["0", "1"])
@pytest.mark.parametrize("T,V,De", [(128, 256, 64), (300, 1000, 64), (257, 777, 200), (513, 2049, 256)])
def test_materialised_logits_nprod2(ops, monkeypatch, T, V, De, two_cta):
    """Element-wise check of the 2-unit product through the dense epilogue: the first thing to look at if the
    fused head disagrees (one K block, one tile in the smallest case).  Bar: the emulated product itself
    (tests/_mixed_ref.py) to accumulation-order noise, and fp64 to the scheme's error budget."""
    import _mixed_ref as R
    monkeypatch.setenv("T4R_GEMM_2CTA", two_cta)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(T, De, generator=g)
    w = torch.randn(V, De, generator=g) * 0.1
    xp, xi = ops.split_planes_mixed(x.cuda())
    wp, wi = ops.split_planes_mixed(w.cuda())
    got = ops.head_logits_mixed(xp, xi, wp, wi, De, inv_temperature=0.5).cpu().double()
    emu = R.product(R.pack(x), R.pack(w)) * 0.5
    ref = (x.double() @ w.double().t()) * 0.5
    scale = ref.abs().max().item()
    assert (got - emu).abs().max().item() < 2e-6 * max(scale, 1.0)
    assert (got - ref).abs().max().item() < 1e-4 * max(scale, 1.0)


@pytest.mark.parametrize("two_cta", ["0", "1"])
@pytest.mark.parametrize("drop,terms", [(128 + 256, ("main",)), (64 + 256, ("cross1",)), (64 + 128, ("cross2",))])
def test_each_partial_product_alone(ops, monkeypatch, drop, terms, two_cta):
    """Bring-up aid: T4R_GEMM_DEBUG bits 64 / 128 / 256 drop the fp16 main product / lo8(A) x hi8(B) / hi8(A) x lo8(B);
    each remaining term alone must equal its emulation (exact products, fp32 vs fp64 accumulation).  If the full
    product is off, this says which instruction descriptor / operand offset is to blame."""
    import _mixed_ref as R
    monkeypatch.setenv("T4R_GEMM_2CTA", two_cta)
    monkeypatch.setenv("T4R_GEMM_DEBUG", str(drop))
    g = torch.Generator().manual_seed(8)
    T, V, De = 257, 777, 200
    x = torch.randn(T, De, generator=g)
    w = torch.randn(V, De, generator=g) * 0.1
    xp, xi = ops.split_planes_mixed(x.cuda())
    wp, wi = ops.split_planes_mixed(w.cuda())
    got = ops.head_logits_mixed(xp, xi, wp, wi, De).cpu().double()
    emu = R.product(R.pack(x), R.pack(w), terms)
    assert (got - emu).abs().max().item() < 2e-6 * max(emu.abs().max().item(), 1e-3)


@pytest.mark.parametrize("kernel", ["single", "pair", "resident"])
@pytest.mark.parametrize("T,V,De,tau", [(200, 10001, 64, 1.0), (517, 30011, 256, 1.0), (64, 999, 128, 0.5),
                                        (600, 123001, 256, 1.0)])
def test_head_full_softmax_nprod2(ops, monkeypatch, T, V, De, tau, kernel):
    """Same cases and bars as test_gpu_parity.py::test_head_full_softmax, on all three head kernels (single CTA,
    CTA pair, CTA pair with the resident A tile); the last case has more (row block x column chunk) units than
    CTA pairs, so the resident kernel's unit-to-unit hand-over of the A slots is exercised."""
    monkeypatch.setenv("T4R_GEMM_2CTA", "0" if kernel == "single" else "1")
    monkeypatch.setenv("T4R_HEAD_RESIDENT", "1" if kernel == "resident" else "0")
    torch.manual_seed(12)
    xt = torch.randn(T, De)
    W = torch.randn(V, De) * 0.1
    y = torch.randint(1, V, (T,))
    ref_loss, ref_logits = O.full_softmax_head(xt, y, W, tau)
    cap = T + 37
    xt_pad = torch.zeros(cap, De); xt_pad[:T] = xt
    y_pad = torch.zeros(cap, dtype=torch.long); y_pad[:T] = y
    count = torch.tensor([T], dtype=torch.int32, device="cuda")
    xp, xi = ops.split_planes_mixed(xt_pad.cuda())
    wp, wi = ops.split_planes_mixed(W.cuda())
    res = ops.head_softmax_ce(xp, xt_pad.cuda(), y_pad.cuda(), wp, W.cuda(), t_dev=count, inv_temperature=1.0 / tau,
                              want_rank=True, nprod=2, xt_inv_scale=xi, w_inv_scale=wi)
    assert abs(res["loss"].item() - ref_loss.item()) < 1e-4
    ref_lse = torch.logsumexp(ref_logits, dim=1)
    assert (res["row_lse"][:T].cpu() - ref_lse).abs().max().item() < 2e-4
    ks = [1, 5, 10, 20]
    ref_rec = O.recall_at_mean(ks, ref_logits, y)
    got_rec = ops.recall_from_ranks(res["row_rank"], ks, count).cpu()
    assert (got_rec - ref_rec).abs().max().item() < 1e-6


@pytest.mark.parametrize("nprod", [3, 1])
def test_resident_head_kernel_matches_the_default_kernel(ops, monkeypatch, nprod):
    """T4R_HEAD_RESIDENT=1 with the shipped arithmetic: same operands, same products, only the operand staging
    differs -> row_lse / loss / ranks must agree with the default CTA-pair kernel to accumulation-order noise."""
    torch.manual_seed(5)
    T, V, De = 700, 140001, 256
    xt = torch.randn(T, De, device="cuda")
    W = torch.randn(V, De, device="cuda") * 0.1
    y = torch.randint(1, V, (T,), device="cuda")
    xp, wp = ops.split_planes(xt), ops.split_planes(W)
    monkeypatch.setenv("T4R_HEAD_RESIDENT", "0")
    a = ops.head_softmax_ce(xp, xt, y, wp, W, want_rank=True, nprod=nprod)
    monkeypatch.setenv("T4R_HEAD_RESIDENT", "1")
    b = ops.head_softmax_ce(xp, xt, y, wp, W, want_rank=True, nprod=nprod)    # general epilogue (ranks)
    c = ops.head_softmax_ce(xp, xt, y, wp, W, want_rank=False, nprod=nprod)   # packed-pair fast path
    for r in (b, c):
        assert (a["row_lse"] - r["row_lse"]).abs().max().item() < 1e-5
        assert abs(a["loss"].item() - r["loss"].item()) < 4e-6   # a few ulp of a loss around 13
    assert torch.equal(a["row_rank"], b["row_rank"])


def test_nprod2_requires_its_scales(ops):
    from transformers4rec_b200 import T4RError
    xt = torch.randn(64, 64, device="cuda")
    W = torch.randn(500, 64, device="cuda")
    xp, xi = ops.split_planes_mixed(xt)
    wp, wi = ops.split_planes_mixed(W)
    y = torch.randint(1, 500, (64,), device="cuda")
    with pytest.raises(T4RError):
        ops.head_softmax_ce(xp, xt, y, wp, W, nprod=2)


def test_model_training_loss_with_mixed_head():
    """Whole model, config-1 shape: task.nprod = 2 must reproduce the oracle's training loss within the parity bar
    and agree with the default arithmetic to 1e-4."""
    from _util import make_pair, mlm_draws, synth_batch
    cards = {"item_id/list": 10001, "category/list": 337}
    dims = {"item_id/list": 64, "category/list": 64}
    cont = tuple(f"cont{i}/list" for i in range(5))
    B, L = 256, 20
    oracle, model = make_pair(cards, dims, "item_id/list", cont, 64, 4, 2, L, weight_scale=0.08)
    batch = synth_batch(B, L, cards, cont, seed=0)
    u, draws = mlm_draws(B, L)
    model.heads[0].body[0].masking.set_draws(u.cuda())
    dev = {k: v.cuda() for k, v in batch.items()}
    task = model.heads[0].prediction_task_dict["next-item"]
    with torch.no_grad():
        ref = oracle(batch, training=True, draws=draws)["loss"].item()
        l3 = model(dev, training=True)["loss"].item()
        task.nprod = 2
        out = model(dev, training=True)
        l2 = out["loss"].item()
        preds = out["predictions"]  # lazily materialised through the 3-product planes
    assert math.isfinite(l2) and abs(l2 - ref) < 1e-3 and abs(l2 - l3) < 1e-4
    assert preds.shape[1] == 10001


def test_head_variants_at_config2_full_size(ops, monkeypatch):
    """BASELINE configs[1] head shape (T = 5120 label rows, V = 1 000 001, De = 256), size-independent properties:
    every opt-in variant (resident-A kernel with the shipped 3-product arithmetic; 2-unit product on the CTA-pair and
    on the resident-A kernel) reproduces the default kernel's per-row log-sum-exp and loss (same logits up to the
    product's error budget), probabilities sum to one, and permuting the table rows -- columns of the GEMM, so a
    different tile / chunk assignment for every class -- leaves every row's log-sum-exp where it was."""
    torch.manual_seed(3)
    T, V, De = 5120, 1_000_001, 256
    xt = torch.nn.functional.layer_norm(torch.randn(T, De, device="cuda"), (De,))
    W = torch.randn(V, De, device="cuda") * 0.05
    y = torch.randint(1, V, (T,), device="cuda")
    xp, wp = ops.split_planes(xt), ops.split_planes(W)
    monkeypatch.setenv("T4R_HEAD_RESIDENT", "0")
    base = ops.head_softmax_ce(xp, xt, y, wp, W)
    perm = torch.randperm(V, device="cuda")
    inv = torch.empty_like(perm); inv[perm] = torch.arange(V, device="cuda")
    Wq = W[perm].contiguous()
    xm, xi = ops.split_planes_mixed(xt)
    wm, wi = ops.split_planes_mixed(W)
    wqm, wqi = ops.split_planes_mixed(Wq)
    runs = {}
    monkeypatch.setenv("T4R_HEAD_RESIDENT", "1")
    runs["resident x3"] = (ops.head_softmax_ce(xp, xt, y, wp, W), 1e-5)
    runs["resident x3, permuted table"] = (ops.head_softmax_ce(xp, xt, inv[y], ops.split_planes(Wq), Wq), 2e-5)
    runs["resident 2-unit"] = (ops.head_softmax_ce(xm, xt, y, wm, W, nprod=2, xt_inv_scale=xi, w_inv_scale=wi), 2e-4)
    runs["resident 2-unit, permuted table"] = (ops.head_softmax_ce(xm, xt, inv[y], wqm, Wq, nprod=2, xt_inv_scale=xi,
                                                                   w_inv_scale=wqi), 2e-4)
    monkeypatch.setenv("T4R_HEAD_RESIDENT", "0")
    runs["pair 2-unit"] = (ops.head_softmax_ce(xm, xt, y, wm, W, nprod=2, xt_inv_scale=xi, w_inv_scale=wi), 2e-4)
    for name, (r, tol) in runs.items():
        assert (r["row_lse"] - base["row_lse"]).abs().max().item() < tol, name
        assert (r["row_loss"] - base["row_loss"]).abs().max().item() < 2 * tol, name
        assert abs(r["loss"].item() - base["loss"].item()) < tol, name
    # probabilities sum to one: materialised 2-unit logits of a slice of rows against the fused log-sum-exp
    lg = ops.head_logits_mixed(xm[:, :128].contiguous(), xi[:128].contiguous(), wm, wi, De)
    p_sum = torch.exp(lg - runs["resident 2-unit"][0]["row_lse"][:128].unsqueeze(1)).sum(1)
    assert (p_sum - 1).abs().max().item() < 1e-3
