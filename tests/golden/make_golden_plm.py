#!/usr/bin/env python
"""Golden vectors for Permutation Language Modeling, produced by executing the UPSTREAM
``transformers4rec/torch/masking.py`` (PermutationLanguageModeling, :501-740) in the authoring container -- same
stub harness as make_golden.py.

Randomness: the upstream code calls ``torch.randint`` (twice per iteration of a data-dependent ``while`` loop, per
session), ``torch.multinomial`` (1-D per session when nothing got masked, 2-D once per batch) and ``torch.randperm``
(once per session).  While it runs they are patched to consume explicit draws in call order.  The draws are defined
PER SESSION (``u_span`` / ``u_start`` [B, NMAX], ``u_force`` / ``u_unmask`` [B], ``perm`` [B, L]: the layout the oracle
and the CUDA kernel use); the sequential queues the upstream code pops from are laid out from them with the iteration
counts of the oracle's own pass -- if the restatement walked the loops differently the upstream run would pop
misaligned draws and the comparison in tests/test_oracle_golden.py would fail.

Usage:  python tests/golden/make_golden_plm.py     (writes tests/golden/reference_vectors_plm.pt)
"""
import importlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (stub harness)

O = MG.O


class patched_plm_draws:
    def __init__(self, randint_us, force_us, unmask_u, perms):
        self.randint_us, self.force_us, self.unmask_u, self.perms = list(randint_us), list(force_us), unmask_u, list(perms)

    def __enter__(self):
        self._ri, self._m, self._rp = torch.randint, torch.multinomial, torch.randperm

        def randint(*args, **k):
            # torch.randint(low, high, size) or torch.randint(high, size)
            if isinstance(args[1], int):
                low, high = args[0], args[1]
            else:
                low, high = 0, args[0]
            return torch.tensor([low + O.randint_from_uniform(self.randint_us.pop(0), high - low)])

        def multinomial(weights, num_samples, replacement=False, **k):
            if weights.ndim == 1:
                return O.pick_kth_set((weights > 0).unsqueeze(0), self.force_us.pop(0).reshape(1))
            return O.pick_kth_set(weights > 0, self.unmask_u).unsqueeze(1)

        def randperm(n, **k):
            return self.perms.pop(0)

        torch.randint, torch.multinomial, torch.randperm = randint, multinomial, randperm
        return self

    def __exit__(self, *a):
        torch.randint, torch.multinomial, torch.randperm = self._ri, self._m, self._rp
        assert not self.randint_us and not self.force_us and not self.perms, "the upstream code consumed fewer draws"


def _pack(m, y):
    """0/1 masks are stored as uint8 (+ their upstream dtype) to keep the fixture small."""
    tm, pm = m.target_mapping, m.perm_mask
    assert bool(((tm == 0) | (tm == 1)).all()) and bool(((pm == 0) | (pm == 1)).all())
    return {"mask_schema": m.mask_schema.clone(), "masked_targets": m.masked_targets.clone(),
            "target_mapping": tm.to(torch.uint8), "target_mapping_dtype": str(tm.dtype),
            "perm_mask": pm.to(torch.uint8), "perm_mask_dtype": str(pm.dtype), "out": y.clone()}


def main():
    assert os.path.isdir(MG.REF)
    MG.install_stubs()
    fake = MG._Anything("transformers4rec.torch.masking")
    sys.modules["transformers4rec.torch.masking"] = fake
    importlib.import_module("transformers4rec.torch.utils.torch_utils")
    del sys.modules["transformers4rec.torch.masking"]
    masking = importlib.import_module("transformers4rec.torch.masking")

    g = torch.Generator().manual_seed(4321)
    B, L, d = 96, 14, 8
    lens = torch.randint(1, L + 1, (B,), generator=g)
    lens[:6] = torch.tensor([1, 1, 2, 2, L, L])          # edge cases: single item, full length
    ids = torch.randint(1, 500, (B, L), generator=g)
    ids = torch.where(torch.arange(L).unsqueeze(0) < lens.unsqueeze(1), ids, torch.zeros_like(ids))
    x = torch.rand((B, L, d), generator=g)
    out = {"item_ids": ids, "x": x, "cases": {}}
    for name, kw in (("default", {}), ("p0.5_span3", {"plm_probability": 0.5, "max_span_length": 3}),
                     ("evalall", {"eval_on_last_item_seq_only": False})):
        draws = {"u_span": torch.rand((B, L), generator=g), "u_start": torch.rand((B, L), generator=g),
                 "u_force": torch.rand((B,), generator=g), "u_unmask": torch.rand((B,), generator=g),
                 "perm": torch.stack([torch.randperm(L, generator=g) for _ in range(B)])}
        okw = {k: v for k, v in kw.items()}
        _, _, _, _, info = O.plm_compute_masked_targets(ids, True, draws=draws, **okw)
        randint_us, force_us = [], []
        for b in range(B):
            for n in range(info["n_iter"][b]):
                randint_us += [draws["u_span"][b, n], draws["u_start"][b, n]]
            if info["forced"][b]:
                force_us.append(draws["u_force"][b])
        torch.manual_seed(9)
        m = masking.PermutationLanguageModeling(hidden_size=d, **kw)
        emb = m.masked_item_embedding.detach().clone()
        case = {"kwargs": kw, "draws": draws, "masked_item_embedding": emb}
        with patched_plm_draws(randint_us, force_us, draws["u_unmask"], list(draws["perm"])):
            with torch.no_grad():
                y = m(x, item_ids=ids, training=True, testing=False)
        case["train"] = _pack(m, y)
        for tag, (tr, te) in (("eval", (False, True)), ("infer", (False, False))):
            with torch.no_grad():
                y = m(x, item_ids=ids, training=tr, testing=te)
            case[tag] = _pack(m, y)
        out["cases"][name] = case
        print(name, "sessions forced:", sum(info["forced"]), "iterations:", sum(info["n_iter"]),
              "labels:", int(case["train"]["mask_schema"].sum()))
    torch.save(out, os.path.join(HERE, "reference_vectors_plm.pt"))
    print("wrote reference_vectors_plm.pt")


if __name__ == "__main__":
    main()
