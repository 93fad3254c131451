"""Host-side mirror of ``transformers4rec/torch/masking.py`` for MLM, CLM and PLM.

Same class names, constructor arguments, attributes (``mask_schema``,
``masked_targets``, ``padding_idx``, ``masked_item_embedding``,
``transformer_arguments``) and registry strings ("mlm"/"masked",
"clm"/"causal").  Label/mask generation runs in ``t4r_mask_mlm`` /
``t4r_mask_clm`` (one integer kernel instead of ~15 ATen launches); replacing
masked rows is fused into the projection epilogue through ``row_code``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Optional

import torch
from torch import nn

from . import _lib, ops


@dataclass
class MaskingInfo:
    schema: torch.Tensor
    targets: torch.Tensor


class _Registry(dict):
    def register_with_multiple_names(self, *names):
        def deco(cls):
            for n in names:
                self[n] = cls
            return cls
        return deco

    def parse(self, name):
        if isinstance(name, str):
            if name not in self:
                raise ValueError(f"{name} is not a registered masking; available: {sorted(self)}")
            return self[name]
        return name


masking_registry = _Registry()


class MaskSequence(nn.Module):
    """masking.py:61-242."""

    def __init__(self, hidden_size: int, padding_idx: int = 0, eval_on_last_item_seq_only: bool = True, **kwargs):
        super().__init__()
        self.padding_idx = padding_idx
        self.hidden_size = hidden_size
        self.eval_on_last_item_seq_only = eval_on_last_item_seq_only
        self.mask_schema: Optional[torch.Tensor] = None
        self.masked_targets: Optional[torch.Tensor] = None
        self.row_code: Optional[torch.Tensor] = None  # uint8, consumed by the fused epilogue
        self.masked_item_embedding = nn.Parameter(torch.Tensor(self.hidden_size))
        torch.nn.init.normal_(self.masked_item_embedding, mean=0, std=0.001)  # masking.py:103-108
        self._draws: Optional[torch.Tensor] = None

    # -- test hook: feed explicit uniform draws u[B, L+2] (see DESIGN.md "Random draws")
    def set_draws(self, u: Optional[torch.Tensor]):
        self._draws = u

    def _compute_masked_targets(self, item_ids, training=False, testing=False) -> MaskingInfo:
        raise NotImplementedError

    def compute_masked_targets(self, item_ids, training=False, testing=False) -> MaskingInfo:
        assert item_ids.ndim == 2, "`item_ids` must have 2 dimensions."
        info = self._compute_masked_targets(item_ids, training=training, testing=testing)
        self.mask_schema, self.masked_targets = info.schema, info.targets
        return info

    def predict_all(self, item_ids) -> MaskingInfo:
        mask, labels, _ = ops.mask_mlm(item_ids, _lib.MLM_EVAL_ALL, self.padding_idx)
        return MaskingInfo(mask, labels)

    def apply_mask_to_inputs(self, inputs, schema, training=False, testing=False):
        """Stand-alone application (when nothing upstream fused it): one kernel that
        rewrites rows according to ``row_code``."""
        raise NotImplementedError

    def forward(self, inputs, item_ids, training=False, testing=False):
        self.compute_masked_targets(item_ids=item_ids, training=training, testing=testing)
        if self.mask_schema is None:
            raise ValueError("`mask_schema must be set.`")
        return self.apply_mask_to_inputs(inputs, self.mask_schema, training=training, testing=testing)

    def forward_output_size(self, input_size):
        return input_size

    def transformer_required_arguments(self) -> Dict[str, Any]:
        return {}

    def transformer_optional_arguments(self) -> Dict[str, Any]:
        return {}

    @property
    def transformer_arguments(self) -> Dict[str, Any]:
        return {**self.transformer_required_arguments(), **self.transformer_optional_arguments()}

    def _apply_codes(self, inputs: torch.Tensor, code: torch.Tensor) -> torch.Tensor:
        B, L, d = inputs.shape
        planes, of = ops.split_planes(inputs.reshape(B * L, d), row_code=code.reshape(-1),
                                      mask_vec=self.masked_item_embedding.detach().float(), want_f32=True)
        out = of.view(B, L, d)
        out._t4r_planes = planes
        return out


@masking_registry.register_with_multiple_names("clm", "causal")
class CausalLanguageModeling(MaskSequence):
    """masking.py:245-337."""

    def __init__(self, hidden_size, padding_idx=0, eval_on_last_item_seq_only=True,
                 train_on_last_item_seq_only=False, **kwargs):
        super().__init__(hidden_size=hidden_size, padding_idx=padding_idx,
                         eval_on_last_item_seq_only=eval_on_last_item_seq_only)
        self.train_on_last_item_seq_only = train_on_last_item_seq_only

    def _mode(self, training, testing):
        if not training and not testing:
            return _lib.CLM_INFERENCE
        if (self.eval_on_last_item_seq_only and not training) or (self.train_on_last_item_seq_only and training):
            return _lib.CLM_LAST
        return _lib.CLM_ALL

    def _compute_masked_targets(self, item_ids, training=False, testing=False) -> MaskingInfo:
        mask, labels, code = ops.mask_clm(item_ids, self._mode(training, testing), self.padding_idx)
        self.row_code = code
        return MaskingInfo(mask, labels)

    def apply_mask_to_inputs(self, inputs, mask_schema, training=False, testing=False):
        return self._apply_codes(inputs, self.row_code)


@masking_registry.register_with_multiple_names("mlm", "masked")
class MaskedLanguageModeling(MaskSequence):
    """masking.py:340-498."""

    def __init__(self, hidden_size, padding_idx=0, eval_on_last_item_seq_only=True, mlm_probability=0.15, **kwargs):
        super().__init__(hidden_size=hidden_size, padding_idx=padding_idx,
                         eval_on_last_item_seq_only=eval_on_last_item_seq_only)
        self.mlm_probability = mlm_probability

    def _mode(self, training, testing):
        if not training and not testing:
            return _lib.MLM_INFERENCE
        if training:
            return _lib.MLM_TRAIN
        return _lib.MLM_EVAL_LAST if self.eval_on_last_item_seq_only else _lib.MLM_EVAL_ALL

    def _compute_masked_targets(self, item_ids, training=False, testing=False) -> MaskingInfo:
        mode = self._mode(training, testing)
        mask, labels, code = ops.mask_mlm(item_ids, mode, self.padding_idx, self.mlm_probability,
                                          u=self._draws if mode == _lib.MLM_TRAIN else None)
        self.row_code = code
        return MaskingInfo(mask, labels)

    def apply_mask_to_inputs(self, inputs, mask_schema, training=False, testing=False):
        if not testing and not training:
            # masking.py:489-492: one extra [MASK] position (copy of the last one, then replaced)
            inputs = torch.cat([inputs, inputs[:, -1, :].unsqueeze(1)], dim=1)
        return self._apply_codes(inputs, self.row_code)


@masking_registry.register_with_multiple_names("plm", "permutation")
class PermutationLanguageModeling(MaskSequence):
    """masking.py:501-740 (XLNet permutation language modeling).  Labels, mask and the permutation attention mask come
    from one kernel (``t4r_mask_plm``, one thread per session; the reference loops over the sessions in python);
    ``target_mapping`` is the identity in every mode the reference builds it in and is only materialised when
    somebody reads the attribute.  ``permute_all=True`` is not supported: the reference leaves ``target_mapping`` all
    zero in that mode (masking.py:592-596), which silences the query stream."""

    def __init__(self, hidden_size: int, padding_idx: int = 0, eval_on_last_item_seq_only: bool = True,
                 plm_probability: float = 1 / 6, max_span_length: int = 5, permute_all: bool = False, **kwargs):
        super().__init__(hidden_size=hidden_size, padding_idx=padding_idx,
                         eval_on_last_item_seq_only=eval_on_last_item_seq_only)
        if permute_all:
            raise NotImplementedError("PermutationLanguageModeling(permute_all=True) is outside the t4r_b200 path: the "
                                      "reference builds an all-zero target_mapping in that mode")
        self.plm_probability = plm_probability
        self.max_span_length = max_span_length
        self.permute_all = permute_all
        self.perm_mask: Optional[torch.Tensor] = None      # uint8 [B, L, L], 1 = query i may not attend key j
        self._plm_draws: Optional[dict] = None

    def set_draws(self, draws: Optional[dict]):
        """Test hook: dict(u_span, u_start [B, L], u_force, u_unmask [B], perm [B, L]) (DESIGN.md "Random draws")."""
        self._plm_draws = draws

    @property
    def target_mapping(self) -> Optional[torch.Tensor]:
        if self.perm_mask is None:
            return None
        B, L, _ = self.perm_mask.shape
        return torch.eye(L, dtype=torch.float32, device=self.perm_mask.device).expand(B, L, L)

    def _compute_masked_targets(self, item_ids, training=False, testing=False) -> MaskingInfo:
        # masking.py:729-737: only `training` selects the mode (evaluation and inference share the masks)
        mode = _lib.PLM_TRAIN if training else (_lib.PLM_EVAL_LAST if self.eval_on_last_item_seq_only
                                                 else _lib.PLM_EVAL_ALL)
        mask, labels, perm_mask = ops.mask_plm(item_ids, mode, self.padding_idx, self.max_span_length,
                                               self.plm_probability, self._plm_draws if training else None)
        self.perm_mask = perm_mask
        # masking.py:155-180 (the base rule PLM inherits): masked rows are replaced in training and evaluation only
        self.row_code = mask.to(torch.uint8) if (training or testing) else torch.zeros_like(mask, dtype=torch.uint8)
        return MaskingInfo(mask, labels)

    def apply_mask_to_inputs(self, inputs, mask_schema, training=False, testing=False):
        if not training and not testing:
            return inputs
        return self._apply_codes(inputs, self.row_code)

    def transformer_required_arguments(self) -> Dict[str, Any]:
        return dict(target_mapping=self.target_mapping, perm_mask=self.perm_mask)
