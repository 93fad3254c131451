"""Transformer configs: host-side mirror of ``transformers4rec/config/transformer.py``
for the two architectures on the hot path (XLNet :420-482, GPT-2 :205-260).  Like
the reference they subclass the Hugging Face config classes, so ``.build(...)``
produces the very same hyper-parameters (incl. the GPT-2 LayerNorm-eps quirk)."""
from __future__ import annotations

import transformers


class _Registry(dict):
    def register(self, name):
        def deco(cls):
            self[name] = cls
            return cls
        return deco

    def parse(self, name):
        if isinstance(name, str):
            if name not in self:
                raise ValueError(f"{name} is not a registered transformer; available: {sorted(self)}")
            return self[name]
        return name


transformer_registry = _Registry()


class T4RecConfig:
    """config/transformer.py:50-135."""

    def to_huggingface_torch_model(self):
        """Reference name kept; returns the t4r_b200 encoder (HF-compatible state dict)."""
        from .block import _encoder_from
        return _encoder_from(self)

    def to_torch_model(self, input_features, *prediction_task, task_blocks=None, task_weights=None,
                       loss_reduction="mean", **kwargs):
        from . import torch as torch4rec
        if not isinstance(input_features, torch4rec.TabularSequenceFeatures):
            raise ValueError("`input_features` must an instance of SequentialTabularFeatures")
        if not all(isinstance(t, torch4rec.PredictionTask) for t in prediction_task):
            raise ValueError("`task` is of the wrong type, please provide one or multiple instance(s) of PredictionTask")
        body = torch4rec.SequentialBlock(input_features,
                                         torch4rec.TransformerBlock(self, masking=input_features.masking))
        return torch4rec.Head(body, *prediction_task, task_blocks=task_blocks, task_weights=task_weights,
                              loss_reduction=loss_reduction).to_model(**kwargs)

    @property
    def transformers_config_cls(self):
        return self.__class__.__bases__[1]


@transformer_registry.register("gtp2")
class GPT2Config(T4RecConfig, transformers.GPT2Config):
    @classmethod
    def build(cls, d_model, n_head, n_layer, total_seq_length, hidden_act="gelu", initializer_range=0.01,
              layer_norm_eps=0.03, dropout=0.3, pad_token=0, log_attention_weights=False, **kwargs):
        return cls(n_embd=d_model, n_inner=d_model * 4, n_layer=n_layer, n_head=n_head,
                   activation_function=hidden_act, initializer_range=initializer_range,
                   layer_norm_eps=layer_norm_eps, resid_pdrop=dropout, embd_pdrop=dropout, attn_pdrop=dropout,
                   n_positions=total_seq_length, n_ctx=total_seq_length, output_attentions=log_attention_weights,
                   vocab_size=1, **kwargs)


transformer_registry["gpt2"] = GPT2Config


@transformer_registry.register("xlnet")
class XLNetConfig(T4RecConfig, transformers.XLNetConfig):
    @classmethod
    def build(cls, d_model, n_head, n_layer, total_seq_length=None, attn_type="bi", hidden_act="gelu",
              initializer_range=0.01, layer_norm_eps=0.03, dropout=0.3, pad_token=0, log_attention_weights=False,
              mem_len=1, **kwargs):
        return cls(d_model=d_model, d_inner=d_model * 4, n_layer=n_layer, n_head=n_head, attn_type=attn_type,
                   ff_activation=hidden_act, initializer_range=initializer_range, layer_norm_eps=layer_norm_eps,
                   dropout=dropout, pad_token_id=pad_token, output_attentions=log_attention_weights, vocab_size=1,
                   mem_len=mem_len, **kwargs)
