"""Blocks: SequentialBlock, MLPBlock/DenseBlock, TransformerBlock and the encoder
parameter holders.

Reference: transformers4rec/torch/block/base.py:236-262 (kwarg routing),
block/mlp.py:30-144, block/transformer.py:76-206.  The encoder parameter
modules keep the Hugging Face state-dict names (``layer.N.rel_attn.q`` ...,
``h.N.attn.c_attn.weight`` ...) so checkpoints load unchanged, but the arithmetic
runs in ``t4r_xlnet_encoder_fwd`` / ``t4r_gpt2_encoder_fwd``.
"""
from __future__ import annotations

import ctypes as C
import inspect
from typing import Optional

import torch
from torch import nn

from . import _lib, ops


def _planes_of(x: torch.Tensor):
    return getattr(x, "_t4r_planes", None)


def right_shift_block(left, right):
    """block/base.py:66-67, :265-311 (``a >> b``): chain two blocks, flattening sequential blocks, so that
    ``features >> MLPBlock([64]) >> TransformerBlock(...)`` builds the body the way the reference's examples do."""
    parts = []
    for side in (left, right):
        parts += list(side) if isinstance(side, SequentialBlock) and not isinstance(side, _BuiltMLP) else [side]
    return SequentialBlock(*parts)


class _RShift:
    def __rshift__(self, other):
        return right_shift_block(self, other)


class SequentialBlock(nn.Sequential, _RShift):
    """block/base.py:160-262: passes ``training`` / ``testing`` (and friends) only to
    layers whose ``forward`` names them."""

    def __init__(self, *args, output_size=None):
        # block/base.py:174-203: buildable blocks (e.g. ``MLPBlock([64])``) are built against the
        # output size of the layer in front of them
        built = []
        for layer in args:
            if not isinstance(layer, nn.Module) and hasattr(layer, "build"):
                if not built or not hasattr(built[-1], "output_size") or built[-1].output_size() is None:
                    raise ValueError("a buildable block needs a preceding layer with a known output size")
                layer = layer.build(built[-1].output_size())
            built.append(layer)
        super().__init__(*built)
        self._static_output_size = output_size
        self.input_size = None

    @property
    def inputs(self):
        from .features import TabularSequenceFeatures
        first = list(self)[0]
        return first if isinstance(first, TabularSequenceFeatures) else None

    def forward(self, input, training=False, testing=False, **kwargs):
        for i, layer in enumerate(self):
            params = inspect.signature(layer.forward).parameters
            if i == len(self) - 1:
                filtered = {k: v for k, v in kwargs.items() if k in params}
            else:
                filtered = {}
            if "training" in params:
                filtered["training"] = training
            if "testing" in params:
                filtered["testing"] = testing
            input = layer(input, **filtered) if filtered else layer(input)
        return input

    def output_size(self, input_size=None):
        if self._static_output_size is not None:
            return self._static_output_size
        last = list(self)[-1]
        if hasattr(last, "output_size"):
            return last.output_size(input_size)
        return None

    def build(self, input_size, **kwargs):
        return self


class DenseBlock(nn.Sequential):
    """block/mlp.py:90-144: Linear (+ activation(inplace=True))."""

    def __init__(self, input_shape, in_features: int, out_features: int, activation=torch.nn.ReLU,
                 use_bias: bool = True, dropout: Optional[float] = None, normalization=None):
        args = [torch.nn.Linear(in_features, out_features, bias=use_bias)]
        if activation:
            args.append(activation(inplace=True) if activation in (nn.ReLU,) else activation())
        if normalization or dropout:
            raise NotImplementedError("MLPBlock normalization/dropout are outside the t4r_b200 hot path")
        super().__init__(*args)
        self._input_shape = input_shape
        self._output_size = out_features
        self._planes = ops.PlaneCache()

    def act_code(self) -> int:
        if len(self) == 1:
            return _lib.ACT_NONE
        if isinstance(self[1], nn.ReLU):
            return _lib.ACT_RELU
        if isinstance(self[1], nn.GELU):
            return _lib.ACT_GELU
        raise NotImplementedError(f"activation {type(self[1]).__name__} is not supported by the fused dense layer")

    def forward(self, x: torch.Tensor):
        lin: nn.Linear = self[0]
        shp = x.shape
        K = shp[-1]
        planes = _planes_of(x)
        if planes is None:
            planes = ops.split_planes(x.reshape(-1, K))
        w_planes = self._planes.get("w", lin.weight)
        y, y_planes, _ = ops.linear(planes, w_planes, K, bias=lin.bias, act=self.act_code())
        y = y.view(*shp[:-1], lin.out_features)
        y._t4r_planes = y_planes
        return y

    def forward_output_size(self, input_size):
        return torch.Size(list(input_size[:-1]) + [self._output_size])


class _BuiltMLP(SequentialBlock):
    def fusable_linear(self):
        return self[0][0] if len(self) == 1 else None

    def fusable_activation(self) -> int:
        return self[0].act_code()

    def output_size(self, input_size=None):
        base = list(self.input_size) if self.input_size is not None else [-1, -1, -1]
        return torch.Size(base[:-1] + [self[-1]._output_size])


class MLPBlock(_RShift):
    """block/mlp.py:30-87 (a BuildableBlock: ``build(input_shape)`` returns the module)."""

    def __init__(self, dimensions, activation=torch.nn.ReLU, use_bias: bool = True, dropout: float = None,
                 normalization: str = None, filter_features=None) -> None:
        if isinstance(dimensions, int):
            dimensions = [dimensions]
        self.normalization = normalization
        self.dropout = dropout
        self.filter_features = filter_features
        self.use_bias = use_bias
        self.activation = activation
        self.dimensions = dimensions

    def build(self, input_shape) -> SequentialBlock:
        layer_input_sizes = list(input_shape[-1:]) + list(self.dimensions[:-1])
        blocks = [DenseBlock(input_shape, i, o, activation=self.activation, use_bias=self.use_bias,
                             dropout=self.dropout, normalization=self.normalization)
                  for i, o in zip(layer_input_sizes, self.dimensions)]
        out = _BuiltMLP(*blocks)
        out.input_size = input_shape
        return out


# --------------------------------------------------------------------------- #
# encoder parameter holders (HF state-dict names)
# --------------------------------------------------------------------------- #
class _XLNetRelAttn(nn.Module):
    def __init__(self, d, H, eps, std):
        super().__init__()
        dh = d // H
        for n in ("q", "k", "v", "o", "r"):
            setattr(self, n, nn.Parameter(torch.empty(d, H, dh).normal_(0.0, std)))
        for n in ("r_r_bias", "r_s_bias", "r_w_bias"):
            setattr(self, n, nn.Parameter(torch.empty(H, dh).normal_(0.0, std)))
        self.seg_embed = nn.Parameter(torch.empty(2, H, dh).normal_(0.0, std))
        self.layer_norm = nn.LayerNorm(d, eps=eps)


class _XLNetFF(nn.Module):
    def __init__(self, d, d_inner, eps, std):
        super().__init__()
        self.layer_norm = nn.LayerNorm(d, eps=eps)
        self.layer_1 = nn.Linear(d, d_inner)
        self.layer_2 = nn.Linear(d_inner, d)
        for lin in (self.layer_1, self.layer_2):
            lin.weight.data.normal_(0.0, std)
            lin.bias.data.zero_()


class _XLNetLayer(nn.Module):
    def __init__(self, d, H, d_inner, eps, std):
        super().__init__()
        self.rel_attn = _XLNetRelAttn(d, H, eps, std)
        self.ff = _XLNetFF(d, d_inner, eps, std)


class XLNetEncoder(nn.Module):
    """Parameters of HF ``XLNetModel`` as the reference configures it
    (config/transformer.py:467-482); initialised like HF ``_init_weights``
    (normal(0, initializer_range), LayerNorm = identity)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        d, H = config.d_model, config.n_head
        std = config.initializer_range
        self.word_embedding = nn.Embedding(config.vocab_size, d)   # unused on this path (vocab_size=1)
        self.mask_emb = nn.Parameter(torch.empty(1, 1, d).normal_(0.0, std))  # unused (no target_mapping)
        self.layer = nn.ModuleList([_XLNetLayer(d, H, config.d_inner, config.layer_norm_eps, std)
                                    for _ in range(config.n_layer)])
        self._planes = ops.PlaneCache()

    def _layer_structs(self):
        d = self.config.d_model
        n = len(self.layer)
        arr = (_lib.XLNetLayer * n)()
        keep = []
        for i, lyr in enumerate(self.layer):
            ra, ff = lyr.rel_attn, lyr.ff
            # B operand rows = output features: [q | k | v] each (H*dh, d_in) = W.reshape(d, HD).T
            sig_params = (ra.q, ra.k, ra.v)
            key = f"qkv{i}"
            ent = self._planes._cache.get(key)
            sig = tuple((p.data_ptr(), p._version) for p in sig_params)
            if ent is None or ent[0] != sig:
                with torch.no_grad():
                    w = torch.cat([p.detach().reshape(d, d).t() for p in sig_params], dim=0).contiguous()
                    self._planes._cache[key] = (sig, ops.split_planes(w))
            wqkv = self._planes._cache[key][1]
            wo = self._planes.get(f"o{i}", ra.o, lambda w: w.reshape(d, d))
            w1 = self._planes.get(f"w1{i}", ff.layer_1.weight)
            w2 = self._planes.get(f"w2{i}", ff.layer_2.weight)
            wr = ra.r.detach().reshape(d, d).contiguous()
            ts = [wqkv, wr, ra.r_w_bias.detach().reshape(-1).contiguous(), ra.r_r_bias.detach().reshape(-1).contiguous(),
                  wo, ra.layer_norm.weight.detach(), ra.layer_norm.bias.detach(), w1, ff.layer_1.bias.detach(), w2,
                  ff.layer_2.bias.detach(), ff.layer_norm.weight.detach(), ff.layer_norm.bias.detach()]
            keep += ts
            for (fname, _), t in zip(_lib.XLNetLayer._fields_, ts):
                setattr(arr[i], fname, t.data_ptr())
        return arr, keep

    def forward(self, inputs_embeds: torch.Tensor, perm_mask: Optional[torch.Tensor] = None, target_mapping=None, **kwargs):
        B, L, d = inputs_embeds.shape
        arr, keep = self._layer_structs()
        if perm_mask is not None:
            # permutation language modeling (HF:xlnet two-stream attention with target_mapping = identity): the
            # content stream h and the query stream g (mask_emb in every row) go through the layers stacked; HF
            # returns g
            x2 = torch.cat([inputs_embeds.reshape(B * L, d).float(),
                            self.mask_emb.detach().reshape(1, d).float().expand(B * L, d)], dim=0)
            out2 = ops.xlnet_encoder_plm(arr, len(self.layer), B, L, d, self.config.n_head,
                                         float(self.config.layer_norm_eps), x2, perm_mask)
            return (out2[B * L:].view(B, L, d),)
        planes = _planes_of(inputs_embeds)
        out, out_planes = ops.xlnet_encoder(arr, len(self.layer), B, L, d, self.config.n_head,
                                            float(self.config.layer_norm_eps), inputs_embeds.reshape(B * L, d), planes,
                                            want_planes=False)
        out = out.view(B, L, d)
        return (out,)


class _GPT2Conv1D(nn.Module):
    def __init__(self, nf, nx, std):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(nx, nf).normal_(0.0, std))
        self.bias = nn.Parameter(torch.zeros(nf))


class _GPT2Attn(nn.Module):
    def __init__(self, d, std):
        super().__init__()
        self.c_attn = _GPT2Conv1D(3 * d, d, std)
        self.c_proj = _GPT2Conv1D(d, d, std)


class _GPT2MLP(nn.Module):
    def __init__(self, d, inner, std):
        super().__init__()
        self.c_fc = _GPT2Conv1D(inner, d, std)
        self.c_proj = _GPT2Conv1D(d, inner, std)


class _GPT2Block(nn.Module):
    def __init__(self, d, inner, eps, std):
        super().__init__()
        self.ln_1 = nn.LayerNorm(d, eps=eps)
        self.attn = _GPT2Attn(d, std)
        self.ln_2 = nn.LayerNorm(d, eps=eps)
        self.mlp = _GPT2MLP(d, inner, std)


class GPT2Encoder(nn.Module):
    """Parameters of HF ``GPT2Model`` as the reference configures it
    (config/transformer.py:244-260).  LayerNorm eps is HF's
    ``layer_norm_epsilon`` (1e-5: the reference's ``layer_norm_eps`` kwarg is not
    read by HF's GPT2Config)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        d = config.n_embd
        std = config.initializer_range
        inner = config.n_inner if config.n_inner is not None else 4 * d
        self.wte = nn.Embedding(config.vocab_size, d)
        self.wpe = nn.Embedding(config.n_positions, d)
        self.wpe.weight.data.normal_(0.0, std)
        self.h = nn.ModuleList([_GPT2Block(d, inner, config.layer_norm_epsilon, std) for _ in range(config.n_layer)])
        self.ln_f = nn.LayerNorm(d, eps=config.layer_norm_epsilon)
        # HF _init_weights: c_proj weights get std / sqrt(2 * n_layer)
        import math
        for blk in self.h:
            for p in (blk.attn.c_proj.weight, blk.mlp.c_proj.weight):
                p.data.normal_(0.0, std / math.sqrt(2 * config.n_layer))
        self._planes = ops.PlaneCache()

    def _layer_structs(self):
        n = len(self.h)
        arr = (_lib.GPT2Layer * n)()
        keep = []
        tr = lambda w: w.t().contiguous()  # Conv1D [in, out] -> [out, in]
        for i, blk in enumerate(self.h):
            ts = [blk.ln_1.weight.detach(), blk.ln_1.bias.detach(),
                  self._planes.get(f"qkv{i}", blk.attn.c_attn.weight, tr), blk.attn.c_attn.bias.detach(),
                  self._planes.get(f"o{i}", blk.attn.c_proj.weight, tr), blk.attn.c_proj.bias.detach(),
                  blk.ln_2.weight.detach(), blk.ln_2.bias.detach(),
                  self._planes.get(f"fc{i}", blk.mlp.c_fc.weight, tr), blk.mlp.c_fc.bias.detach(),
                  self._planes.get(f"pr{i}", blk.mlp.c_proj.weight, tr), blk.mlp.c_proj.bias.detach()]
            keep += ts
            for (fname, _), t in zip(_lib.GPT2Layer._fields_, ts):
                setattr(arr[i], fname, t.data_ptr())
        return arr, keep

    def forward(self, inputs_embeds: torch.Tensor, **kwargs):
        B, L, d = inputs_embeds.shape
        if L > self.config.n_positions:
            raise ValueError(f"sequence length {L} exceeds n_positions {self.config.n_positions}")
        arr, keep = self._layer_structs()
        out, _ = ops.gpt2_encoder(arr, len(self.h), B, L, d, self.config.n_head, float(self.config.layer_norm_epsilon),
                                  self.wpe.weight.detach(), self.ln_f.weight.detach(), self.ln_f.bias.detach(),
                                  inputs_embeds.reshape(B * L, d))
        return (out.view(B, L, d),)


def _encoder_from(transformer) -> nn.Module:
    """block/transformer.py:100-107: accept a T4RecConfig, an HF config or an HF model."""
    import transformers as hf
    if isinstance(transformer, (XLNetEncoder, GPT2Encoder)):
        return transformer
    if isinstance(transformer, hf.PreTrainedModel):
        enc = _encoder_from(transformer.config)
        missing, unexpected = enc.load_state_dict(transformer.state_dict(), strict=False)
        if missing:
            raise ValueError(f"cannot map HF model onto the t4r_b200 encoder, missing: {missing}")
        return enc.to(next(transformer.parameters()).device)
    if isinstance(transformer, hf.XLNetConfig):
        if getattr(transformer, "attn_type", "bi") != "bi" or transformer.ff_activation != "gelu":
            raise NotImplementedError("t4r_b200 XLNet encoder supports attn_type='bi', ff_activation='gelu'")
        return XLNetEncoder(transformer)
    if isinstance(transformer, hf.GPT2Config):
        if transformer.activation_function != "gelu":
            raise NotImplementedError("t4r_b200 GPT-2 encoder supports activation_function='gelu'")
        return GPT2Encoder(transformer)
    raise NotImplementedError(
        f"{type(transformer).__name__}: only XLNet and GPT-2 are on the t4r_b200 hot path (SURVEY §2 row 9)")


class TransformerBlock(nn.Module, _RShift):
    """block/transformer.py:76-206."""

    def __init__(self, transformer, masking=None, prepare_module=None):
        super().__init__()
        self.transformer = _encoder_from(transformer)
        if masking is not None:
            from .masking import CausalLanguageModeling, MaskedLanguageModeling
            # utils/torch_utils.py:441-473 MappingTransformerMasking: GPT-2 is CLM-only
            if isinstance(self.transformer, GPT2Encoder) and isinstance(masking, MaskedLanguageModeling):
                raise ValueError(f"{masking.__class__.__name__} is not supported by: the GPT2Config architecture")
            required = list(masking.transformer_required_arguments().keys())
            if required and isinstance(self.transformer, XLNetEncoder) and set(required) <= {"target_mapping", "perm_mask"}:
                required = []  # permutation language modeling: XLNet's two-stream forward takes both
            if required:
                raise ValueError(f"{masking.__class__.__name__} requires the parameters: {', '.join(required)} "
                                 f"in the {type(self.transformer)} signature")
        self.masking = masking
        self.prepare_module = None  # GPT2Prepare's tril head_mask is implied by the causal kernel

    @classmethod
    def from_registry(cls, transformer: str, d_model: int, n_head: int, n_layer: int, total_seq_length: int,
                      masking=None):
        from .config import transformer_registry
        _t = transformer_registry.parse(transformer).build(d_model=d_model, n_head=n_head, n_layer=n_layer,
                                                           total_seq_length=total_seq_length)
        return cls(_t, masking)

    def forward(self, inputs_embeds, **kwargs):
        # block/transformer.py:185-196: the masking's transformer arguments travel with the call
        perm_mask = getattr(self.masking, "perm_mask", None) if self.masking is not None else None
        if perm_mask is not None:
            return self.transformer(inputs_embeds=inputs_embeds, perm_mask=perm_mask)[0]
        return self.transformer(inputs_embeds=inputs_embeds)[0]

    def _get_name(self):
        return "TansformerBlock"

    def forward_output_size(self, input_size):
        assert len(input_size) == 3
        return torch.Size([input_size[0], input_size[1], self.transformer.config.hidden_size])

    def output_size(self, input_size=None):
        return torch.Size([-1, -1, self.transformer.config.hidden_size])
