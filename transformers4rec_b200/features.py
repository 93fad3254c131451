"""Input block: host-side mirror of ``TabularSequenceFeatures`` and its parts.

Reference: transformers4rec/torch/features/sequence.py:97-296 (orchestration),
features/embedding.py:51-257,416-514 (tables), features/continuous.py:60-63,
tabular/aggregation.py:35-47 (concat in sorted-name order), block/mlp.py:123-144
(projection).  ``forward`` issues three kernels for what the reference runs as
F gathers + cat + GEMM + ReLU + ~15 masking ops + where:
  t4r_mask_*  ->  t4r_embed_concat_fwd  ->  t4r_linear_fwd (bias+ReLU+mask epilogue).

The other input-block options of the reference -- soft one-hot embeddings of
continuous features (features/embedding.py:517-556), ``continuous_projection``
(features/tabular.py:88-118), per-feature LayerNorm (tabular/transformations.py:95-141), the
element-wise aggregations (tabular/aggregation.py:139-193) and StochasticSwapNoise
(tabular/transformations.py:29-92) -- run through the general kernel ``t4r_input_block_fwd`` /
``t4r_swap_noise`` instead of the specialised gather.
"""
from __future__ import annotations

import os
from functools import partial
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import torch
from torch import nn

from . import _lib, ops
from .block import MLPBlock, SequentialBlock
from .masking import MaskSequence, masking_registry
from .schema import Schema, Tags, categorical_cardinalities


class TableConfig:
    """features/embedding.py:416-480."""

    def __init__(self, vocabulary_size: int, dim: int, initializer: Optional[Callable] = None,
                 combiner: str = "mean", name: Optional[str] = None):
        if not isinstance(vocabulary_size, int) or vocabulary_size < 1:
            raise ValueError("Invalid vocabulary_size {}.".format(vocabulary_size))
        if not isinstance(dim, int) or dim < 1:
            raise ValueError("Invalid dim {}.".format(dim))
        if combiner not in ("mean", "sum", "sqrtn"):
            raise ValueError("Invalid combiner {}".format(combiner))
        if (initializer is not None) and (not callable(initializer)):
            raise ValueError("initializer must be callable if specified.")
        self.initializer_or_none = initializer
        self.initializer = initializer or partial(torch.nn.init.normal_, mean=0.0, std=0.05)
        self.vocabulary_size = vocabulary_size
        self.dim = dim
        self.combiner = combiner
        self.name = name


class FeatureConfig:
    """features/embedding.py:483-514."""

    def __init__(self, table: TableConfig, max_sequence_length: int = 0, name: Optional[str] = None):
        self.table = table
        self.max_sequence_length = max_sequence_length
        self.name = name


class TabularLayerNorm(nn.Module):
    """tabular/transformations.py:95-141: one ``nn.LayerNorm(dim)`` per feature, applied before the
    aggregation (features of width 1 are skipped, like upstream)."""

    def __init__(self, features_dim: Optional[Dict[str, int]] = None):
        super().__init__()
        self.feature_layer_norm = nn.ModuleDict()
        for fname, dim in (features_dim or {}).items():
            if dim == 1:
                continue
            self.feature_layer_norm[fname] = nn.LayerNorm(normalized_shape=dim)

    @classmethod
    def from_feature_config(cls, feature_config: Dict[str, "FeatureConfig"]):
        return cls({name: f.table.dim for name, f in feature_config.items()})

    def params(self, name: str):
        if name in self.feature_layer_norm:
            m = self.feature_layer_norm[name]
            return m.weight, m.bias
        return None

    def forward(self, inputs: Dict[str, torch.Tensor], **kwargs):
        out = {}
        for key, val in inputs.items():
            ln = self.params(key)
            if ln is None:
                out[key] = val
                continue
            dim = val.shape[-1]
            y, _, _ = ops.input_block([dict(kind=_lib.FEAT_DENSE, dim=dim, col=0, input=val, ln=ln)],
                                      val.numel() // dim, 1, dim, ln_eps=self.feature_layer_norm[key].eps)
            out[key] = y.view(val.shape)
        return out


def _parse_post(post, feature_config) -> Optional[TabularLayerNorm]:
    if post is None:
        return None
    if isinstance(post, str):
        if post not in ("layer-norm",):
            raise NotImplementedError(f"post transformation '{post}' is not on the t4r_b200 path (layer-norm only)")
        return TabularLayerNorm.from_feature_config(feature_config)
    if isinstance(post, TabularLayerNorm):
        return post
    raise NotImplementedError(f"post transformation {type(post).__name__} is not on the t4r_b200 path")


class SoftEmbedding(nn.Module):
    """features/embedding.py:517-556: a continuous scalar becomes the softmax(Linear(1, n)(x))
    weighted mean of n embedding rows (soft one-hot encoding)."""

    def __init__(self, num_embeddings: int, embeddings_dim: int, emb_initializer=None):
        assert num_embeddings > 0, "The number of embeddings for soft embeddings needs to be greater than 0"
        assert embeddings_dim > 0, "The embeddings dim for soft embeddings needs to be greater than 0"
        super().__init__()
        self.embedding_table = nn.Embedding(num_embeddings, embeddings_dim)
        if emb_initializer:
            emb_initializer(self.embedding_table.weight)
        self.projection_layer = nn.Linear(1, num_embeddings, bias=True)

    def feature(self, x: torch.Tensor, **extra) -> dict:
        return dict(kind=_lib.FEAT_SOFT, dim=self.embedding_table.embedding_dim, input=x,
                    table=self.embedding_table.weight, soft_w=self.projection_layer.weight,
                    soft_b=self.projection_layer.bias, **extra)

    def forward(self, input_numeric: torch.Tensor):
        dim = self.embedding_table.embedding_dim
        y, _, _ = ops.input_block([self.feature(input_numeric, col=0)], input_numeric.numel(), 1, dim)
        return y.view(*input_numeric.shape, dim)


class SoftEmbeddingFeatures(nn.Module):
    """features/embedding.py:278-412: one SoftEmbedding per continuous feature, LayerNorm'd by default."""

    def __init__(self, feature_config: Dict[str, FeatureConfig], layer_norm: bool = True, post=None):
        super().__init__()
        self.feature_config = feature_config
        self.features = list(feature_config.keys())
        self.embedding_tables = nn.ModuleDict({
            name: SoftEmbedding(f.table.vocabulary_size, f.table.dim, f.table.initializer_or_none)
            for name, f in feature_config.items()})
        self.post = TabularLayerNorm.from_feature_config(feature_config) if layer_norm else _parse_post(post, feature_config)

    @classmethod
    def from_schema(cls, schema: Schema, soft_embedding_cardinalities=None, soft_embedding_cardinality_default: int = 10,
                    soft_embedding_dims=None, soft_embedding_dim_default: int = 8, embeddings_initializers=None,
                    layer_norm: bool = True, combiner: str = "mean", tags=None, **kwargs):
        if tags:
            schema = schema.select_by_tag(tags)
        soft_embedding_cardinalities = soft_embedding_cardinalities or {}
        soft_embedding_dims = soft_embedding_dims or {}
        embeddings_initializers = embeddings_initializers or {}
        cardinalities = categorical_cardinalities(schema)
        feature_config = {}
        for col_name in schema.column_names:
            if col_name in cardinalities:
                continue
            feature_config[col_name] = FeatureConfig(TableConfig(
                vocabulary_size=soft_embedding_cardinalities.get(col_name, soft_embedding_cardinality_default),
                dim=soft_embedding_dims.get(col_name, soft_embedding_dim_default), name=col_name, combiner=combiner,
                initializer=embeddings_initializers.get(col_name, None)))
        if not feature_config:
            return None
        return cls(feature_config, layer_norm=layer_norm)

    def output_dims(self) -> Dict[str, int]:
        return {n: f.table.dim for n, f in self.feature_config.items()}

    def forward(self, inputs, **kwargs):
        out = {k: self.embedding_tables[k](inputs[k]) for k in self.features if k in inputs}
        return self.post(out) if self.post is not None else out


class ContinuousProjection(nn.Module):
    """features/tabular.py:88-118 (project_continuous_features): the scalar features are concatenated
    (sorted names) and sent through an MLP; the result is ONE feature named "continuous_projection"."""

    name = "continuous_projection"

    def __init__(self, continuous: "ContinuousFeatures", dimensions, sequence_length: int = -1):
        super().__init__()
        if isinstance(dimensions, int):
            dimensions = [dimensions]
        self.continuous = continuous
        self.features = sorted(continuous.features)
        self.mlp = MLPBlock(list(dimensions)).build(torch.Size([-1, sequence_length, len(self.features)]))
        self.dim = int(list(dimensions)[-1])

    def project(self, inputs, expand) -> torch.Tensor:
        """-> fp32 [M, dim]; ``expand`` flattens a feature to [M] (repeating context features)."""
        vals = [(expand(inputs[n]), i) for i, n in enumerate(self.features)]
        M = vals[0][0].numel()
        x, _, _ = ops.embed_concat([], vals, M, len(vals), want_f32=True, want_planes=False)
        return self.mlp(x.view(M, 1, len(vals))).reshape(M, self.dim)

    def forward(self, inputs, **kwargs):
        some = inputs[self.features[0]]
        y = self.project(inputs, lambda v: v.reshape(-1))
        return {self.name: y.view(*some.shape, self.dim)}


class StochasticSwapNoise(nn.Module):
    """tabular/transformations.py:29-92.  Training-time augmentation of the RAW inputs: a position that
    is not padding is, with probability ``replacement_prob``, overwritten by a value drawn from the
    non-padded values of the same feature in the batch.  Integer/copy work: bit-exact for given draws.
    ``set_draws({name: (u, perm)})`` injects the draws (tests); otherwise they come from torch's
    CUDA generator, drawn in the same order upstream draws them (bernoulli, then randperm)."""

    def __init__(self, schema=None, pad_token: int = 0, replacement_prob: float = 0.1):
        super().__init__()
        self.schema = schema
        self.pad_token = pad_token
        self.replacement_prob = replacement_prob
        self._draws = None

    def set_draws(self, draws):
        self._draws = draws

    def get_padding_mask_from_item_id(self, inputs, pad_token=0):
        item = self.schema.select_by_tag(Tags.ITEM_ID).column_names[0] if self.schema is not None else None
        return (inputs[item] != pad_token) if item is not None else None

    def augment(self, input_tensor: torch.Tensor, mask: Optional[torch.Tensor] = None, draws=None) -> torch.Tensor:
        if not self.training:
            return input_tensor
        eff = mask
        if mask is not None and input_tensor.dim() == mask.dim() - 1:
            eff = mask[:, 0]
        if draws is None:
            u = torch.rand(input_tensor.shape, device=input_tensor.device)
            n_pool = int(eff.sum()) if eff is not None else input_tensor.numel()
            perm = torch.randperm(n_pool, device=input_tensor.device)
        else:
            u, perm = draws
        x = input_tensor if input_tensor.dtype in (torch.int64, torch.float32) else (
            input_tensor.long() if not input_tensor.is_floating_point() else input_tensor.float())
        return ops.swap_noise(x, mask, u, perm, self.replacement_prob).to(input_tensor.dtype)

    def forward(self, inputs, input_mask: Optional[torch.Tensor] = None, **kwargs):
        if self.schema is not None and input_mask is None:
            input_mask = self.get_padding_mask_from_item_id(inputs, self.pad_token)
        if isinstance(inputs, dict):
            d = self._draws or {}
            return {k: self.augment(v, input_mask, d.get(k)) for k, v in inputs.items()}
        return self.augment(inputs, input_mask, self._draws)


class SequenceEmbeddingFeatures(nn.Module):
    """features/sequence.py:43-90 + features/embedding.py:51-257: owns one
    ``nn.Embedding(V, dim, padding_idx)`` per categorical feature (state-dict keys
    ``embedding_tables.<feature>.weight``) and remembers ``item_seq``."""

    def __init__(self, feature_config: Dict[str, FeatureConfig], item_id: Optional[str] = None, padding_idx: int = 0,
                 post=None, shard_item_table: bool = False, device=None):
        super().__init__()
        self.post = _parse_post(post, feature_config)
        self.padding_idx = padding_idx
        self.item_id = item_id
        self.feature_config = feature_config
        tables = {}
        for name, feature in feature_config.items():
            if shard_item_table and name == item_id:
                # BASELINE configs 4-5: the item table (= the tied output layer) is row-sharded over the ranks
                from .distributed import ShardedEmbedding
                tables[name] = ShardedEmbedding(feature.table.vocabulary_size, feature.table.dim, padding_idx,
                                                initializer=feature.table.initializer, device=device)
                continue
            emb = nn.Embedding(feature.table.vocabulary_size, feature.table.dim, padding_idx=padding_idx, device=device)
            if feature.table.initializer is not None:
                feature.table.initializer(emb.weight)
            tables[name] = emb
        self.embedding_tables = nn.ModuleDict(tables)
        self.item_seq: Optional[torch.Tensor] = None

    @property
    def item_embedding_table(self):
        assert self.item_id is not None
        return self.embedding_tables[self.item_id]

    @classmethod
    def from_schema(cls, schema: Schema, embedding_dims=None, embedding_dim_default: int = 64,
                    infer_embedding_sizes: bool = False, infer_embedding_sizes_multiplier: float = 2.0,
                    embeddings_initializers=None, combiner="mean", tags=None, item_id=None, padding_idx=0, post=None,
                    shard_item_table: bool = False, device=None, **kwargs):
        """features/embedding.py:103-221."""
        if tags:
            schema = schema.select_by_tag(tags)
        _item_id = schema.select_by_tag(Tags.ITEM_ID)
        if not item_id and len(_item_id) > 0:
            if len(_item_id) > 1:
                raise ValueError("Multiple columns with tag ITEM_ID found. Please specify the item_id column name.")
            item_id = list(_item_id)[0].name
        embedding_dims = dict(embedding_dims or {})
        cardinalities = categorical_cardinalities(schema)
        if infer_embedding_sizes:
            for k, card in cardinalities.items():
                # utils/torch_utils.py get_embedding_size_from_cardinality: ceil(card**0.25 * multiplier)
                if k not in embedding_dims:
                    import math
                    embedding_dims[k] = int(math.ceil(math.pow(card, 0.25) * infer_embedding_sizes_multiplier))
        embeddings_initializers = embeddings_initializers or {}
        feature_config = {}
        for key, cardinality in cardinalities.items():
            feature_config[key] = FeatureConfig(TableConfig(
                vocabulary_size=cardinality, dim=embedding_dims.get(key, embedding_dim_default), name=key,
                combiner=combiner, initializer=embeddings_initializers.get(key, None)))
        if not feature_config:
            return None
        return cls(feature_config, item_id=item_id, padding_idx=padding_idx, post=post,
                   shard_item_table=shard_item_table, device=device)

    def item_ids(self, inputs) -> torch.Tensor:
        return inputs[self.item_id]

    def is_sharded(self, name: str) -> bool:
        return hasattr(self.embedding_tables[name], "lookup")

    def shard_item_table(self, group=None):
        """Replace the replicated item table by this rank's row block.  A head built with
        ``weight_tying=True`` follows automatically: it reads ``item_embedding_table`` at call time."""
        from .distributed import ShardedEmbedding
        assert self.item_id is not None
        if not self.is_sharded(self.item_id):
            full = self.embedding_tables[self.item_id].weight.detach()
            self.embedding_tables[self.item_id] = ShardedEmbedding.from_full(full, group, self.padding_idx)
        return self

    def output_dims(self) -> Dict[str, int]:
        return {n: f.table.dim for n, f in self.feature_config.items()}

    def forward(self, inputs, **kwargs):
        """Stand-alone use (dict of [B, L, dim] tensors, like the reference); the fused
        path in TabularSequenceFeatures does not go through here."""
        out = {}
        for name in self.feature_config:
            ids = inputs[name]
            shp = tuple(ids.shape)
            if self.is_sharded(name):
                out[name] = self.embedding_tables[name](ids)
                continue
            table = self.embedding_tables[name].weight
            of, _, _ = ops.embed_concat([(table.detach(), ids.reshape(-1), 0)], [], ids.numel(), table.shape[1], True,
                                        False)
            out[name] = of.view(*shp, table.shape[1])
        if self.item_id:
            self.item_seq = self.item_ids(inputs)
        return self.post(out) if self.post is not None else out


class ContinuousFeatures(nn.Module):
    """features/continuous.py:26-69: scalars become a trailing unit dimension."""

    def __init__(self, features: List[str]):
        super().__init__()
        self.features = list(features)

    @classmethod
    def from_schema(cls, schema: Schema, tags=None, **kwargs):
        if tags:
            schema = schema.select_by_tag(tags)
        if not schema.column_names:
            return None
        return cls(schema.column_names)

    def forward(self, inputs, **kwargs):
        return {k: inputs[k].unsqueeze(-1) for k in self.features if k in inputs}


class TabularSequenceFeatures(nn.Module):
    def __rshift__(self, other):   # block/base.py:66-67: ``features >> block``
        from .block import right_shift_block
        return right_shift_block(self, other)

    """features/sequence.py:97-296."""

    EMBEDDING_MODULE_CLASS = SequenceEmbeddingFeatures
    CONTINUOUS_MODULE_CLASS = ContinuousFeatures
    SOFT_EMBEDDING_MODULE_CLASS = SoftEmbeddingFeatures
    AGGREGATIONS = {"concat": _lib.AGG_CONCAT, "element-wise-sum": _lib.AGG_SUM,
                    "element-wise-sum-item-multi": _lib.AGG_SUM_ITEM_MULTI}

    def __init__(self, continuous_module=None, categorical_module=None, pretrained_embedding_module=None,
                 projection_module=None, masking: Optional[MaskSequence] = None, aggregation: Optional[str] = None,
                 schema: Optional[Schema] = None, pre=None, **kwargs):
        super().__init__()
        if isinstance(pre, str):
            if pre not in ("stochastic-swap-noise", "ssn"):
                raise NotImplementedError(f"pre transformation '{pre}' is not on the t4r_b200 path")
            pre = StochasticSwapNoise(schema=schema)
        self.pre = pre
        if pretrained_embedding_module is not None:
            raise NotImplementedError("pretrained embeddings are outside the t4r_b200 hot path (SURVEY §2 row 1)")
        to_merge = {}
        if continuous_module is not None:
            to_merge["continuous_module"] = continuous_module
        if categorical_module is not None:
            to_merge["categorical_module"] = categorical_module
        assert to_merge != {}, "Please provide at least one input layer"
        self.to_merge = nn.ModuleDict(to_merge)
        if aggregation is not None and aggregation not in self.AGGREGATIONS:
            raise NotImplementedError(f"aggregation '{aggregation}' is not on the t4r_b200 path "
                                      f"({', '.join(self.AGGREGATIONS)})")
        self.aggregation = aggregation
        self.schema = schema
        self.projection_module = projection_module
        self.set_masking(masking)
        self._planes = ops.PlaneCache()

    # ----------------------------------------------------------------- builders
    @classmethod
    def from_schema(cls, schema: Schema, continuous_tags=(Tags.CONTINUOUS,), categorical_tags=(Tags.CATEGORICAL,),
                    pretrained_embeddings_tags=(Tags.EMBEDDING,), aggregation: Optional[str] = None,
                    automatic_build: bool = True, max_sequence_length: Optional[int] = None,
                    continuous_projection=None, continuous_soft_embeddings: bool = False, projection=None,
                    d_output: Optional[int] = None, masking=None, **kwargs) -> "TabularSequenceFeatures":
        """features/sequence.py:140-229 + features/tabular.py:120-214."""
        cont = None
        if continuous_tags:
            if continuous_soft_embeddings:
                soft_kwargs = {k: v for k, v in kwargs.items() if k in (
                    "soft_embedding_cardinalities", "soft_embedding_cardinality_default", "soft_embedding_dims",
                    "soft_embedding_dim_default", "layer_norm")}
                cont = cls.SOFT_EMBEDDING_MODULE_CLASS.from_schema(schema, tags=continuous_tags, **soft_kwargs)
            else:
                cont = cls.CONTINUOUS_MODULE_CLASS.from_schema(schema, tags=continuous_tags)
        emb_kwargs = {k: v for k, v in kwargs.items() if k in (
            "embedding_dims", "embedding_dim_default", "infer_embedding_sizes", "infer_embedding_sizes_multiplier",
            "embeddings_initializers", "combiner", "item_id", "padding_idx", "post", "shard_item_table", "device")}
        cat = cls.EMBEDDING_MODULE_CLASS.from_schema(schema, tags=categorical_tags, **emb_kwargs) if categorical_tags else None
        if continuous_projection:
            if not automatic_build:
                raise ValueError("Continuous feature projection can only be done with automatic_build")
            if not isinstance(cont, ContinuousFeatures):
                raise ValueError("continuous_projection needs plain continuous features")
            cont = ContinuousProjection(cont, continuous_projection, max_sequence_length or -1)
        output = cls(continuous_module=cont, categorical_module=cat, aggregation=aggregation, schema=schema,
                     pre=kwargs.get("pre"))
        output.max_sequence_length = max_sequence_length
        if d_output and projection:
            raise ValueError("You cannot specify both d_output and projection at the same time")
        if (projection or masking or d_output) and not aggregation:
            output.aggregation = "concat"
        hidden_size = output.output_size()
        if d_output and not projection:
            projection = MLPBlock([d_output])
        if projection is not None and hasattr(projection, "build"):
            projection = projection.build(hidden_size)
        if projection is not None:
            output.projection_module = projection
            hidden_size = projection.output_size()
        if isinstance(masking, str):
            mk_kwargs = {k: v for k, v in kwargs.items() if k in (
                "padding_idx", "eval_on_last_item_seq_only", "mlm_probability", "train_on_last_item_seq_only")}
            masking = masking_registry.parse(masking)(hidden_size=output.output_size()[-1], **mk_kwargs)
        if masking and not getattr(output, "item_id", None):
            raise ValueError("For masking a categorical_module is required including an item_id.")
        output.set_masking(masking)
        return output

    @property
    def masking(self):
        return self._masking

    def set_masking(self, value):
        self._masking = value

    @property
    def categorical_module(self):
        return self.to_merge["categorical_module"] if "categorical_module" in self.to_merge else None

    @property
    def continuous_module(self):
        return self.to_merge["continuous_module"] if "continuous_module" in self.to_merge else None

    @property
    def item_id(self) -> Optional[str]:
        cm = self.categorical_module
        return getattr(cm, "item_id", None) if cm is not None else None

    @property
    def item_embedding_table(self):
        cm = self.categorical_module
        return getattr(cm, "item_embedding_table", None) if cm is not None else None

    # ----------------------------------------------------------------- shapes
    def _layout(self) -> Tuple[List[Tuple[str, str, int, int]], int]:
        """(name, kind, first column, width) in sorted-name order (aggregation.py:42-47); for the
        element-wise aggregations every column offset is 0 and the total is the common width."""
        widths = {}
        kinds = {}
        if self.categorical_module is not None:
            for n, dim in self.categorical_module.output_dims().items():
                widths[n], kinds[n] = dim, "cat"
        cont = self.continuous_module
        if isinstance(cont, ContinuousProjection):
            widths[cont.name], kinds[cont.name] = cont.dim, "dense"
        elif isinstance(cont, SoftEmbeddingFeatures):
            for n, dim in cont.output_dims().items():
                widths[n], kinds[n] = dim, "soft"
        elif cont is not None:
            for n in cont.features:
                widths[n], kinds[n] = 1, "cont"
        elementwise = self.aggregation in ("element-wise-sum", "element-wise-sum-item-multi")
        if elementwise and len(set(widths.values())) > 1:
            # aggregation.py:120-134 (_check_inputs_last_dim_equal)
            raise ValueError(f"The last dim of all input features is not equal, which is"
                             f" required for element-wise aggregation: {widths}")
        col = 0
        out = []
        for n in sorted(widths.keys()):
            out.append((n, kinds[n], 0 if elementwise else col, widths[n]))
            col += widths[n]
        return out, (next(iter(widths.values())) if elementwise else col)

    def output_size(self, input_size=None):
        L = getattr(self, "max_sequence_length", None) or -1
        if self.projection_module is not None:
            return torch.Size([-1, L, self.projection_module.output_size()[-1]])
        return torch.Size([-1, L, self._layout()[1]])

    def forward_output_size(self, input_size=None):
        return self.output_size(input_size)

    def build(self, *args, **kwargs):
        return self

    # ----------------------------------------------------------------- forward
    def forward(self, inputs: Dict[str, torch.Tensor], training: bool = False, testing: bool = False, **kwargs):
        if self.pre is not None:
            inputs = self.pre(inputs)
        layout, C_width = self._layout()
        cm = self.categorical_module
        cont = self.continuous_module
        seq_t = inputs[cm.item_id] if (cm is not None and cm.item_id) else max(
            (inputs[n] for n, kind, *_ in layout if kind != "dense"), key=lambda t: t.dim())
        B, L = seq_t.shape[0], seq_t.shape[1]
        M = B * L
        if cm is not None and cm.item_id:
            cm.item_seq = inputs[cm.item_id]  # features/embedding.py:244-245 (side channel for the head)

        if not (self.masking or self.projection_module or self.aggregation):
            # no aggregation requested: behave like MergeTabular (dict of per-feature tensors)
            out = {}
            if cm is not None:
                out.update(cm(inputs))
            if cont is not None:
                out.update(cont(inputs))
            return out

        # 1. labels / mask first: the projection epilogue needs the row codes
        row_code = None
        mask_vec = None
        inference_mlm = False
        if self.masking:
            self.masking.compute_masked_targets(cm.item_seq, training=training, testing=testing)
            row_code = self.masking.row_code
            mask_vec = self.masking.masked_item_embedding.detach().float()
            inference_mlm = row_code.shape[1] != L  # MLM inference appends one position

        def is_context(v):
            # TabularAggregation._expand_non_sequential_features (tabular/base.py:53-63): context
            # features [B] are repeated for every position of the sequence
            return v.dim() == 1 or (v.dim() == 2 and v.shape[1] == 1 and L != 1)

        def seq(v):
            if is_context(v):
                v = v.reshape(B, 1).expand(B, L)
            return v.reshape(-1)

        agg = self.AGGREGATIONS[self.aggregation or "concat"]
        cat_ln = getattr(cm, "post", None) if cm is not None else None
        cont_ln = getattr(cont, "post", None) if cont is not None else None
        sharded = [n for n, kind, *_ in layout if kind == "cat" and cm.is_sharded(n)]
        plain = agg == _lib.AGG_CONCAT and all(kind in ("cat", "cont") for _, kind, *_ in layout) and not (
            cat_ln is not None and len(cat_ln.feature_layer_norm) > 0) and not sharded

        def gather(want_f32: bool, want_planes: bool):
            if plain:  # the specialised HBM-roofline gather
                cats, conts = [], []
                for name, kind, col, width in layout:
                    if kind == "cat":
                        cats.append((cm.embedding_tables[name].weight.detach(), seq(inputs[name]), col))
                    else:
                        conts.append((seq(inputs[name]), col))
                return ops.embed_concat(cats, conts, M, C_width, want_f32=want_f32, want_planes=want_planes)
            feats = []
            item_feature = -1
            if sharded and len(layout) == 1 and agg == _lib.AGG_CONCAT and not (cat_ln is not None and cat_ln.params(sharded[0])):
                # item-only input block over a row-sharded table (config 4): the exchange's second gather
                # already emits the rows in session order, fp32 and split planes
                rows, planes = cm.embedding_tables[sharded[0]].lookup(seq(inputs[sharded[0]]))
                return (rows if want_f32 else None), (planes if want_planes else None), None
            for i, (name, kind, col, width) in enumerate(layout):
                if kind == "cat":
                    v = inputs[name]
                    ln = cat_ln.params(name) if cat_ln is not None else None
                    if name in sharded:  # rows arrive through the all-to-all, then enter as a dense feature
                        rows, _ = cm.embedding_tables[name].lookup(v)
                        f = dict(kind=_lib.FEAT_DENSE, dim=width, col=col, input=rows, per_session=is_context(v), ln=ln)
                    else:
                        f = dict(kind=_lib.FEAT_CAT, dim=width, col=col, input=v, per_session=is_context(v),
                                 table=cm.embedding_tables[name].weight, ln=ln)
                    if name == cm.item_id:
                        item_feature = i
                elif kind == "cont":
                    v = inputs[name]
                    f = dict(kind=_lib.FEAT_CONT, dim=1, col=col, input=v, per_session=is_context(v))
                elif kind == "soft":
                    v = inputs[name]
                    f = cont.embedding_tables[name].feature(v, col=col, per_session=is_context(v),
                                                            ln=cont_ln.params(name) if cont_ln is not None else None)
                else:  # continuous_projection: its MLP runs first, the result enters as a dense feature
                    f = dict(kind=_lib.FEAT_DENSE, dim=width, col=col, input=cont.project(inputs, seq))
                feats.append(f)
            if agg == _lib.AGG_SUM_ITEM_MULTI and item_feature < 0:
                raise ValueError("element-wise-sum-item-multi needs the item-id feature")
            return ops.input_block(feats, M, L, C_width, agg=agg, item_feature=item_feature, want_f32=want_f32,
                                   want_planes=want_planes)

        check = os.environ.get("T4R_CHECK_IDS", "0") == "1"
        proj = self._projection_linear()
        if proj is not None:
            _, planes, self._id_err = gather(False, True)
            if check:
                self.check_ids()
            w_planes = self._planes.get("proj", proj.weight)
            fuse_mask = row_code is not None and not inference_mlm
            x, x_planes, _ = ops.linear(planes, w_planes, C_width, bias=proj.bias, act=self._projection_act(),
                                        row_code=row_code if fuse_mask else None,
                                        mask_vec=mask_vec if fuse_mask else None)
            x = x.view(B, L, -1)
            if row_code is not None and not fuse_mask:
                x = self.masking.apply_mask_to_inputs(x, self.masking.mask_schema, training=training, testing=testing)
            else:
                x._t4r_planes = x_planes
            return x

        concat, _, self._id_err = gather(True, False)
        if check:
            self.check_ids()
        x = concat.view(B, L, C_width)
        if self.projection_module is not None:
            x = self.projection_module(x)
        if self.masking:
            x = self.masking.apply_mask_to_inputs(x, self.masking.mask_schema, training=training, testing=testing)
        return x

    def check_ids(self) -> None:
        """``nn.Embedding`` raises on ids outside the table (features/embedding.py:226-249 -> ``IndexError: index out of
        range in self``).  The fused gather cannot raise from the device: it reads row 0 for such an id and raises a
        flag that this method turns into the reference's error (one host sync; also run on every forward when
        ``T4R_CHECK_IDS=1``)."""
        err = getattr(self, "_id_err", None)
        if err is not None and int(err.item()) != 0:
            raise IndexError("index out of range in self")

    def _projection_linear(self) -> Optional[nn.Linear]:
        """The fused path handles the reference's default projection: a single
        DenseBlock = Linear (+ReLU).  Anything else runs as a generic module."""
        pm = self.projection_module
        if pm is None:
            return None
        lin = getattr(pm, "fusable_linear", None)
        return lin() if callable(lin) else None

    def _projection_act(self) -> int:
        return self.projection_module.fusable_activation()
