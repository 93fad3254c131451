"""Input block: host-side mirror of ``TabularSequenceFeatures`` and its parts.

Reference: transformers4rec/torch/features/sequence.py:97-296 (orchestration),
features/embedding.py:51-257,416-514 (tables), features/continuous.py:60-63,
tabular/aggregation.py:35-47 (concat in sorted-name order), block/mlp.py:123-144
(projection).  ``forward`` issues three kernels for what the reference runs as
F gathers + cat + GEMM + ReLU + ~15 masking ops + where:
  t4r_mask_*  ->  t4r_embed_concat_fwd  ->  t4r_linear_fwd (bias+ReLU+mask epilogue).
"""
from __future__ import annotations

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


class SequenceEmbeddingFeatures(nn.Module):
    """features/sequence.py:43-90 + features/embedding.py:51-257: owns one
    ``nn.Embedding(V, dim, padding_idx)`` per categorical feature (state-dict keys
    ``embedding_tables.<feature>.weight``) and remembers ``item_seq``."""

    def __init__(self, feature_config: Dict[str, FeatureConfig], item_id: Optional[str] = None, padding_idx: int = 0):
        super().__init__()
        self.padding_idx = padding_idx
        self.item_id = item_id
        self.feature_config = feature_config
        tables = {}
        for name, feature in feature_config.items():
            emb = nn.Embedding(feature.table.vocabulary_size, feature.table.dim, padding_idx=padding_idx)
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
                    embeddings_initializers=None, combiner="mean", tags=None, item_id=None, padding_idx=0, **kwargs):
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
        return cls(feature_config, item_id=item_id, padding_idx=padding_idx)

    def item_ids(self, inputs) -> torch.Tensor:
        return inputs[self.item_id]

    def output_dims(self) -> Dict[str, int]:
        return {n: f.table.dim for n, f in self.feature_config.items()}

    def forward(self, inputs, **kwargs):
        """Stand-alone use (dict of [B, L, dim] tensors, like the reference); the fused
        path in TabularSequenceFeatures does not go through here."""
        out = {}
        for name in self.feature_config:
            ids = inputs[name]
            shp = tuple(ids.shape)
            table = self.embedding_tables[name].weight
            of, _, _ = ops.embed_concat([(table.detach(), ids.reshape(-1), 0)], [], ids.numel(), table.shape[1], True,
                                        False)
            out[name] = of.view(*shp, table.shape[1])
        if self.item_id:
            self.item_seq = self.item_ids(inputs)
        return out


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
    """features/sequence.py:97-296."""

    EMBEDDING_MODULE_CLASS = SequenceEmbeddingFeatures
    CONTINUOUS_MODULE_CLASS = ContinuousFeatures

    def __init__(self, continuous_module=None, categorical_module=None, pretrained_embedding_module=None,
                 projection_module=None, masking: Optional[MaskSequence] = None, aggregation: Optional[str] = None,
                 schema: Optional[Schema] = None, **kwargs):
        super().__init__()
        if pretrained_embedding_module is not None:
            raise NotImplementedError("pretrained embeddings are outside the t4r_b200 hot path (SURVEY §2 row 1)")
        to_merge = {}
        if continuous_module is not None:
            to_merge["continuous_module"] = continuous_module
        if categorical_module is not None:
            to_merge["categorical_module"] = categorical_module
        assert to_merge != {}, "Please provide at least one input layer"
        self.to_merge = nn.ModuleDict(to_merge)
        if aggregation not in (None, "concat"):
            raise NotImplementedError(f"aggregation '{aggregation}' is not on the t4r_b200 hot path (concat only)")
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
        """features/sequence.py:140-229."""
        if continuous_projection:
            raise NotImplementedError("continuous_projection is not on the t4r_b200 hot path yet (SURVEY §8f)")
        if continuous_soft_embeddings:
            raise NotImplementedError("soft embeddings are outside the t4r_b200 hot path (SURVEY §2 row 1)")
        cont = cls.CONTINUOUS_MODULE_CLASS.from_schema(schema, tags=continuous_tags) if continuous_tags else None
        emb_kwargs = {k: v for k, v in kwargs.items() if k in (
            "embedding_dims", "embedding_dim_default", "infer_embedding_sizes", "infer_embedding_sizes_multiplier",
            "embeddings_initializers", "combiner", "item_id", "padding_idx")}
        cat = cls.EMBEDDING_MODULE_CLASS.from_schema(schema, tags=categorical_tags, **emb_kwargs) if categorical_tags else None
        output = cls(continuous_module=cont, categorical_module=cat, aggregation=aggregation, schema=schema)
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
        """(name, kind, first column, width) in sorted-name order (aggregation.py:42-47)."""
        widths = {}
        kinds = {}
        if self.categorical_module is not None:
            for n, dim in self.categorical_module.output_dims().items():
                widths[n], kinds[n] = dim, "cat"
        if self.continuous_module is not None:
            for n in self.continuous_module.features:
                widths[n], kinds[n] = 1, "cont"
        col = 0
        out = []
        for n in sorted(widths.keys()):
            out.append((n, kinds[n], col, widths[n]))
            col += widths[n]
        return out, col

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
        layout, C_width = self._layout()
        cm = self.categorical_module
        seq_t = inputs[cm.item_id] if (cm is not None and cm.item_id) else max(
            (inputs[n] for n, *_ in layout), key=lambda t: t.dim())
        B, L = seq_t.shape[0], seq_t.shape[1]
        M = B * L
        if cm is not None and cm.item_id:
            cm.item_seq = inputs[cm.item_id]  # features/embedding.py:244-245 (side channel for the head)

        if not (self.masking or self.projection_module):
            # no aggregation requested: behave like MergeTabular (dict of per-feature tensors)
            out = {}
            if cm is not None:
                out.update(cm(inputs))
            if self.continuous_module is not None:
                out.update(self.continuous_module(inputs))
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

        def seq(v):
            # TabularAggregation._expand_non_sequential_features (tabular/base.py:53-63): context
            # features [B] are repeated for every position of the sequence
            if v.dim() == 1 or (v.dim() == 2 and v.shape[1] == 1 and L != 1):
                v = v.reshape(B, 1).expand(B, L)
            return v.reshape(-1)

        cats, conts = [], []
        for name, kind, col, width in layout:
            if kind == "cat":
                cats.append((cm.embedding_tables[name].weight.detach(), seq(inputs[name]), col))
            else:
                conts.append((seq(inputs[name]), col))

        proj = self._projection_linear()
        if proj is not None:
            _, planes, self._id_err = ops.embed_concat(cats, conts, M, C_width, want_f32=False, want_planes=True)
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

        concat, _, self._id_err = ops.embed_concat(cats, conts, M, C_width, want_f32=True, want_planes=False)
        x = concat.view(B, L, C_width)
        if self.projection_module is not None:
            x = self.projection_module(x)
        if self.masking:
            x = self.masking.apply_mask_to_inputs(x, self.masking.mask_schema, training=training, testing=testing)
        return x

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
