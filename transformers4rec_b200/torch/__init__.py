"""Drop-in namespace: ``import transformers4rec_b200.torch as tr`` exposes the names
``transformers4rec.torch`` exports for the session-transformer path
(reference: transformers4rec/torch/__init__.py:77-132)."""
from ..block import (DenseBlock, GPT2Encoder, MLPBlock, SequentialBlock, TransformerBlock,  # noqa: F401
                     XLNetEncoder)
from ..config import GPT2Config, T4RecConfig, XLNetConfig, transformer_registry  # noqa: F401
from ..features import (ContinuousFeatures, ContinuousProjection, FeatureConfig, SequenceEmbeddingFeatures,  # noqa: F401
                        SoftEmbedding, SoftEmbeddingFeatures, StochasticSwapNoise, TableConfig, TabularLayerNorm,
                        TabularSequenceFeatures)
from ..masking import (CausalLanguageModeling, MaskedLanguageModeling, MaskSequence,  # noqa: F401
                       PermutationLanguageModeling, masking_registry)
from ..model import GraphedForward, Head, Model  # noqa: F401
from ..prediction_task import (LogUniformSampler, NextItemPredictionTask, PredictionTask)  # noqa: F401
from ..ranking_metric import (AvgPrecisionAt, DCGAt, MeanReciprocalRankAt, NDCGAt, PrecisionAt, RecallAt,  # noqa: F401
                              ranking_metrics_registry)
from ..schema import ColumnSchema, Schema, Tags  # noqa: F401
from ..padding import pad_batch, pad_inputs  # noqa: F401

from ..training import FusedAdamW, FusedTrainingStep, training_loss  # noqa: F401,E402  (N3: backward + optimizer step)
