"""Streaming ranking metrics computed from the label's rank.

Reference: transformers4rec/torch/ranking_metric.py:30-147 and
utils/torch_utils.py:226-238 one-hot the labels to [T, V] and run ``torch.topk``;
for a single relevant item per row Recall@k is just ``rank < k``, where ``rank`` =
number of classes scoring above the label (ties: lower id first, matching the
order ``torch.topk`` yields on the reference's known-answer tests).  The rank is
produced inside the fused head kernel, so [T, V] is never materialised.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _lib, ops


class RankingMetric:
    def __init__(self, top_ks=None, labels_onehot=False):
        self.top_ks = top_ks or [2, 5]
        if not isinstance(self.top_ks, (list, tuple)):
            self.top_ks = [self.top_ks]
        self.labels_onehot = labels_onehot
        self.metric_mean: List[torch.Tensor] = []
        self.metric_rows: List[torch.Tensor] = []

    def reset(self):
        self.metric_mean = []
        self.metric_rows = []

    def update_from_ranks(self, row_rank: torch.Tensor, t_dev: Optional[torch.Tensor] = None):
        m = self._from_ranks(row_rank, t_dev)
        self.metric_mean.append(m)
        # rows behind this batch's mean (device-side count when the caller has one: no host sync here)
        rows = t_dev.reshape(-1)[:1].to(m.dtype) if t_dev is not None else m.new_full((1,), float(row_rank.numel()))
        self.metric_rows.append(rows.to(m.device))

    def update(self, preds: torch.Tensor, target: torch.Tensor, **kwargs):
        """Materialised scores [T, V] + class-id labels [T]."""
        if not self.labels_onehot:
            target = target.view(-1, target.size(-1)).float().argmax(dim=-1)
        preds = preds.reshape(-1, preds.size(-1)).float()
        target = target.reshape(-1).long()
        tgt = preds.gather(1, target.unsqueeze(1))
        ids = torch.arange(preds.size(1), device=preds.device).unsqueeze(0)
        rank = ((preds > tgt) | ((preds == tgt) & (ids < target.unsqueeze(1)))).sum(dim=1).int()
        self.update_from_ranks(rank)

    def __call__(self, preds, target, **kwargs):
        self.update(preds, target, **kwargs)
        return self.metric_mean[-1]

    def compute(self):
        # ranking_metric.py:61-63 concatenates the per-ROW results of every update and takes their mean: a
        # row-weighted mean of the per-batch means (equal to the plain mean only when every batch has the same T)
        means = torch.stack(self.metric_mean, dim=0)
        rows = torch.cat(self.metric_rows).reshape(-1, 1)
        return (means * rows).sum(0) / rows.sum().clamp(min=1.0)

    def _from_ranks(self, row_rank, t_dev):
        raise NotImplementedError


class RecallAt(RankingMetric):
    """ranking_metric.py:106-147."""

    def _from_ranks(self, row_rank, t_dev):
        outs = []
        ks = list(self.top_ks)
        for i in range(0, len(ks), 4):
            outs.append(ops.recall_from_ranks(row_rank, ks[i:i + 4], t_dev))
        return torch.cat(outs)


class PrecisionAt(RankingMetric):
    """ranking_metric.py:73-103 with one relevant item per row: [rank < k] / k."""

    def _from_ranks(self, row_rank, t_dev):
        return ops.metrics_from_ranks(row_rank, self.top_ks, _lib.METRIC_PRECISION, t_dev)


class AvgPrecisionAt(RankingMetric):
    """ranking_metric.py:150-190 with one relevant item per row: [rank < k] / (rank + 1)."""

    def _from_ranks(self, row_rank, t_dev):
        return ops.metrics_from_ranks(row_rank, self.top_ks, _lib.METRIC_RR, t_dev)


class MeanReciprocalRankAt(AvgPrecisionAt):
    """ranking_metric.py:284-319: identical to AvgPrecisionAt when a row has a single relevant item."""


class DCGAt(RankingMetric):
    """ranking_metric.py:193-238 with one relevant item per row: [rank < k] / log2(rank + 2)."""

    def _from_ranks(self, row_rank, t_dev):
        return ops.metrics_from_ranks(row_rank, self.top_ks, _lib.METRIC_DCG, t_dev)


class NDCGAt(DCGAt):
    """ranking_metric.py:241-281: the ideal DCG of a single relevant item is 1, so NDCG == DCG."""


ranking_metrics_registry = {
    "precision_at": PrecisionAt, "precision": PrecisionAt,
    "recall_at": RecallAt, "recall": RecallAt,
    "avg_precision_at": AvgPrecisionAt, "avg_precision": AvgPrecisionAt, "map": AvgPrecisionAt,
    "dcg_at": DCGAt, "dcg": DCGAt,
    "ndcg_at": NDCGAt, "ndcg": NDCGAt,
    "mrr_at": MeanReciprocalRankAt, "mrr": MeanReciprocalRankAt,
}
