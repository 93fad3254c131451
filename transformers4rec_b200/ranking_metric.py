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

from . import ops


class RankingMetric:
    def __init__(self, top_ks=None, labels_onehot=False):
        self.top_ks = top_ks or [2, 5]
        if not isinstance(self.top_ks, (list, tuple)):
            self.top_ks = [self.top_ks]
        self.labels_onehot = labels_onehot
        self.metric_mean: List[torch.Tensor] = []

    def reset(self):
        self.metric_mean = []

    def update_from_ranks(self, row_rank: torch.Tensor, t_dev: Optional[torch.Tensor] = None):
        self.metric_mean.append(self._from_ranks(row_rank, t_dev))

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
        # ranking_metric.py:61-63: mean over batches of the per-batch means
        return torch.stack(self.metric_mean, dim=0).mean(0)

    def _from_ranks(self, row_rank, t_dev):
        raise NotImplementedError


class RecallAt(RankingMetric):
    """ranking_metric.py:111-147."""

    def _from_ranks(self, row_rank, t_dev):
        outs = []
        ks = list(self.top_ks)
        for i in range(0, len(ks), 4):
            outs.append(ops.recall_from_ranks(row_rank, ks[i:i + 4], t_dev))
        return torch.cat(outs)


class NDCGAt(RankingMetric):
    """ranking_metric.py:247-319 specialised to one relevant item: 1/log2(rank+2) if rank < k."""

    def _from_ranks(self, row_rank, t_dev):
        r = row_rank.float()
        n = row_rank.numel() if t_dev is None else None
        valid = torch.ones_like(r, dtype=torch.bool) if t_dev is None else (
            torch.arange(r.numel(), device=r.device) < t_dev)
        gain = 1.0 / torch.log2(r + 2.0)
        cnt = valid.sum().clamp(min=1)
        return torch.stack([(gain * ((r < k) & valid)).sum() / cnt for k in self.top_ks])


class AvgPrecisionAt(RankingMetric):
    """ranking_metric.py:150-196 specialised to one relevant item: 1/(rank+1) if rank < k."""

    def _from_ranks(self, row_rank, t_dev):
        r = row_rank.float()
        valid = torch.ones_like(r, dtype=torch.bool) if t_dev is None else (
            torch.arange(r.numel(), device=r.device) < t_dev)
        cnt = valid.sum().clamp(min=1)
        return torch.stack([((1.0 / (r + 1.0)) * ((r < k) & valid)).sum() / cnt for k in self.top_ks])


MeanReciprocalRankAt = AvgPrecisionAt  # identical for a single relevant item
