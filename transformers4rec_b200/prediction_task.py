"""Next-item head: host-side mirror of ``NextItemPredictionTask`` and friends.

Reference: transformers4rec/torch/model/prediction_task.py:306-512 (task),
:589-699 (``_NextItemPredictionTask``: tied logits, temperature, sampled softmax),
:702-861 (``LogUniformSampler``); model/base.py:35-150 (``PredictionTask``).
The arithmetic runs in ``t4r_compact_targets`` -> ``t4r_gather_rows_split`` ->
(``t4r_linear_fwd`` for the task_block) -> ``t4r_head_softmax_ce_fwd``: logits are
reduced to (log-sum-exp, label logit, label rank) tile by tile and [T, V] is never
written.  ``predictions`` stays available lazily (``t4r_head_logits``).
"""
from __future__ import annotations

import math

import logging
from typing import Dict, Iterable, Optional

import torch
from torch import nn

from . import _lib, ops
from .block import MLPBlock
from .masking import MaskedLanguageModeling
from .ranking_metric import AvgPrecisionAt, NDCGAt, RecallAt

LOG = logging.getLogger("transformers4rec_b200")


class LazyOutputs(dict):
    """The reference returns ``{"loss", "labels", "predictions"}``.  ``loss`` is
    computed eagerly (asynchronously); ``labels`` needs the label count on the host
    and ``predictions`` materialises [T, V], so both are produced on first access."""

    def __init__(self, eager: Dict, lazy: Dict):
        super().__init__(eager)
        self._lazy = lazy

    def __getitem__(self, k):
        if not dict.__contains__(self, k) and k in self._lazy:
            dict.__setitem__(self, k, self._lazy[k]())
        return dict.__getitem__(self, k)

    def __contains__(self, k):
        return dict.__contains__(self, k) or k in self._lazy

    def get(self, k, default=None):
        return self[k] if k in self else default

    def keys(self):
        return list(dict.keys(self)) + [k for k in self._lazy if not dict.__contains__(self, k)]


class PredictionTask(nn.Module):
    """model/base.py:35-150 (the parts the next-item path uses)."""

    def __init__(self, loss=None, metrics=None, task_block=None, task_name=None):
        super().__init__()
        self.loss = loss
        self.metrics = list(metrics) if metrics is not None else []
        self.task_block = task_block
        self.task_name = task_name
        self.pre = None

    def build(self, body, input_size, inputs=None, device=None, task_block=None, pre=None):
        if task_block is not None:
            if hasattr(task_block, "build") and not isinstance(task_block, nn.Module):
                task_block = task_block.build(input_size)
            self.task_block = task_block
        self.input_size = input_size
        if device is not None:
            self.to(device)
        return self

    def to_head(self, body, inputs=None, **kwargs):
        from .model import Head
        return Head(body, self, inputs=inputs, **kwargs)

    def to_model(self, body, inputs=None, **kwargs):
        from .model import Head, Model
        return Model(Head(body, self, inputs=inputs, **kwargs), **kwargs)

    def metric_name(self, metric) -> str:
        name = type(metric).__name__
        snake = "".join(["_" + c.lower() if c.isupper() else c for c in name]).lstrip("_")
        return f"{self.task_name}/{snake}" if self.task_name else snake

    def reset_metrics(self):
        for m in self.metrics:
            m.reset()


class LogUniformSampler(nn.Module):
    """model/prediction_task.py:702-861.  The draw itself (``multinomial`` over V
    entries, ``unique``, truncate) is torch plumbing on the device, as in the
    reference; tests inject the raw draws to compare bit-exactly."""

    def __init__(self, max_n_samples: int, max_id: int, min_id: int = 0, unique_sampling: bool = True,
                 n_samples_multiplier_before_unique: int = 2):
        super().__init__()
        if max_id <= 0:
            raise ValueError("max_id must be a positive integer.")
        if max_n_samples <= 0:
            raise ValueError("n_sample must be a positive integer.")
        self.max_id = max_id
        self.min_id = min_id
        self.unique_sampling = unique_sampling
        self.max_n_samples = max_n_samples
        self.n_sample = max_n_samples
        if self.unique_sampling:
            self.n_sample = int(self.n_sample * n_samples_multiplier_before_unique)
        with torch.no_grad():
            dist = self.get_log_uniform_distr(max_id, min_id)
            self.register_buffer("dist", dist)
            unique_sampling_dist = self.get_unique_sampling_distr(dist, self.n_sample)
            self.register_buffer("unique_sampling_dist", unique_sampling_dist)
            eff = unique_sampling_dist if unique_sampling else dist
            # -log(q + 1e-16): the logQ correction of :684-687, precomputed per class
            self.register_buffer("neg_log_q", -torch.log(eff + 1e-16))

    def get_log_uniform_distr(self, max_id: int, min_id: int = 0) -> torch.Tensor:
        log_indices = torch.arange(1.0, max_id - min_id + 2.0, 1.0).log_()
        probs = (log_indices[1:] - log_indices[:-1]) / log_indices[-1]
        if min_id > 0:
            probs = torch.cat([torch.zeros([min_id], dtype=probs.dtype), probs], dim=0)
        return probs

    def get_unique_sampling_distr(self, dist, n_sample):
        return (-(-dist.double().log1p_() * n_sample).expm1_()).float()

    def draw(self) -> torch.Tensor:
        if self.dist.numel() <= (1 << 20):
            return torch.multinomial(self.dist, self.n_sample, replacement=True)
        # torch.multinomial stops at 2^24 categories (the reference's sampler cannot run BASELINE config 5's
        # 50 M items at all) and costs milliseconds per draw long before that (measured: 4 ms of a 10.8 ms config-5
        # step at 6.25 M categories).  Same distribution by inverting its CDF, log(k + 2) / log(R + 1) for the k-th
        # id of the range (:766-787): k = floor(exp(u * log(R + 1))) - 1, in fp64.
        R = self.max_id - self.min_id
        u = torch.rand(self.n_sample, dtype=torch.float64, device=self.dist.device)
        k = torch.exp(u * math.log(R + 1.0)).floor().long().sub_(1).clamp_(0, R - 1)
        return k + self.min_id

    def sample(self, labels: torch.Tensor, raw_draws: Optional[torch.Tensor] = None):
        if not torch.is_tensor(labels):
            raise TypeError("Labels must be a torch.Tensor.")
        if labels.dtype != torch.long:
            raise ValueError("Labels must be a tensor of dtype long.")
        with torch.no_grad():
            raw = self.draw() if raw_draws is None else raw_draws
            neg_samples = raw.unique()[: self.max_n_samples].to(labels.device)
            dist = self.unique_sampling_dist if self.unique_sampling else self.dist
            return neg_samples, dist[labels], dist[neg_samples]

    def forward(self, labels):
        return self.sample(labels)


class NextItemPredictionTask(PredictionTask):
    """model/prediction_task.py:306-512."""

    def __init__(self, loss: nn.Module = None, metrics: Iterable = None, task_block=None, task_name: str = "next-item",
                 weight_tying: bool = False, softmax_temperature: float = 1, padding_idx: int = 0,
                 target_dim: int = None, sampled_softmax: Optional[bool] = False, max_n_samples: Optional[int] = 100):
        loss = loss if loss is not None else nn.CrossEntropyLoss()
        if not isinstance(loss, nn.CrossEntropyLoss) or loss.reduction != "mean" or loss.weight is not None:
            raise NotImplementedError("the fused head implements nn.CrossEntropyLoss(label_smoothing=...) with mean "
                                      "reduction and no class weights")
        label_smoothing = float(getattr(loss, "label_smoothing", 0.0))
        if label_smoothing and sampled_softmax:
            raise NotImplementedError("label smoothing with sampled softmax is not on the t4r_b200 hot path")
        if metrics is None:
            metrics = (NDCGAt(top_ks=[10, 20], labels_onehot=True), AvgPrecisionAt(top_ks=[10, 20], labels_onehot=True),
                       RecallAt(top_ks=[10, 20], labels_onehot=True))
        super().__init__(loss=loss, metrics=metrics, task_block=task_block, task_name=task_name)
        self.label_smoothing = label_smoothing
        self.softmax_temperature = softmax_temperature
        self.weight_tying = weight_tying
        self.padding_idx = padding_idx
        self.target_dim = target_dim
        self.sampled_softmax = sampled_softmax
        self.max_n_samples = max_n_samples
        self.item_embedding_table = None
        self.masking = None
        self.output_layer = None
        self.sampler = None
        # tensor-core arithmetic of the GEMMs on this task: 3 = split-bf16, three products per MAC; 1 = plain bf16;
        # 2 (default) = the 2-unit product in the TRAINING full-softmax head, replicated or row-sharded (fp16 x fp16 +
        # two e4m3 cross terms, csrc/t4r_mixed_pack.cuh) -- every other GEMM of the task (task_block, sampled head,
        # evaluation ranks, predictions, serving) runs with 3.  Measured on a B200 against fp64 (tools/precision_gpu.py,
        # profiles/r2_head_precision.json; 2048 rows x 200 k classes x 256): max row-loss error 7.3e-6 (2) vs 5.9e-6
        # (3) vs 3.1e-6 for torch's own fp32 sgemm at the reference's initialisation scale, 3.6e-4 vs 1.4e-4 at
        # |logit| <= 52, identical mean-loss error in every regime -- two orders inside the 1e-3 parity bar where
        # the bar is meaningful, at 2/3 of the tensor time (config-2 head 4.63 -> 3.76 ms).
        self.nprod = 2
        self._planes = ops.PlaneCache()
        self._last = None
        self._neg_draws = None

    def build(self, body, input_size, device=None, inputs=None, task_block=None, pre=None):
        """model/prediction_task.py:369-417."""
        if not len(input_size) == 3 or isinstance(input_size, dict):
            raise ValueError(f"NextItemPredictionTask needs a 3-dim vector as input, found:{input_size}")
        if not inputs:
            inputs = body.inputs
        if not getattr(inputs, "item_id", None):
            raise ValueError("For Item Prediction task a categorical_module including an item_id column is required.")
        self.embeddings = inputs.categorical_module
        if not self.target_dim:
            self.target_dim = self.embeddings.item_embedding_table.num_embeddings
        task_block = task_block or self.task_block
        if self.weight_tying:
            self.item_embedding_table = self.embeddings.item_embedding_table
            item_dim = self.item_embedding_table.weight.shape[1]
            if input_size[-1] != item_dim and not task_block:
                LOG.warning(f"Projecting inputs of NextItemPredictionTask to'{item_dim}' As weight tying requires the "
                            f"input dimension '{input_size[-1]}' to be equal to the item-id embedding dimension '{item_dim}'")
                task_block = MLPBlock([item_dim], activation=None)
        self.masking = inputs.masking
        if not self.masking:
            raise ValueError("The input block should contain a masking schema for training and evaluation")
        self.padding_idx = self.masking.padding_idx
        super().build(body, input_size, device=device, inputs=inputs, task_block=task_block)
        head_in = self.task_block.output_size()[-1] if self.task_block is not None else input_size[-1]
        if not self.weight_tying:
            # :636-638 own output layer, kaiming_uniform(a=sqrt(5))
            import math
            self.output_layer = nn.Parameter(torch.empty(self.target_dim, head_in))
            torch.nn.init.kaiming_uniform_(self.output_layer, a=math.sqrt(5))
        if self.sampled_softmax:
            self.sampler = LogUniformSampler(max_n_samples=self.max_n_samples, max_id=self.target_dim,
                                             min_id=self.padding_idx + 1, unique_sampling=True)
        if device is not None:
            self.to(device)
        return self

    # ------------------------------------------------------------------ helpers
    def output_weight(self) -> torch.Tensor:
        if self.weight_tying:
            self.item_embedding_table = self.embeddings.item_embedding_table  # may have been swapped for its shard
            return self.item_embedding_table.weight
        return self.output_layer

    def set_negative_draws(self, raw_draws: Optional[torch.Tensor]):
        """Test hook: the multinomial output ids the sampler would have drawn."""
        self._neg_draws = raw_draws

    def _dense_nprod(self) -> int:
        return 3 if self.nprod == 2 else self.nprod

    def _inv_tau(self) -> float:
        return 1.0 / float(self.softmax_temperature) if self.softmax_temperature else 1.0

    def _task_block_rows(self, planes, m_dev):
        """task_block (Linear d -> item_dim, no activation by default) on compacted rows."""
        blocks = list(self.task_block)
        xf = None
        K = None
        for blk in blocks:
            lin = blk[0]
            K = lin.in_features
            xf, planes, _ = ops.linear(planes, blk._planes.get("w", lin.weight), K, bias=lin.bias, act=blk.act_code(),
                                       m_dev=m_dev, nprod=self._dense_nprod())
        return xf, planes

    # ------------------------------------------------------------------ forward
    def forward(self, inputs: torch.Tensor, targets=None, training=False, testing=False, top_k=None, **kwargs):
        if isinstance(inputs, (tuple, list)):
            inputs = inputs[0]
        x = inputs
        B, L, d = x.shape
        W = self.output_weight()
        Wd = W.detach()
        mixed_ok = self.nprod == 2 and training and not self.sampled_softmax
        mixed_head = mixed_ok and not self._sharded()
        if mixed_head:
            w_planes = None                                # the mixed head keeps only its own 1x copy of the table
        elif mixed_ok:
            w_planes = self._planes.get_mixed("W", W)      # sharded full softmax: (planes, inverse scales) of the shard
        else:
            w_planes = self._planes.get("W", W)
        inv_tau = self._inv_tau()

        if training or testing:
            labels_2d = self.masking.masked_targets
            tgt_rows, tgt_labels, count = ops.compact_targets(labels_2d, self.padding_idx)
            cap = B * L
            want_rank = bool(testing and not training)
            peer_head = None
            if self._sharded() and not (self.sampled_softmax and training):
                # over NVLink peer memory the label rows go straight into the window the other ranks pull from
                peer_head = self._peer_head(cap, Wd.shape[1], x.device)
            direct = peer_head is not None and self.task_block is None and d == Wd.shape[1]
            xt_planes, xt_f32 = ops.gather_rows_split(x.reshape(B * L, d), tgt_rows, count, cap, want_f32=True,
                                                      **(dict(out_f32=peer_head.mail_x) if direct else {}))
            if self.task_block is not None:
                xt_f32, xt_planes = self._task_block_rows(xt_planes, count)
            if self._sharded():
                if peer_head is not None or (self.sampled_softmax and training and self.item_embedding_table.peer_view() is not None):
                    return self._forward_sharded_peer(peer_head, xt_planes, xt_f32, tgt_labels, count, training, want_rank,
                                                      w_planes, inv_tau)
                return self._forward_sharded(xt_planes, xt_f32, tgt_labels, count, training, want_rank, w_planes, inv_tau)
            if self.sampled_softmax and training:
                neg, _, _ = self.sampler.sample(tgt_labels[:1], raw_draws=self._neg_draws)
                S = neg.numel()
                neg_planes, _ = ops.gather_rows_split(Wd, neg, None, S, want_f32=False)
                col_bias = self.sampler.neg_log_q[neg].contiguous()
                pos = ops.label_logit(xt_f32, Wd, tgt_labels, t_dev=count, class_bias=self.sampler.neg_log_q,
                                      inv_temperature=inv_tau)
                res = ops.head_softmax_ce(xt_planes, xt_f32, tgt_labels, neg_planes, None, t_dev=count,
                                          inv_temperature=inv_tau, col_bias=col_bias, col_ids=neg, col_ids_sorted_unique=True,
                                          hit_value=float(torch.finfo(torch.float16).min / 100.0), pos_logit=pos,
                                          nprod=self._dense_nprod())
                self._last = dict(xt_planes=xt_planes, w_planes=neg_planes, count=count, neg=neg, pos=pos,
                                  col_bias=col_bias, labels=tgt_labels, De=Wd.shape[1], sampled=True)
            elif mixed_head:
                w_mixed, w_inv = self._planes.get_mixed("W", W)
                xt_mixed, xt_inv = ops.split_planes_mixed(xt_f32, count=count)
                res = ops.head_softmax_ce(xt_mixed, xt_f32, tgt_labels, w_mixed, Wd, t_dev=count,
                                          inv_temperature=inv_tau, want_rank=want_rank, nprod=2,
                                          label_smoothing=self.label_smoothing, xt_inv_scale=xt_inv, w_inv_scale=w_inv)
                self._last = dict(xt_planes=xt_planes, w_planes=None, count=count, labels=tgt_labels,
                                  De=Wd.shape[1], sampled=False)
            else:
                res = ops.head_softmax_ce(xt_planes, xt_f32, tgt_labels, w_planes, Wd, t_dev=count,
                                          inv_temperature=inv_tau, want_rank=want_rank, nprod=self._dense_nprod(),
                                          label_smoothing=self.label_smoothing)
                self._last = dict(xt_planes=xt_planes, w_planes=w_planes, count=count, labels=tgt_labels,
                                  De=Wd.shape[1], sampled=False)
            self._last.update(res)
            loss = res["loss"].reshape(())
            out = LazyOutputs({"loss": loss}, {"labels": self._lazy_labels, "predictions": self._lazy_predictions})
            out.row_rank, out.count = res["row_rank"], count
            return out

        # inference (:452-470): hidden state at the next-item position, full scores
        item_seq = self.embeddings.item_seq
        non_pad = item_seq != self.padding_idx
        rows_ids = torch.arange(item_seq.size(0), dtype=torch.long, device=item_seq.device)
        last = non_pad.sum(dim=1) if isinstance(self.masking, MaskedLanguageModeling) else non_pad.sum(dim=1) - 1
        flat_idx = (rows_ids * x.shape[1] + last).int()
        xs_planes, xs_f32 = ops.gather_rows_split(x.reshape(-1, d), flat_idx, None, B, want_f32=True)
        if self.task_block is not None:
            xs_f32, xs_planes = self._task_block_rows(xs_planes, None)
        if self._sharded():
            # serving over the row-sharded table: per-shard top-k + one exchange of (score, id) candidates
            if top_k is None:
                raise NotImplementedError("scores [B, V] are not materialised over a row-sharded item table: "
                                          "pass top_k (Model.top_k / forward(..., top_k=k))")
            from . import distributed as D
            table = self.item_embedding_table
            return D.sharded_topk(xs_f32, table.weight.detach(), table.num_embeddings, int(top_k), table.group,
                                  w_planes=w_planes, inv_tau=inv_tau)
        scores = ops.head_logits(xs_planes, w_planes, Wd.shape[1], inv_temperature=inv_tau, nprod=self._dense_nprod())
        if top_k is None:
            return scores
        return ops.topk(scores, top_k)

    # ------------------------------------------------------------------ row-sharded table (configs 4-5)
    def _sharded(self) -> bool:
        return self.weight_tying and hasattr(self.item_embedding_table, "lookup")

    def _forward_sharded(self, xt_planes, xt_f32, tgt_labels, count, training, want_rank, w_planes, inv_tau):
        """the label rows of all ranks against every rank's rows of the tied table.  Full softmax:
        all-gather of the label rows + one all-gather of per-row (lse, label-logit) pairs.  Sampled softmax:
        identical negatives on every rank (rank 0's draws are broadcast), their rows and the positives' rows
        arrive through the table's row exchange, after which the head is data parallel."""
        import torch.distributed as dist

        from . import distributed as D
        table = self.item_embedding_table
        Wd = table.weight.detach()
        if self.label_smoothing:
            raise NotImplementedError("label smoothing over a row-sharded table is not on the t4r_b200 path")
        T = int(count.item())  # the collectives need host-side sizes
        xt = xt_f32[:T].contiguous()
        y = tgt_labels[:T].contiguous()
        if self.sampled_softmax and training:
            if self._neg_draws is not None:
                raw = self._neg_draws
            else:
                raw = self.sampler.draw()
                dist.broadcast(raw, src=dist.get_global_rank(table.group, 0) if table.group is not None else 0,
                               group=table.group)
            neg, _, _ = self.sampler.sample(y[:1], raw_draws=raw)
            S = neg.numel()
            _, neg_planes = table.lookup(neg)
            wy, _ = table.lookup(y, ragged=True)
            col_bias = self.sampler.neg_log_q[neg].contiguous()
            pos = ops.label_logit(xt, wy, torch.arange(T, device=xt.device), class_bias=self.sampler.neg_log_q[y].contiguous(),
                                  inv_temperature=inv_tau)
            xp = xt_planes[:, :T].contiguous()
            res = ops.head_softmax_ce(xp, xt, y, neg_planes, None, inv_temperature=inv_tau, col_bias=col_bias, col_ids=neg, col_ids_sorted_unique=True,
                                      hit_value=float(torch.finfo(torch.float16).min / 100.0), pos_logit=pos,
                                      nprod=self._dense_nprod())
            tot = torch.stack([res["row_loss"][:T].sum(), torch.tensor(float(T), device=xt.device)])
            dist.all_reduce(tot, group=table.group)
            loss = (tot[0] / tot[1].clamp(min=1.0)).reshape(())
            self._last = dict(count=count, labels=tgt_labels, sampled=True, sharded=True, S=S)
            out = LazyOutputs({"loss": loss}, {"labels": self._lazy_labels, "predictions": self._no_sharded_predictions})
            out.row_rank, out.count = None, count
            return out
        res = D.sharded_softmax_ce(xt, y, Wd, table.num_embeddings, table.group, w_planes=w_planes, inv_tau=inv_tau,
                                   want_rank=want_rank)
        loss = res[1].reshape(())
        self._last = dict(count=count, labels=tgt_labels, sampled=False, sharded=True)
        out = LazyOutputs({"loss": loss}, {"labels": self._lazy_labels, "predictions": self._no_sharded_predictions})
        out.row_rank = res[3] if want_rank else None
        out.count = None if want_rank else count  # ranks are already trimmed to this rank's T rows
        return out

    def _peer_head(self, cap: int, De: int, device):
        """The head's peer-memory windows for this capacity (created once: a collective), or None when peer memory
        is not available (CPU / gloo, T4R_PEER=0, failed mapping) -- then the NCCL formulation below runs."""
        from . import distributed as D
        table = self.item_embedding_table
        if self.label_smoothing or table.peer_view() is None:
            return None
        ph = getattr(self, "_peer_head_state", None)
        if ph is None or ph.cap != cap or ph.De != De or ph.mail_x.device != device:
            ph = D.PeerHead(cap, De, device, table.group)
            self._peer_head_state = ph
        return ph if ph.ok else None

    def _forward_sharded_peer(self, ph, xt_planes, xt_f32, tgt_labels, count, training, want_rank, w_planes, inv_tau):
        """SURVEY 8e over NVLink peer memory (csrc/t4r_peer.cu): no bulk collective and no host synchronisation.
        Full softmax: the ranks pull each other's label rows, score them against their V/world rows and read each
        other's (lse, label logit, rank count) statistics -- two 4-byte NCCL collectives order the steps.  Sampled
        softmax: the negatives' and the positives' rows are read from their owners' shards, after which the head is
        data parallel; only (sum of row losses, T) is all-reduced."""
        import torch.distributed as dist

        from . import distributed as D
        table = self.item_embedding_table
        Wd = table.weight.detach()
        cap = tgt_labels.numel()
        if self.sampled_softmax and training:
            if self._neg_draws is not None:
                raw = self._neg_draws
            else:
                raw = self.sampler.draw()
                dist.broadcast(raw, src=dist.get_global_rank(table.group, 0) if table.group is not None else 0,
                               group=table.group)
            neg, _, _ = self.sampler.sample(tgt_labels[:1], raw_draws=raw)
            S = neg.numel()
            _, neg_planes = table.lookup(neg)
            wy, _ = table.lookup(tgt_labels, count=count)
            col_bias = self.sampler.neg_log_q[neg].contiguous()
            pos = ops.label_logit(xt_f32, wy, torch.arange(cap, device=xt_f32.device), t_dev=count,
                                  class_bias=self.sampler.neg_log_q[tgt_labels].contiguous(), inv_temperature=inv_tau)
            res = ops.head_softmax_ce(xt_planes, xt_f32, tgt_labels, neg_planes, None, t_dev=count, inv_temperature=inv_tau,
                                      col_bias=col_bias, col_ids=neg, col_ids_sorted_unique=True, hit_value=float(torch.finfo(torch.float16).min / 100.0),
                                      pos_logit=pos, nprod=self._dense_nprod())
            n = count.to(torch.float32)
            tot = torch.cat([res["loss"].reshape(1) * n, n])
            dist.all_reduce(tot, group=table.group)
            loss = (tot[0] / tot[1].clamp(min=1.0)).reshape(())
            self._last = dict(count=count, labels=tgt_labels, sampled=True, sharded=True, S=S)
            out = LazyOutputs({"loss": loss}, {"labels": self._lazy_labels, "predictions": self._no_sharded_predictions})
            out.row_rank, out.count = None, count
            return out
        if xt_f32.data_ptr() != ph.mail_x.data_ptr():
            ph.mail_x.copy_(xt_f32)
        ph.mail_y.copy_(tgt_labels)

        def label_logits_over_all_shards(xg, yg, t_total):
            wy, _ = table.lookup(yg, count=t_total)            # W[y] from the label's owner
            return ops.label_logit(xg, wy, torch.arange(yg.numel(), device=xg.device), t_dev=t_total,
                                   inv_temperature=inv_tau)
        res = D.peer_softmax_ce(ph, count, Wd, table.num_embeddings, w_planes=w_planes, inv_tau=inv_tau,
                                want_rank=want_rank, rank_tgt_fn=label_logits_over_all_shards)
        loss = res["loss"].reshape(())
        self._last = dict(count=count, labels=tgt_labels, sampled=False, sharded=True, t_total=res["t_total"])
        out = LazyOutputs({"loss": loss}, {"labels": self._lazy_labels, "predictions": self._no_sharded_predictions})
        if want_rank:   # this rank's rows sit at [my_start, my_start + count) of the rank-major order
            idx = (torch.arange(cap, device=xt_f32.device) + res["my_start"].long()).clamp_(max=ph.cap_g - 1)
            out.row_rank = res["row_rank"][idx].contiguous()
        else:
            out.row_rank = None
        out.count = count
        return out

    def _no_sharded_predictions(self):
        raise NotImplementedError("predictions [T, V] are not materialised over a row-sharded table")

    def _lazy_labels(self):
        T = int(self._last["count"].item())
        if self._last.get("sampled"):
            # model/prediction_task.py:693-696: with sampled softmax the positive sits in column 0 of the
            # [T, 1 + S] predictions and the returned targets are all zero
            return torch.zeros_like(self._last["labels"][:T])
        return self._last["labels"][:T]

    def _lazy_predictions(self):
        st = self._last
        T = int(st["count"].item())
        w_planes = st["w_planes"] if st["w_planes"] is not None else self._planes.get("W", self.output_weight())
        logits = ops.head_logits(st["xt_planes"], w_planes, st["De"], t_dev=st["count"],
                                 inv_temperature=self._inv_tau(), nprod=self._dense_nprod())[:T]
        if not st["sampled"]:
            return logits
        # sampled softmax layout of :693: [positive | negatives], logQ-corrected, hits removed
        neg = logits * (1.0 / self._inv_tau()) + st["col_bias"].unsqueeze(0)
        hits = st["labels"][:T].unsqueeze(1) == st["neg"].unsqueeze(0)
        neg[hits] = torch.finfo(torch.float16).min / 100.0
        return torch.cat([st["pos"][:T].unsqueeze(1), neg * self._inv_tau()], dim=1)

    def remove_pad_3d(self, inp_tensor, non_pad_mask):
        inp_tensor = inp_tensor.flatten(end_dim=1)
        fl = torch.masked_select(inp_tensor, non_pad_mask.unsqueeze(1).expand_as(inp_tensor))
        return fl.view(-1, inp_tensor.size(1))

    # ------------------------------------------------------------------ metrics
    def calculate_metrics(self, predictions=None, targets=None) -> Dict[str, torch.Tensor]:
        """:481-492.  With the fused head the label ranks of the last eval forward are
        used; materialised ``predictions``/``targets`` are accepted too."""
        outputs = {}
        ranks = getattr(predictions, "row_rank", None) if predictions is not None else None
        count = getattr(predictions, "count", None)
        for metric in self.metrics:
            if ranks is not None:
                metric.update_from_ranks(ranks, count)
                outputs[self.metric_name(metric)] = metric.metric_mean[-1]
            else:
                outputs[self.metric_name(metric)] = metric(predictions, targets)
        return outputs

    def compute_metrics(self):
        """:494-512."""
        metrics = {self.metric_name(m): m.compute() for m in self.metrics if getattr(m, "top_ks", None)}
        topks = {self.metric_name(m): m.top_ks for m in self.metrics}
        results = {}
        for name, metric in metrics.items():
            if len(metric.size()) == 0:
                metric = metric.unsqueeze(0)
            for measure, k in zip(metric, topks[name]):
                results[f"{name}_{k}"] = measure
        return results
