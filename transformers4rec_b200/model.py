"""Head / Model: the call chain of model/base.py:371-407 (Head.forward) and
:544-598 (Model.forward) for the next-item path, plus the two callers SURVEY §8f N3 names on
the model itself: ``Model.fit`` (model/base.py:669-717) and ``Model.evaluate``
(:719-738), and the checkpoint surface ``Model.save`` / ``Model.load`` (:839-922: the
state dict, whose key names follow the reference's module tree).  Schemas and the HF
Trainer integration are host orchestration outside the hot path
and are not mirrored."""
from __future__ import annotations

import inspect
import logging
import os
import pathlib
from typing import Dict, List, Optional, Union

import numpy as np
import torch
from torch import nn

from .block import SequentialBlock
from .prediction_task import PredictionTask

LOG = logging.getLogger("transformers4rec_b200")


def _batches(dataloader):
    """The reference iterates ``dataloader.dataset`` when handed a torch DataLoader (model/base.py:680-683: Merlin's
    loader is an IterableDataset that yields whole batches).  A DataLoader over a map-style dataset is iterated
    itself, so that its collated batches arrive."""
    if isinstance(dataloader, torch.utils.data.DataLoader) and isinstance(dataloader.dataset,
                                                                         torch.utils.data.IterableDataset):
        return dataloader.dataset
    return dataloader


class Head(nn.Module):
    """model/base.py:227-445."""

    def __init__(self, body, prediction_tasks: Union[List[PredictionTask], PredictionTask, None] = None,
                 *more_tasks, task_blocks=None, task_weights: Optional[List[float]] = None,
                 loss_reduction: str = "mean", inputs=None, **kwargs):
        super().__init__()
        self.body = body
        self.loss_reduction = loss_reduction
        tasks = []
        if prediction_tasks is not None:
            tasks = list(prediction_tasks) if isinstance(prediction_tasks, (list, tuple)) else [prediction_tasks]
        tasks += list(more_tasks)
        self.prediction_task_dict = nn.ModuleDict()
        for t in tasks:
            self.prediction_task_dict[t.task_name] = t
        self._task_weights = {}
        if task_weights:
            for t, w in zip(tasks, task_weights):
                self._task_weights[t.task_name] = w
        self.build(inputs=inputs, task_blocks=task_blocks)

    def build(self, inputs=None, device=None, task_blocks=None):
        """model/base.py:279-304."""
        if not getattr(self.body, "output_size", lambda: None)():
            raise ValueError("Can't infer output-size of the body")
        input_size = self.body.output_size()
        for name, task in self.prediction_task_dict.items():
            task_block = task_blocks
            if task_blocks and isinstance(task_blocks, dict) and name in task_blocks:
                task_block = task_blocks[name]
            task.build(self.body, input_size, inputs=inputs, device=device, task_block=task_block)

    def forward(self, body_outputs, training=False, testing=False, targets=None, call_body=False, top_k=None, **kwargs):
        outputs = {}
        if call_body:
            body_outputs = self.body(body_outputs, training=training, testing=testing, **kwargs)
        if training or testing:
            losses = []
            for name, task in self.prediction_task_dict.items():
                task_output = task(body_outputs, targets=targets, training=training, testing=testing, **kwargs)
                if len(self.prediction_task_dict) == 1:
                    outputs = task_output
                    # model/base.py:404-407: stack().mean() of a single loss is the loss itself
                    return outputs
                losses.append(task_output["loss"] * self._task_weights.get(name, 1.0))
                outputs[name] = task_output
            loss_tensor = torch.stack(losses)
            loss = loss_tensor.mean() if self.loss_reduction == "mean" else (
                loss_tensor.sum() if self.loss_reduction == "sum" else loss_tensor)
            return {"loss": loss, **outputs}
        for name, task in self.prediction_task_dict.items():
            outputs[name] = task(body_outputs, targets=targets, training=training, testing=testing, top_k=top_k,
                                 **kwargs)
        if len(outputs) == 1:
            return next(iter(outputs.values()))
        return outputs

    def to_model(self, **kwargs):
        return Model(self, **kwargs)

    def calculate_metrics(self, predictions, targets=None):
        out = {}
        for name, task in self.prediction_task_dict.items():
            out.update(task.calculate_metrics(predictions, targets))
        return out

    def compute_metrics(self, mode=None):
        out = {}
        for task in self.prediction_task_dict.values():
            out.update(task.compute_metrics())
        return out

    def reset_metrics(self):
        for task in self.prediction_task_dict.values():
            task.reset_metrics()


class GraphedForward:
    """The forward-only pass of a model (``model(batch, training=..., testing=...)``: loss, label ranks) captured once
    into a CUDA graph and replayed: no host work per step besides copying the inputs into the captured buffers.  For the
    small configurations the eager step is launch-bound (BASELINE config 1: 27 launches, 0.44 ms eager vs 0.185 ms
    replayed); at config-2 size the two are equal.  The forward has no host synchronisation (the label count stays on
    the device), which is what makes it capturable.

    Fixed at capture time: input shapes / dtypes, the mode flags, and the WEIGHTS' split planes -- call ``recapture()``
    after the parameters change (an optimizer step, ``load_state_dict``).  The random draws of MLM masking are made
    inside the graph by torch's graph-safe generator, so every replay masks differently, as eager calls do."""

    def __init__(self, model: "Model", example_batch: Dict[str, torch.Tensor], training: bool = True, testing: bool = False,
                 warmup: int = 2):
        if not all(v.is_cuda for v in example_batch.values()):
            raise ValueError("GraphedForward needs CUDA input tensors")
        self.model, self.training, self.testing, self.warmup = model, training, testing, int(warmup)
        self.static = {k: v.clone() for k, v in example_batch.items()}
        self.graph = None
        self.recapture()

    def recapture(self):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.no_grad(), torch.cuda.stream(side):      # warm-up off the capture: workspaces, plane caches, lazy inits
            for _ in range(max(1, self.warmup)):
                self.model(self.static, training=self.training, testing=self.testing)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            out = self.model(self.static, training=self.training, testing=self.testing)
            self.loss = out["loss"]
            self.row_rank, self.count = getattr(out, "row_rank", None), getattr(out, "count", None)
        return self

    def __call__(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """-> the loss tensor (a captured buffer: read it, or copy it, before the next call)."""
        for k, buf in self.static.items():
            v = batch[k]
            if v.shape != buf.shape or v.dtype != buf.dtype:
                raise ValueError(f"GraphedForward was captured for {k!r} of shape {tuple(buf.shape)} / {buf.dtype}, "
                                 f"got {tuple(v.shape)} / {v.dtype}")
            buf.copy_(v, non_blocking=True)
        self.graph.replay()
        return self.loss


class Model(nn.Module):
    """model/base.py:448-598."""

    def __init__(self, *head: Head, head_weights=None, head_reduction: str = "mean", **kwargs):
        super().__init__()
        self.heads = nn.ModuleList(head)
        self.head_weights = head_weights or [1.0] * len(head)
        self.head_reduction = head_reduction
        self.top_k = kwargs.get("top_k", None)
        self.max_sequence_length = kwargs.get("max_sequence_length", None)
        self.name = kwargs.get("name", None)

    def graphed(self, example_batch: Dict[str, torch.Tensor], training: bool = True, testing: bool = False) -> GraphedForward:
        """CUDA-graph replay of the forward-only pass for batches shaped like ``example_batch`` (see GraphedForward)."""
        return GraphedForward(self, example_batch, training=training, testing=testing)

    def enable_fused_training(self, on: bool = True, head_chunk: int = 32768):
        """N3: route ``model(batch, training=True)`` through the fused training step whenever autograd is recording, so
        that the usual ``loss = model(batch, training=True)["loss"]; loss.backward(); optimizer.step()`` (HF Trainer's
        compute_loss / training_step, trainer.py:315-338 in the reference) trains on the t4r kernels.  Opt-in; with it
        off (the default) or under ``torch.no_grad()`` the forward-only path runs as before.  The step reads the model's
        structure once, here: call it after the model is final (device placement, ``shard_item_table``, ``pre`` ...)."""
        if on:
            from .training import FusedTrainingStep
            self._fused_step = FusedTrainingStep(self, head_chunk=head_chunk)
        else:
            self._fused_step = None
        return self

    def forward(self, inputs: Dict[str, torch.Tensor], targets=None, training=False, testing=False, **kwargs):
        # model/base.py:546-548: floating inputs are cast to fp32
        for name, val in inputs.items():
            if torch.is_floating_point(val) and val.dtype != torch.float32:
                inputs[name] = val.to(torch.float32)
        # model/base.py:551: ragged (__values/__offsets) inputs become dense padded tensors
        if any(k.endswith("__offsets") for k in inputs):
            from .padding import pad_inputs
            inputs = pad_inputs(inputs, max_sequence_length=self.max_sequence_length)
        if training and not testing and getattr(self, "_fused_step", None) is not None and torch.is_grad_enabled():
            from .training import training_loss
            step = self._fused_step
            loss = training_loss(self, inputs, step)
            from .prediction_task import LazyOutputs
            out = LazyOutputs({"loss": loss, "labels": step.labels[:step.T]}, {})
            out.row_rank, out.count = step.row_rank, None    # label ranks among the training logits, when asked for
            return out
        if len(self.heads) == 1:
            # :574-576 stack().mean() over one head is the identity
            return self.heads[0](inputs, call_body=True, targets=targets, training=training, testing=testing,
                                 top_k=self.top_k, **kwargs)
        outs = [h(inputs, call_body=True, targets=targets, training=training, testing=testing, top_k=self.top_k,
                  **kwargs) for h in self.heads]
        if training or testing:
            losses = torch.stack([o["loss"] * w for o, w in zip(outs, self.head_weights)])
            return {"loss": losses.mean() if self.head_reduction == "mean" else losses.sum(), "heads": outs}
        return outs

    # ------------------------------------------------------------------ fit / evaluate (model/base.py:669-738)
    def fit(self, dataloader, optimizer=None, eval_dataloader=None, num_epochs=1, amp=False, train=True, verbose=True,
            compute_metric=True):
        """model/base.py:669-717: ``num_epochs`` passes over ``dataloader`` (batches ``(x, y)``), one optimizer step
        per batch; returns the mean loss of every epoch as a numpy array.

        Training runs on the fused step (``enable_fused_training``: forward + backward on the t4r kernels); the
        default optimizer is ``FusedAdamW`` with lr 1e-3 and no weight decay, i.e. the reference's ``torch.optim.Adam``
        default computed by the t4r update kernel.  An optimizer class is instantiated on ``self.parameters()``
        (as in the reference), an instance is used as is.  The fused head never materialises the [T, V] training
        logits: with ``compute_metric`` the training head also emits every label's rank among them (replicated full
        softmax) and the streaming metrics are updated from those ranks; with sampled softmax or a row-sharded table
        train-mode metrics are not accumulated -- use ``eval_dataloader`` / ``evaluate``.  ``amp`` is accepted and ignored: the
        path's GEMMs are fixed split-bf16 tensor-core products at fp32 accuracy, there is no autocast switch."""
        if optimizer is None:
            from .training import FusedAdamW
            optimizer = FusedAdamW(self.parameters(), lr=1e-3, weight_decay=0.0)
        elif inspect.isclass(optimizer):
            optimizer = optimizer(self.parameters())
        if amp:
            LOG.warning("Model.fit: amp=True has no effect on the t4r path (fp32-grade tensor-core products)")
        if train and getattr(self, "_fused_step", None) is None:
            self.enable_fused_training()
        if train:
            self._fused_step.want_rank = bool(compute_metric)
        self.train(mode=train)
        epoch_losses = []
        with torch.set_grad_enabled(mode=train):
            for _ in range(num_epochs):
                losses = []
                it = iter(_batches(dataloader))
                if verbose:
                    from tqdm import tqdm
                    it = tqdm(it)
                for x, y in it:
                    output = self(x, targets=y, training=True)
                    losses.append(float(output["loss"].detach()))
                    if compute_metric and (not train or getattr(output, "row_rank", None) is not None):
                        self._update_metrics(output)
                    if train:
                        optimizer.zero_grad()
                        output["loss"].backward()
                        optimizer.step()
                if verbose:
                    if compute_metric:
                        print(self.compute_metrics(mode="train"))
                    if eval_dataloader:
                        print(self.evaluate(eval_dataloader, verbose=False))
                        self.train(mode=train)
                epoch_losses.append(np.mean(losses))
        return np.array(epoch_losses)

    def evaluate(self, dataloader, targets=None, training=False, testing=True, verbose=True, mode="eval"):
        """model/base.py:719-738: reset the metrics, run every batch ``(x, y)`` through the evaluation forward, update
        the streaming metrics (from the label ranks the fused head produced -- no [T, V] logits are materialised) and
        return ``compute_metrics``."""
        it = iter(_batches(dataloader))
        if verbose:
            from tqdm import tqdm
            it = tqdm(it)
        self.reset_metrics()
        with torch.no_grad():
            for x, y in it:
                output = self(x, targets=y, training=training, testing=testing)
                self._update_metrics(output)
        return self.compute_metrics(mode=mode)

    def _update_metrics(self, output):
        # the fused head hands back the label ranks with its outputs; without them (e.g. the reference's plain
        # ``predictions`` / ``labels`` pair) the metrics take the materialised tensors, model/base.py:704-707
        if getattr(output, "row_rank", None) is not None:
            return self.calculate_metrics(output)
        return self.calculate_metrics(output["predictions"], targets=output["labels"])

    # ------------------------------------------------------------------ checkpoint surface (model/base.py:839-922)
    def save(self, path: Union[str, os.PathLike], model_name: str = "t4rec_model_class"):
        """model/base.py:839-856: the state dict (weights only) goes to ``{path}/{model_name}.pt``; the architecture is
        rebuilt by the caller before ``Model.load``.  A row-sharded item table is saved as THIS rank's rows."""
        export_path = pathlib.Path(path)
        export_path.mkdir(exist_ok=True)
        torch.save(self.state_dict(), export_path / (model_name + ".pt"))

    @classmethod
    def load(cls, state_dict: Dict[str, torch.Tensor], heads, head_weights=None, head_reduction: str = "mean",
             optimizer=None, name: Optional[str] = None, max_sequence_length: Optional[int] = None,
             top_k: Optional[int] = None, strict: bool = True) -> "Model":
        """model/base.py:858-922: a Model over already-built ``heads`` carrying ``state_dict`` (no file I/O here)."""
        if not isinstance(heads, (list, nn.ModuleList)):
            heads = [heads]
        model = cls(*heads, head_weights=head_weights, head_reduction=head_reduction, name=name,
                    max_sequence_length=max_sequence_length, top_k=top_k)
        if not isinstance(state_dict, dict) or not all(isinstance(v, torch.Tensor) for v in state_dict.values()):
            raise ValueError("`state_dict` must be a dictionary of parameter (torch) tensors.")
        model.load_state_dict(state_dict, strict=strict)
        return model

    def calculate_metrics(self, predictions, targets=None):
        out = {}
        for h in self.heads:
            out.update(h.calculate_metrics(predictions, targets))
        return out

    def compute_metrics(self, mode=None):
        out = {}
        for h in self.heads:
            out.update(h.compute_metrics(mode))
        return out

    def reset_metrics(self):
        for h in self.heads:
            h.reset_metrics()
