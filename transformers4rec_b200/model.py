"""Head / Model: the call chain of model/base.py:371-407 (Head.forward) and
:544-598 (Model.forward) for the next-item path.  fit/save/load, schemas and the
HF Trainer integration are host orchestration outside the hot path (SURVEY §2
row 12) and are not mirrored."""
from __future__ import annotations

from typing import Dict, List, Optional, Union

import torch
from torch import nn

from .block import SequentialBlock
from .prediction_task import PredictionTask


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


class Model(nn.Module):
    """model/base.py:448-598."""

    def __init__(self, *head: Head, head_weights=None, head_reduction: str = "mean", **kwargs):
        super().__init__()
        self.heads = nn.ModuleList(head)
        self.head_weights = head_weights or [1.0] * len(head)
        self.head_reduction = head_reduction
        self.top_k = kwargs.get("top_k", None)
        self.max_sequence_length = kwargs.get("max_sequence_length", None)

    def enable_fused_training(self, on: bool = True, head_chunk: int = 32768):
        """N3: route ``model(batch, training=True)`` through the fused training step whenever autograd is recording, so
        that the usual ``loss = model(batch, training=True)["loss"]; loss.backward(); optimizer.step()`` (HF Trainer's
        compute_loss / training_step, trainer.py:315-338 in the reference) trains on the t4r kernels.  Opt-in; with it
        off (the default) or under ``torch.no_grad()`` the forward-only path runs as before."""
        if on:
            from .training import FusedTrainingStep
            self._fused_step = FusedTrainingStep(self, head_chunk=head_chunk)
        else:
            self._fused_step = None
        return self

    def forward(self, inputs: Dict[str, torch.Tensor], targets=None, training=False, testing=False, **kwargs):
        if training and not testing and getattr(self, "_fused_step", None) is not None and torch.is_grad_enabled():
            from .training import training_loss
            step = self._fused_step
            loss = training_loss(self, inputs, step)
            return {"loss": loss, "labels": step.labels[:step.T]}
        # model/base.py:546-548: floating inputs are cast to fp32
        for name, val in inputs.items():
            if torch.is_floating_point(val) and val.dtype != torch.float32:
                inputs[name] = val.to(torch.float32)
        # model/base.py:551: ragged (__values/__offsets) inputs become dense padded tensors
        if any(k.endswith("__offsets") for k in inputs):
            from .padding import pad_inputs
            inputs = pad_inputs(inputs, max_sequence_length=self.max_sequence_length)
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
