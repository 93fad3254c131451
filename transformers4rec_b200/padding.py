"""Ragged ingest: ``__values`` / ``__offsets`` -> dense right-padded tensors.

Host-side mirror of transformers4rec/torch/utils/padding.py (``pad_batch`` :71-122,
``pad_inputs`` :125-164, called first thing by ``Model.forward``, model/base.py:551).
The reference builds a sparse COO tensor and densifies it; here one kernel
(``t4r_pad_ragged``) writes the padded / truncated rows directly.  Like the reference,
``pad_inputs`` reads the longest row length back to the host to size the output."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib
from ._lib import check, ptr


def _squeeze(t: torch.Tensor) -> torch.Tensor:
    return t.squeeze(1) if t.dim() == 2 else t


def _pad(values: Optional[torch.Tensor], offsets: Optional[torch.Tensor], rows: int, in_len: int, pad_len: int):
    src = values if values is not None else offsets
    if not src.is_cuda:
        raise _lib.T4RError("t4r_b200 padding needs CUDA tensors (no CPU fallback)")
    if values.dtype in (torch.int64,):
        eb, dtype = 8, torch.int64
    elif values.dtype in (torch.float32,):
        eb, dtype = 4, torch.float32
    elif values.dtype in (torch.int32, torch.int16, torch.uint8, torch.bool):
        values, eb, dtype = values.long(), 8, torch.int64
    else:
        values, eb, dtype = values.float(), 4, torch.float32
    values = values.contiguous()
    if offsets is not None:
        offsets = offsets.long().contiguous()
    out = torch.empty((rows, pad_len), dtype=dtype, device=values.device)
    check(_lib.load().t4r_pad_ragged(ptr(values), ptr(offsets), rows, in_len, pad_len, eb, ptr(out),
                                     torch.cuda.current_stream().cuda_stream), "t4r_pad_ragged")
    return out


def pad_ragged_tensor(values: torch.Tensor, offsets: torch.Tensor, padding_length: int) -> torch.Tensor:
    """utils/padding.py:48-68."""
    values, offsets = _squeeze(values), _squeeze(offsets)
    if values.dim() != 1:
        raise NotImplementedError("3-D ragged inputs (pretrained embeddings) are outside the t4r_b200 hot path")
    return _pad(values, offsets, offsets.numel() - 1, 0, padding_length)


def pad_dense_tensor(t: torch.Tensor, length: int) -> torch.Tensor:
    """utils/padding.py:20-30 (F.pad with a possibly negative amount = truncate)."""
    if t.dim() != 2:
        return t
    if t.shape[1] == length:
        return t
    return _pad(t, None, t.shape[0], t.shape[1], length)


def pad_batch(batch: Dict[str, torch.Tensor], padding_lengths: Dict[str, int]) -> Dict[str, torch.Tensor]:
    """utils/padding.py:71-122."""
    out = {}
    for key, value in batch.items():
        if key.endswith("__offsets"):
            col = key[: -len("__offsets")]
            length = padding_lengths.get(col)
            if length is None:
                raise ValueError(f"Found ragged column '{col}' with unspecified padding length. "
                                 "Please provide a padding length for this feature "
                                 "to be converted to a dense tensor. ")
            out[col] = pad_ragged_tensor(batch[f"{col}__values"], value, length)
        elif key.endswith("__values"):
            continue
        else:
            length = padding_lengths.get(key)
            out[key] = pad_dense_tensor(value, length) if length is not None else value
    return out


def pad_inputs(inputs: Dict[str, torch.Tensor], max_sequence_length: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """utils/padding.py:125-164."""
    batch_max = 0
    for key, val in inputs.items():
        if key.endswith("__offsets"):
            off = _squeeze(val)
            batch_max = max(int(torch.max(off[1:] - off[:-1])), batch_max)
    length = batch_max if max_sequence_length is None else min(max_sequence_length, batch_max)
    if length > 0:
        lengths = {k[: -len("__offsets")]: length for k in inputs if k.endswith("__offsets")}
        if lengths:
            inputs = pad_batch(inputs, lengths)
    return inputs
