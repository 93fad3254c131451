"""Functional wrappers: torch CUDA tensors in, C-ABI calls out.

PyTorch is plumbing here (device memory, streams); every computation below runs
in ``libt4r_b200.so``.  All functions require CUDA tensors and raise otherwise --
there is no eager/CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import check, ptr


def round_up64(k: int) -> int:
    return (k + 63) // 64 * 64


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.T4RError("t4r_b200 ops need CUDA tensors (no CPU fallback); got a tensor on " + str(t.device))


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


class Workspace:
    """Grow-only byte buffer per (device, tag) so steady-state steps never allocate."""

    def __init__(self):
        self._bufs: Dict[Tuple[str, str], torch.Tensor] = {}

    def get(self, tag: str, nbytes: int, device) -> torch.Tensor:
        key = (str(device), tag)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            self._bufs[key] = buf
        return buf

    def tensor(self, tag: str, shape: Sequence[int], dtype, device) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        buf = self.get(tag, max(nbytes, 256), device)
        return buf[:nbytes].view(dtype).view(*shape)


WS = Workspace()
# bench.py sets this to a (start, stop) pair of torch.cuda.Event(enable_timing=True) to time
# the logits/LSE GEMM of every head call on its own stream
HEAD_EVENTS = None


# --------------------------------------------------------------------------- #
# operand packing
# --------------------------------------------------------------------------- #
def split_planes(x: torch.Tensor, row_code: Optional[torch.Tensor] = None, mask_vec: Optional[torch.Tensor] = None,
                 want_f32: bool = False, out: Optional[torch.Tensor] = None):
    """fp32 [rows, K] -> bf16 planes [2, rows, Kp] (optionally applying row codes)."""
    _need_cuda(x)
    x = _f32c(x)
    rows, K = x.shape
    Kp = round_up64(K)
    planes = out if out is not None else torch.empty((2, rows, Kp), dtype=torch.bfloat16, device=x.device)
    of = torch.empty((rows, K), dtype=torch.float32, device=x.device) if want_f32 else None
    check(_lib.load().t4r_split_planes(ptr(x), rows, K, K, ptr(row_code), ptr(mask_vec), ptr(of), ptr(planes),
                                       _stream()), "t4r_split_planes")
    return (planes, of) if want_f32 else planes


def split_planes_mixed(x: torch.Tensor, count: Optional[torch.Tensor] = None):
    """fp32 [rows, K] -> (int16 words [2, rows, Kp], fp32 [rows] inverse row scales): the operands of the 2-unit
    product (``nprod=2``: fp16 x fp16 + two e4m3 cross terms, csrc/t4r_mixed_pack.cuh).  ``count`` (device int32):
    only the first round_up(count, 256) rows are packed (the rest of the outputs is left untouched)."""
    _need_cuda(x, count)
    x = _f32c(x)
    rows, K = x.shape
    Kp = round_up64(K)
    planes = torch.empty((2, rows, Kp), dtype=torch.int16, device=x.device)
    inv = torch.empty((rows,), dtype=torch.float32, device=x.device)
    check(_lib.load().t4r_split_planes_mixed_n(ptr(x), rows, K, K, ptr(count), ptr(planes), ptr(inv), _stream()),
          "t4r_split_planes_mixed_n")
    return planes, inv


def split_planes_mixed_host(x: torch.Tensor):
    """The same packing code compiled for the host (CPU tensors; test infrastructure, no CUDA call)."""
    assert not x.is_cuda
    x = _f32c(x)
    rows, K = x.shape
    Kp = round_up64(K)
    planes = torch.empty((2, rows, Kp), dtype=torch.int16)
    inv = torch.empty((rows,), dtype=torch.float32)
    check(_lib.load().t4r_debug_split_planes_mixed_host(ptr(x), rows, K, K, ptr(planes), ptr(inv)),
          "t4r_debug_split_planes_mixed_host")
    return planes, inv


class PlaneCache:
    """Split-bf16 copies of weights, refreshed when the parameter changes
    (``data_ptr`` / ``_version``).  ``transform`` maps the parameter to its [N, K]
    (nn.Linear-style) layout before packing."""

    def __init__(self):
        self._cache = {}

    def get(self, key: str, param: torch.Tensor, transform=None) -> torch.Tensor:
        sig = (param.data_ptr(), param._version, tuple(param.shape), str(param.device))
        ent = self._cache.get(key)
        if ent is not None and ent[0] == sig:
            return ent[1]
        with torch.no_grad():
            w = param.detach()
            if transform is not None:
                w = transform(w)
            planes = split_planes(w)
        self._cache[key] = (sig, planes)
        return planes

    def get_mixed(self, key: str, param: torch.Tensor):
        """(mixed planes, inverse row scales) of a parameter for the 2-unit product, cached like ``get``."""
        sig = (param.data_ptr(), param._version, tuple(param.shape), str(param.device), "mixed")
        ent = self._cache.get(key + "#mixed")
        if ent is not None and ent[0] == sig:
            return ent[1]
        with torch.no_grad():
            val = split_planes_mixed(param.detach())
        self._cache[key + "#mixed"] = (sig, val)
        return val

    def drop(self, key: str):
        self._cache.pop(key, None)
        self._cache.pop(key + "#mixed", None)

    def clear(self):
        self._cache.clear()


# --------------------------------------------------------------------------- #
# K1
# --------------------------------------------------------------------------- #
def embed_concat(cats: List[Tuple[torch.Tensor, torch.Tensor, int]], conts: List[Tuple[torch.Tensor, int]], M: int,
                 C_width: int, want_f32: bool, want_planes: bool):
    """cats: (table [rows, dim], ids [M] int64, first column); conts: (values [M] f32, column)."""
    lib = _lib.load()
    fl = _lib.FeatureList()
    fl.n_cat, fl.n_cont = len(cats), len(conts)
    if len(cats) > _lib.T4R_MAX_FEATURES or len(conts) > _lib.T4R_MAX_FEATURES:
        raise _lib.T4RError(f"at most {_lib.T4R_MAX_FEATURES} categorical / continuous features per call")
    keep = []
    dev = None
    for i, (table, ids, col) in enumerate(cats):
        _need_cuda(table, ids)
        table = _f32c(table)
        ids = ids.reshape(-1)
        if ids.dtype != torch.int64:
            ids = ids.long()
        ids = ids.contiguous()
        if ids.numel() != M:
            raise _lib.T4RError(f"embed_concat: feature {i} carries {ids.numel()} ids, expected M = {M}")
        keep += [table, ids]
        fl.table[i], fl.ids[i] = table.data_ptr(), ids.data_ptr()
        fl.table_rows[i], fl.dim[i], fl.cat_col[i] = table.shape[0], table.shape[1], col
        dev = table.device
    for i, (vals, col) in enumerate(conts):
        _need_cuda(vals)
        vals = _f32c(vals.reshape(-1))
        if vals.numel() != M:
            raise _lib.T4RError(f"embed_concat: continuous feature {i} carries {vals.numel()} values, expected M = {M}")
        keep.append(vals)
        fl.cont[i], fl.cont_col[i] = vals.data_ptr(), col
        dev = dev or vals.device
    out_f32 = torch.empty((M, C_width), dtype=torch.float32, device=dev) if want_f32 else None
    planes = torch.empty((2, M, round_up64(C_width)), dtype=torch.bfloat16, device=dev) if want_planes else None
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib.t4r_embed_concat_fwd(C.byref(fl), M, C_width, ptr(out_f32), ptr(planes), ptr(err), _stream()),
          "t4r_embed_concat_fwd")
    return out_f32, planes, err


def input_block(feats: List[dict], M: int, L: int, C_width: int, agg: int = _lib.AGG_CONCAT, item_feature: int = -1,
                ln_eps: float = 1e-5, want_f32: bool = True, want_planes: bool = False):
    """General input block (t4r_input_block_fwd).  ``feats``: dicts in sorted-name order with keys
    kind, dim, col, input and, by kind, table / soft_w / soft_b / card, optional ln=(gamma, beta),
    per_session."""
    if not feats or len(feats) > _lib.T4R_MAX_FEATURES:
        raise _lib.T4RError(f"1..{_lib.T4R_MAX_FEATURES} features per call")
    arr = (_lib.Feature * len(feats))()
    keep = []
    dev = None

    def f32(t):
        t = _f32c(t.detach())
        keep.append(t)
        return t.data_ptr()

    for i, f in enumerate(feats):
        a = arr[i]
        a.kind, a.dim, a.col = int(f["kind"]), int(f["dim"]), int(f.get("col", 0))
        a.per_session = 1 if f.get("per_session") else 0
        x = f["input"]
        _need_cuda(x)
        dev = x.device
        if a.kind == _lib.FEAT_CAT:
            x = x.reshape(-1)
            x = (x if x.dtype == torch.int64 else x.long()).contiguous()
            keep.append(x)
            a.input = x.data_ptr()
        else:
            a.input = f32(x.reshape(-1) if a.kind != _lib.FEAT_DENSE else x.reshape(-1, a.dim))
        if a.kind in (_lib.FEAT_CAT, _lib.FEAT_SOFT):
            _need_cuda(f["table"])
            a.table = f32(f["table"])
            a.card = int(f["table"].shape[0])
        if a.kind == _lib.FEAT_SOFT:
            a.soft_w, a.soft_b = f32(f["soft_w"].reshape(-1)), f32(f["soft_b"].reshape(-1))
        if f.get("ln") is not None:
            a.ln_gamma, a.ln_beta = f32(f["ln"][0]), f32(f["ln"][1])
    out_f32 = torch.empty((M, C_width), dtype=torch.float32, device=dev) if want_f32 else None
    planes = torch.empty((2, M, round_up64(C_width)), dtype=torch.bfloat16, device=dev) if want_planes else None
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    check(_lib.load().t4r_input_block_fwd(arr, len(feats), M, L, agg, item_feature, ln_eps, C_width, ptr(out_f32),
                                          ptr(planes), ptr(err), _stream()), "t4r_input_block_fwd")
    return out_f32, planes, err


def swap_noise(values: torch.Tensor, keep_mask: Optional[torch.Tensor], u: torch.Tensor, perm: torch.Tensor,
               replacement_prob: float) -> torch.Tensor:
    """StochasticSwapNoise.augment with explicit draws: ``u`` uniforms (shape of ``values``), ``perm`` a
    permutation of range(number of kept positions)."""
    _need_cuda(values, keep_mask, u, perm)
    if values.dtype not in (torch.int64, torch.float32):
        raise _lib.T4RError(f"swap_noise: int64 or float32 tensors (got {values.dtype})")
    v = values.contiguous()
    n = v.numel()
    inner, stride, km = 1, 1, None
    if keep_mask is not None:
        km = keep_mask
        if v.dim() == km.dim() - 1:  # transformations.py:63-65: context feature, mask[:, 0]
            stride = km.shape[1] if km.dim() > 1 else 1
            km = km.contiguous()
        else:
            km = km.contiguous()
            inner = max(1, n // km.numel())
        km = km.view(torch.uint8) if km.dtype == torch.bool else km.to(torch.uint8)
    u = _f32c(u).reshape(-1)
    perm = perm.long().contiguous()
    out = torch.empty_like(v)
    scratch = torch.empty(n, dtype=torch.int32, device=v.device)
    check(_lib.load().t4r_swap_noise(ptr(v), v.element_size(), n, ptr(km), stride, inner, ptr(u), float(replacement_prob),
                                     ptr(perm), ptr(scratch), ptr(out), _stream()), "t4r_swap_noise")
    return out


def metrics_from_ranks(row_rank: torch.Tensor, ks: Sequence[int], kind: int, t_dev=None) -> torch.Tensor:
    _need_cuda(row_rank)
    outs = []
    ks = [int(k) for k in ks]
    for i in range(0, len(ks), 4):
        part = ks[i:i + 4]
        out = torch.empty(len(part), dtype=torch.float32, device=row_rank.device)
        arr = (C.c_int32 * len(part))(*part)
        check(_lib.load().t4r_metrics_from_ranks(ptr(row_rank), ptr(t_dev), row_rank.numel(), kind, arr, len(part),
                                                 ptr(out), _stream()), "t4r_metrics_from_ranks")
        outs.append(out)
    return torch.cat(outs) if len(outs) > 1 else outs[0]


# --------------------------------------------------------------------------- #
# K3
# --------------------------------------------------------------------------- #
def mask_mlm(item_ids: torch.Tensor, mode: int, padding_idx: int = 0, mlm_probability: float = 0.15,
             u: Optional[torch.Tensor] = None):
    _need_cuda(item_ids, u)
    ids = item_ids.long().contiguous()
    B, L = ids.shape
    Lo = L + 1 if mode == _lib.MLM_INFERENCE else L
    if mode == _lib.MLM_TRAIN:
        if u is None:
            u = torch.rand((B, L + 2), dtype=torch.float32, device=ids.device)
        u = _f32c(u)
        assert tuple(u.shape) == (B, L + 2), "u must be [B, L+2]"
    mask = torch.empty((B, Lo), dtype=torch.bool, device=ids.device)
    labels = torch.empty((B, Lo), dtype=torch.int64, device=ids.device)
    code = torch.empty((B, Lo), dtype=torch.uint8, device=ids.device)
    check(_lib.load().t4r_mask_mlm(ptr(ids), B, L, padding_idx, mode, mlm_probability, ptr(u), ptr(mask), ptr(labels),
                                   ptr(code), _stream()), "t4r_mask_mlm")
    return mask, labels, code


def mask_clm(item_ids: torch.Tensor, mode: int, padding_idx: int = 0):
    _need_cuda(item_ids)
    ids = item_ids.long().contiguous()
    B, L = ids.shape
    mask = torch.empty((B, L), dtype=torch.bool, device=ids.device)
    labels = torch.empty((B, L), dtype=torch.int64, device=ids.device)
    code = torch.empty((B, L), dtype=torch.uint8, device=ids.device)
    check(_lib.load().t4r_mask_clm(ptr(ids), B, L, padding_idx, mode, ptr(mask), ptr(labels), ptr(code), _stream()),
          "t4r_mask_clm")
    return mask, labels, code


def plm_context_lengths(max_span_length: int, plm_probability: float):
    """masking.py:608 ``int(span_length / plm_probability)`` for span 0..max (index 0 unused), python float math."""
    return [0] + [int(sp / plm_probability) for sp in range(1, max_span_length + 1)]


def _plm_call(fn, ids, mode, padding_idx, max_span, ctx_len, draws, stream_arg):
    B, L = ids.shape
    dev = ids.device
    mask = torch.empty((B, L), dtype=torch.bool, device=dev)
    labels = torch.empty((B, L), dtype=torch.int64, device=dev)
    perm_mask = torch.empty((B, L, L), dtype=torch.uint8, device=dev)
    keep = []
    ptrs = [None] * 5
    if mode == _lib.PLM_TRAIN:
        for i, (key, dt) in enumerate((("u_span", torch.float32), ("u_start", torch.float32), ("u_force", torch.float32),
                                       ("u_unmask", torch.float32), ("perm", torch.int32))):
            t = draws[key].to(dt).contiguous()
            assert tuple(t.shape) == ((B, L) if key in ("u_span", "u_start", "perm") else (B,)), key
            keep.append(t)
            ptrs[i] = ptr(t)
    arr = (C.c_int32 * (max_span + 1))(*[int(v) for v in ctx_len]) if mode == _lib.PLM_TRAIN else None
    check(fn(ptr(ids), B, L, padding_idx, mode, max_span, arr, *ptrs, ptr(mask), ptr(labels), ptr(perm_mask), *stream_arg),
          "t4r_mask_plm")
    return mask, labels, perm_mask


def mask_plm(item_ids: torch.Tensor, mode: int, padding_idx: int = 0, max_span_length: int = 5,
             plm_probability: float = 1 / 6, draws: Optional[dict] = None):
    """PLM labels + permutation mask (t4r_mask_plm).  ``draws`` (training): dict(u_span, u_start [B, L], u_force,
    u_unmask [B], perm [B, L]); generated on the device when omitted.  Returns (mask_schema bool, masked_targets i64,
    perm_mask uint8 [B, L, L])."""
    _need_cuda(item_ids)
    ids = item_ids.long().contiguous()
    B, L = ids.shape
    if mode == _lib.PLM_TRAIN and draws is None:
        dev = ids.device
        draws = {"u_span": torch.rand((B, L), device=dev), "u_start": torch.rand((B, L), device=dev),
                 "u_force": torch.rand((B,), device=dev), "u_unmask": torch.rand((B,), device=dev),
                 "perm": torch.argsort(torch.rand((B, L), device=dev), dim=1)}   # a uniform random permutation per row
    return _plm_call(_lib.load().t4r_mask_plm, ids, mode, padding_idx, max_span_length,
                     plm_context_lengths(max_span_length, plm_probability), draws, (_stream(),))


def mask_plm_host(item_ids: torch.Tensor, mode: int, padding_idx: int = 0, max_span_length: int = 5,
                  plm_probability: float = 1 / 6, draws: Optional[dict] = None):
    """The same code compiled for the host (CPU tensors; test infrastructure)."""
    assert not item_ids.is_cuda
    ids = item_ids.long().contiguous()
    return _plm_call(_lib.load().t4r_debug_mask_plm_host, ids, mode, padding_idx, max_span_length,
                     plm_context_lengths(max_span_length, plm_probability), draws, ())


def compact_targets(masked_targets: torch.Tensor, padding_idx: int = 0):
    _need_cuda(masked_targets)
    mt = masked_targets.long().contiguous().reshape(-1)
    n = mt.numel()
    rows = torch.empty(n, dtype=torch.int32, device=mt.device)
    labels = torch.empty(n, dtype=torch.int64, device=mt.device)
    count = torch.empty(1, dtype=torch.int32, device=mt.device)
    check(_lib.load().t4r_compact_targets(ptr(mt), n, padding_idx, ptr(rows), ptr(labels), ptr(count), _stream()),
          "t4r_compact_targets")
    return rows, labels, count


def gather_rows_split(x2d: torch.Tensor, idx: torch.Tensor, count: Optional[torch.Tensor], cap: int,
                      want_f32: bool = True, out_f32: Optional[torch.Tensor] = None):
    """``out_f32``: caller-owned fp32 [cap, K] destination (e.g. a peer window, distributed.PeerHead)."""
    _need_cuda(x2d, idx, count)
    x2d = _f32c(x2d)
    K = x2d.shape[1]
    planes = torch.empty((2, cap, round_up64(K)), dtype=torch.bfloat16, device=x2d.device)
    if out_f32 is not None:
        assert out_f32.shape == (cap, K) and out_f32.dtype == torch.float32 and out_f32.is_contiguous()
        of = out_f32
    else:
        of = torch.empty((cap, K), dtype=torch.float32, device=x2d.device) if want_f32 else None
    lib = _lib.load()
    if idx.dtype == torch.int32:
        check(lib.t4r_gather_rows_split(ptr(x2d), K, K, ptr(idx), ptr(count), cap, ptr(of), ptr(planes), _stream()),
              "t4r_gather_rows_split")
    else:
        assert count is None
        idx = idx.long().contiguous()
        check(lib.t4r_gather_rows_split_i64(ptr(x2d), K, K, ptr(idx), cap, ptr(of), ptr(planes), _stream()),
              "t4r_gather_rows_split_i64")
    return planes, of


# --------------------------------------------------------------------------- #
# K2
# --------------------------------------------------------------------------- #
def linear(x_planes: torch.Tensor, w_planes: torch.Tensor, K: int, *, bias=None, act: int = _lib.ACT_NONE,
           row_code=None, mask_vec=None, residual=None, ln=None, ln_eps: float = 0.0, want_f32=True,
           want_planes=True, want_pre_ln=False, m_dev=None, nprod: int = 3):
    """Y = epilogue(X W^T): x_planes [2, M, Kp], w_planes [2, N, Kp]."""
    _need_cuda(x_planes, w_planes)
    M, N = x_planes.shape[1], w_planes.shape[1]
    dev = x_planes.device
    a = _lib.LinearArgs()
    a.M, a.N, a.K = M, N, K
    a.x_planes, a.w_planes = ptr(x_planes), ptr(w_planes)
    a.m_dev = ptr(m_dev)
    keep = []
    if bias is not None:
        bias = _f32c(bias.detach()); keep.append(bias); a.bias = ptr(bias)
    a.act = act
    if row_code is not None:
        rc = row_code.reshape(-1).contiguous(); mv = _f32c(mask_vec.detach()); keep += [rc, mv]
        a.row_code, a.mask_vec = ptr(rc), ptr(mv)
    if residual is not None:
        residual = _f32c(residual); keep.append(residual); a.residual = ptr(residual)
    if ln is not None:
        g, b = _f32c(ln[0].detach()), _f32c(ln[1].detach()); keep += [g, b]
        a.ln_gamma, a.ln_beta, a.ln_eps = ptr(g), ptr(b), ln_eps
    out_f32 = torch.empty((M, N), dtype=torch.float32, device=dev) if want_f32 else None
    out_pre = torch.empty((M, N), dtype=torch.float32, device=dev) if want_pre_ln else None
    out_planes = torch.empty((2, M, round_up64(N)), dtype=torch.bfloat16, device=dev) if want_planes else None
    a.out_f32, a.out_pre_ln, a.out_planes = ptr(out_f32), ptr(out_pre), ptr(out_planes)
    a.nprod = nprod
    check(_lib.load().t4r_linear_fwd(C.byref(a), _stream()), "t4r_linear_fwd")
    return out_f32, out_planes, out_pre


def ffn(x_planes: torch.Tensor, w1_planes: torch.Tensor, b1: torch.Tensor, w2_planes: torch.Tensor, b2, ln, ln_eps: float,
        *, residual=None, want_f32=True, want_planes=False, want_pre_ln=False):
    """LayerNorm(residual + gelu(X W1^T + b1) W2^T + b2) in one kernel; residual None -> X itself."""
    _need_cuda(x_planes, w1_planes, w2_planes)
    M, d = x_planes.shape[1], x_planes.shape[2]
    hidden = w1_planes.shape[1]
    dev = x_planes.device
    b1 = _f32c(b1.detach())
    b2 = _f32c(b2.detach()) if b2 is not None else None
    g, b = _f32c(ln[0].detach()), _f32c(ln[1].detach())
    residual = _f32c(residual) if residual is not None else None
    out_f32 = torch.empty((M, d), dtype=torch.float32, device=dev) if want_f32 else None
    out_pre = torch.empty((M, d), dtype=torch.float32, device=dev) if want_pre_ln else None
    out_planes = torch.empty((2, M, d), dtype=torch.bfloat16, device=dev) if want_planes else None
    check(_lib.load().t4r_ffn_fwd(ptr(x_planes), M, d, hidden, ptr(w1_planes), ptr(b1), ptr(w2_planes), ptr(b2),
                                  ptr(residual), ptr(g), ptr(b), ln_eps, ptr(out_pre), ptr(out_f32), ptr(out_planes),
                                  _stream()), "t4r_ffn_fwd")
    return out_f32, out_planes, out_pre


def debug_sgemm_nt(A: torch.Tensor, B: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(A, B, bias)
    A, B = _f32c(A), _f32c(B)
    M, K = A.shape
    N = B.shape[0]
    Cm = torch.empty((M, N), dtype=torch.float32, device=A.device)
    check(_lib.load().t4r_debug_sgemm_nt(ptr(A), ptr(B), ptr(bias), ptr(Cm), M, N, K, _stream()), "t4r_debug_sgemm_nt")
    return Cm


# --------------------------------------------------------------------------- #
# encoders
# --------------------------------------------------------------------------- #
def xlnet_encoder(layers_struct, n_layer: int, B: int, L: int, d: int, n_head: int, eps: float, x_f32: torch.Tensor,
                  x_planes: Optional[torch.Tensor], want_planes: bool = False):
    _need_cuda(x_f32, x_planes)
    lib = _lib.load()
    x_f32 = _f32c(x_f32)
    dev = x_f32.device
    nbytes = lib.t4r_xlnet_encoder_workspace_bytes(B, L, d, n_head)
    ws = WS.get("xlnet", nbytes, dev)
    out = torch.empty((B * L, d), dtype=torch.float32, device=dev)
    out_planes = torch.empty((2, B * L, d), dtype=torch.bfloat16, device=dev) if want_planes else None
    check(lib.t4r_xlnet_encoder_fwd(layers_struct, n_layer, B, L, d, n_head, eps, ptr(x_f32), ptr(x_planes), ptr(out),
                                    ptr(out_planes), ptr(ws), ws.numel(), _stream()), "t4r_xlnet_encoder_fwd")
    return out, out_planes


def xlnet_encoder_plm(layers_struct, n_layer: int, B: int, L: int, d: int, n_head: int, eps: float, x2_f32: torch.Tensor,
                      perm_mask: torch.Tensor):
    """Two-stream XLNet forward (PLM): ``x2_f32`` [2 B L, d] = content-stream rows then query-stream rows; returns the
    same layout (the second half is HF's output[0])."""
    _need_cuda(x2_f32, perm_mask)
    lib = _lib.load()
    x2_f32 = _f32c(x2_f32)
    pm = perm_mask.to(torch.uint8).contiguous()
    dev = x2_f32.device
    nbytes = lib.t4r_xlnet_encoder_workspace_bytes(2 * B, L, d, n_head)
    ws = WS.get("xlnet_plm", nbytes, dev)
    out = torch.empty((2 * B * L, d), dtype=torch.float32, device=dev)
    check(lib.t4r_xlnet_encoder_plm_fwd(layers_struct, n_layer, B, L, d, n_head, eps, ptr(x2_f32), ptr(pm), ptr(out),
                                        ptr(ws), ws.numel(), _stream()), "t4r_xlnet_encoder_plm_fwd")
    return out


def gpt2_encoder(layers_struct, n_layer: int, B: int, L: int, d: int, n_head: int, eps: float, wpe, lnf_g, lnf_b,
                 x_f32: torch.Tensor, want_planes: bool = False):
    _need_cuda(x_f32)
    lib = _lib.load()
    x_f32 = _f32c(x_f32)
    dev = x_f32.device
    nbytes = lib.t4r_gpt2_encoder_workspace_bytes(B, L, d, n_head)
    ws = WS.get("gpt2", nbytes, dev)
    out = torch.empty((B * L, d), dtype=torch.float32, device=dev)
    out_planes = torch.empty((2, B * L, d), dtype=torch.bfloat16, device=dev) if want_planes else None
    check(lib.t4r_gpt2_encoder_fwd(layers_struct, n_layer, B, L, d, n_head, eps, ptr(wpe), ptr(lnf_g), ptr(lnf_b),
                                   ptr(x_f32), ptr(out), ptr(out_planes), ptr(ws), ws.numel(), _stream()),
          "t4r_gpt2_encoder_fwd")
    return out, out_planes


# --------------------------------------------------------------------------- #
# head
# --------------------------------------------------------------------------- #
def head_softmax_ce(xt_planes, xt_f32, labels, w_planes, w_f32, *, t_dev=None, inv_temperature=1.0, col_bias=None,
                    col_ids=None, hit_value=0.0, pos_logit=None, v_offset=0, want_rank=False, want_loss=True,
                    nprod=3, events=None, label_smoothing=0.0, rank_tgt=None, xt_inv_scale=None, w_inv_scale=None,
                    out_stats: Optional[torch.Tensor] = None, col_ids_sorted_unique: bool = False):
    """Fused logits + log-sum-exp + CE.  Returns dict(row_lse,row_tgt,row_loss,loss,row_rank).
    ``out_stats``: caller-owned fp32 [3, T_cap] buffer that receives row_lse | row_tgt | row_rank (int32 bits) --
    the layout ``peer_combine_lse`` reads from every shard's window."""
    _need_cuda(xt_planes, w_planes)
    lib = _lib.load()
    T_cap = xt_planes.shape[1]
    V = w_planes.shape[1]
    dev = xt_planes.device
    De = xt_f32.shape[1] if xt_f32 is not None else w_f32.shape[1]
    a = _lib.HeadArgs()
    a.T_cap, a.t_dev, a.De, a.V = T_cap, ptr(t_dev), De, V
    a.xt_planes, a.xt_f32, a.labels = ptr(xt_planes), ptr(xt_f32), ptr(labels)
    a.w_planes, a.w_f32 = ptr(w_planes), ptr(w_f32)
    a.inv_temperature = inv_temperature
    a.col_bias, a.col_ids, a.hit_value, a.pos_logit = ptr(col_bias), ptr(col_ids), hit_value, ptr(pos_logit)
    a.v_offset = v_offset
    if out_stats is not None:
        assert out_stats.shape == (3, T_cap) and out_stats.dtype == torch.float32 and out_stats.is_contiguous()
        row_lse, row_tgt = out_stats[0], out_stats[1]
        row_rank = out_stats[2].view(torch.int32) if want_rank else None
    else:
        row_lse = torch.empty(T_cap, dtype=torch.float32, device=dev)
        row_tgt = torch.empty(T_cap, dtype=torch.float32, device=dev)
        row_rank = torch.empty(T_cap, dtype=torch.int32, device=dev) if want_rank else None
    row_loss = torch.empty(T_cap, dtype=torch.float32, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev) if want_loss else None
    a.row_lse, a.row_tgt, a.row_loss, a.loss, a.row_rank = ptr(row_lse), ptr(row_tgt), ptr(row_loss), ptr(loss), ptr(row_rank)
    nbytes = lib.t4r_head_workspace_bytes(T_cap, V, De)
    ws = WS.get("head", nbytes, dev)
    a.workspace, a.workspace_bytes = ptr(ws), ws.numel()
    a.nprod = nprod
    a.label_smoothing = float(label_smoothing)
    a.rank_tgt = ptr(rank_tgt)
    a.col_ids_sorted_unique = 1 if (col_ids is not None and col_ids_sorted_unique) else 0
    if nprod == 2:
        if xt_inv_scale is None or w_inv_scale is None:
            raise _lib.T4RError("head_softmax_ce: nprod=2 needs the mixed planes' inverse row scales")
        a.xt_inv_scale, a.w_inv_scale = ptr(xt_inv_scale), ptr(w_inv_scale)
    ev = events if events is not None else HEAD_EVENTS
    if ev is not None:
        a.ev_gemm_start, a.ev_gemm_stop = ev[0].cuda_event, ev[1].cuda_event
    check(lib.t4r_head_softmax_ce_fwd(C.byref(a), _stream()), "t4r_head_softmax_ce_fwd")
    return {"row_lse": row_lse, "row_tgt": row_tgt, "row_loss": row_loss, "loss": loss, "row_rank": row_rank}


def label_logit(xt_f32, w_f32, labels, *, t_dev=None, class_bias=None, inv_temperature=1.0, v_offset=0):
    _need_cuda(xt_f32, w_f32, labels)
    T_cap, De = xt_f32.shape
    out = torch.empty(T_cap, dtype=torch.float32, device=xt_f32.device)
    check(_lib.load().t4r_label_logit(ptr(xt_f32), ptr(w_f32), ptr(labels), T_cap, ptr(t_dev), De, w_f32.shape[0],
                                      ptr(class_bias), inv_temperature, v_offset, ptr(out), _stream()), "t4r_label_logit")
    return out


def head_logits(xt_planes, w_planes, De: int, *, t_dev=None, inv_temperature=1.0, nprod=3) -> torch.Tensor:
    _need_cuda(xt_planes, w_planes)
    T_cap, V = xt_planes.shape[1], w_planes.shape[1]
    out = torch.zeros((T_cap, V), dtype=torch.float32, device=xt_planes.device)
    check(_lib.load().t4r_head_logits(ptr(xt_planes), ptr(w_planes), T_cap, ptr(t_dev), V, De, inv_temperature,
                                      ptr(out), V, nprod, _stream()), "t4r_head_logits")
    return out


def head_logits_mixed(xt_planes, xt_inv_scale, w_planes, w_inv_scale, De: int, *, t_dev=None,
                      inv_temperature=1.0) -> torch.Tensor:
    """Materialised logits from the operands of the 2-unit product (``split_planes_mixed``)."""
    _need_cuda(xt_planes, w_planes, xt_inv_scale, w_inv_scale)
    T_cap, V = xt_planes.shape[1], w_planes.shape[1]
    out = torch.zeros((T_cap, V), dtype=torch.float32, device=xt_planes.device)
    check(_lib.load().t4r_head_logits_mixed(ptr(xt_planes), ptr(w_planes), T_cap, ptr(t_dev), V, De, inv_temperature,
                                            ptr(out), V, ptr(xt_inv_scale), ptr(w_inv_scale), _stream()),
          "t4r_head_logits_mixed")
    return out


def recall_from_ranks(row_rank: torch.Tensor, ks: Sequence[int], t_dev=None) -> torch.Tensor:
    _need_cuda(row_rank)
    out = torch.empty(len(ks), dtype=torch.float32, device=row_rank.device)
    arr = (C.c_int32 * len(ks))(*[int(k) for k in ks])
    check(_lib.load().t4r_recall_from_ranks(ptr(row_rank), ptr(t_dev), row_rank.numel(), arr, len(ks), ptr(out),
                                            _stream()), "t4r_recall_from_ranks")
    return out


def topk(logits: torch.Tensor, k: int):
    _need_cuda(logits)
    logits = _f32c(logits)
    rows, V = logits.shape
    scores = torch.empty((rows, k), dtype=torch.float32, device=logits.device)
    ids = torch.empty((rows, k), dtype=torch.int64, device=logits.device)
    check(_lib.load().t4r_topk(ptr(logits), rows, V, V, k, ptr(scores), ptr(ids), _stream()), "t4r_topk")
    return scores, ids


def combine_shard_lse(parts: torch.Tensor, t_dev=None):
    """parts [world, T_cap, 2] (lse, label-logit) -> (row_loss [T_cap], loss [1])."""
    _need_cuda(parts)
    parts = _f32c(parts)
    world, T_cap, _ = parts.shape
    row_loss = torch.empty(T_cap, dtype=torch.float32, device=parts.device)
    loss = torch.empty(1, dtype=torch.float32, device=parts.device)
    check(_lib.load().t4r_combine_shard_lse(ptr(parts), world, T_cap, ptr(t_dev), ptr(row_loss), ptr(loss), _stream()),
          "t4r_combine_shard_lse")
    return row_loss, loss


# --------------------------------------------------------------------------- #
# K11 / K12 over NVLink peer memory (csrc/t4r_peer.cu).  ``peers`` arguments are objects with a ``.struct``
# (_lib.PeerPtrs: the same buffer on every rank as addressable from this process) -- distributed.PeerView, or
# ``local_peer_view`` below for single-process use.
# --------------------------------------------------------------------------- #
class LocalPeerView:
    """A "group" whose ranks' buffers all live in THIS process (single-GPU tests / world size 1)."""

    def __init__(self, tensors: Sequence[torch.Tensor], rank: int = 0):
        if not 1 <= len(tensors) <= _lib.T4R_MAX_PEERS:
            raise _lib.T4RError(f"1..{_lib.T4R_MAX_PEERS} ranks")
        _need_cuda(*tensors)
        self.tensors = [t if t.is_contiguous() else t.contiguous() for t in tensors]
        self.world, self.rank = len(tensors), rank
        self.local = self.tensors[rank]
        self.struct = _lib.PeerPtrs()
        self.struct.world, self.struct.rank = self.world, rank
        for r, t in enumerate(self.tensors):
            self.struct.base[r] = t.data_ptr()


def peer_gather_rows(shards, V: int, rows_per_shard: int, K: int, ids: torch.Tensor, count: Optional[torch.Tensor] = None,
                     pad_id: int = -1, want_f32: bool = True, want_planes: bool = True, err_flag=None):
    """Rows ``ids`` of a table row-sharded over the ranks of ``shards`` -> (fp32 [n, K] or None, planes or None)."""
    _need_cuda(ids, count)
    ids = ids.reshape(-1)
    if ids.dtype != torch.int64:
        ids = ids.long()
    ids = ids.contiguous()
    n, dev = ids.numel(), ids.device
    of = torch.empty((n, K), dtype=torch.float32, device=dev) if want_f32 else None
    planes = torch.empty((2, n, round_up64(K)), dtype=torch.bfloat16, device=dev) if want_planes else None
    if n == 0:
        return of, planes
    check(_lib.load().t4r_peer_gather_rows(C.byref(shards.struct), int(V), int(rows_per_shard), int(K), ptr(ids), ptr(count),
                                           n, int(pad_id), ptr(of), ptr(planes), ptr(err_flag), _stream()),
          "t4r_peer_gather_rows")
    return of, planes


def peer_pull_rows(mail_x, mail_y, counts: torch.Tensor, cap: int, K: int, want_f32: bool = True):
    """Label rows of every rank, rank-major and compact -> dict(x fp32 [world*cap, K] or None, planes, labels,
    t_total int32[1], my_start int32[1])."""
    _need_cuda(counts)
    assert counts.dtype == torch.int32 and counts.numel() == mail_x.world
    dev, cap_g = counts.device, mail_x.world * cap
    of = torch.empty((cap_g, K), dtype=torch.float32, device=dev) if want_f32 else None
    planes = torch.empty((2, cap_g, round_up64(K)), dtype=torch.bfloat16, device=dev)
    labels = torch.empty(cap_g, dtype=torch.int64, device=dev)
    meta = torch.empty(2, dtype=torch.int32, device=dev)
    check(_lib.load().t4r_peer_pull_rows(C.byref(mail_x.struct), C.byref(mail_y.struct), ptr(counts), int(cap), int(K),
                                         ptr(of), ptr(planes), ptr(labels), ptr(meta[0:1]), ptr(meta[1:2]), _stream()),
          "t4r_peer_pull_rows")
    return {"x": of, "planes": planes, "labels": labels, "t_total": meta[0:1], "my_start": meta[1:2]}


def peer_combine_lse(stats, cap_g: int, t_total: torch.Tensor, with_rank: bool = False):
    """Every shard's [3, cap_g] statistics -> (row_loss [cap_g], loss [1], row_rank int32 [cap_g] or None)."""
    _need_cuda(t_total)
    dev = t_total.device
    row_loss = torch.empty(cap_g, dtype=torch.float32, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    row_rank = torch.empty(cap_g, dtype=torch.int32, device=dev) if with_rank else None
    check(_lib.load().t4r_peer_combine_lse(C.byref(stats.struct), int(cap_g), ptr(t_total), int(bool(with_rank)),
                                           ptr(row_loss), ptr(row_rank), ptr(loss), _stream()), "t4r_peer_combine_lse")
    return row_loss, loss, row_rank


# --------------------------------------------------------------------------- #
# N3: primitives of the training step (csrc/t4r_train.cu; composed in training.py).
# Public functions take CUDA tensors only.  ``host_twin(name)`` returns the same entry point running its per-item code
# in a host loop on CPU tensors -- test infrastructure (tests/test_abi_and_host.py), never used by the package.
# --------------------------------------------------------------------------- #
def _tr(on_host, *ts):
    if on_host:
        assert all(t is None or not t.is_cuda for t in ts)
        return (None, 1)
    _need_cuda(*ts)
    return (_stream(), 0)


def transpose(x, _on_host=False):
    x = _f32c(x)
    tail = _tr(_on_host, x)
    R, Cc = x.shape
    out = torch.empty((Cc, R), dtype=torch.float32, device=x.device)
    check(_lib.load().t4r_train_transpose(ptr(x), R, Cc, ptr(out), *tail), "t4r_train_transpose")
    return out


def act_fwd(kind, x, _on_host=False):
    x = _f32c(x)
    tail = _tr(_on_host, x)
    y = torch.empty_like(x)
    check(_lib.load().t4r_train_act_fwd(int(kind), ptr(x), ptr(y), x.numel(), *tail), "t4r_train_act_fwd")
    return y


def act_bwd(kind, pre, dy, _on_host=False):
    pre, dy = _f32c(pre), _f32c(dy)
    tail = _tr(_on_host, pre, dy)
    dx = torch.empty_like(pre)
    check(_lib.load().t4r_train_act_bwd(int(kind), ptr(pre), ptr(dy), ptr(dx), pre.numel(), *tail), "t4r_train_act_bwd")
    return dx


def add_positions(x, wpe, B, L, _on_host=False):
    x, wpe = _f32c(x), _f32c(wpe)
    tail = _tr(_on_host, x, wpe)
    d = x.shape[1]
    y = torch.empty_like(x)
    check(_lib.load().t4r_train_add_positions(ptr(x), ptr(wpe), B, L, d, ptr(y), *tail), "t4r_train_add_positions")
    return y


def sum_over_sessions(x, B, L, _on_host=False):
    x = _f32c(x)
    tail = _tr(_on_host, x)
    d = x.shape[1]
    out = torch.empty((L, d), dtype=torch.float32, device=x.device)
    check(_lib.load().t4r_train_sum_sessions(ptr(x), B, L, d, ptr(out), *tail), "t4r_train_sum_sessions")
    return out


def apply_row_codes(y, code, mask_vec, _on_host=False):
    y, mask_vec = _f32c(y), _f32c(mask_vec)
    code = code.reshape(-1).to(torch.uint8).contiguous()
    tail = _tr(_on_host, y, code, mask_vec)
    out = torch.empty_like(y)
    check(_lib.load().t4r_train_row_codes_fwd(ptr(y), ptr(code), ptr(mask_vec), y.shape[0], y.shape[1], ptr(out), *tail),
          "t4r_train_row_codes_fwd")
    return out


def row_codes_bwd(dx, code, _on_host=False):
    dx = _f32c(dx)
    code = code.reshape(-1).to(torch.uint8).contiguous()
    tail = _tr(_on_host, dx, code)
    dy = torch.empty_like(dx)
    tmp = torch.empty_like(dx)
    dmask = torch.empty((dx.shape[1],), dtype=torch.float32, device=dx.device)
    check(_lib.load().t4r_train_row_codes_bwd(ptr(dx), ptr(code), dx.shape[0], dx.shape[1], ptr(dy), ptr(dmask), ptr(tmp),
                                              *tail), "t4r_train_row_codes_bwd")
    return dmask, dy


def gather_rows(x, idx, _on_host=False):
    x = _f32c(x)
    idx = idx.to(torch.int32).contiguous()
    tail = _tr(_on_host, x, idx)
    out = torch.empty((idx.numel(), x.shape[1]), dtype=torch.float32, device=x.device)
    check(_lib.load().t4r_train_gather_rows(ptr(x), ptr(idx), idx.numel(), x.shape[1], ptr(out), *tail), "t4r_train_gather_rows")
    return out


def scatter_rows(src, idx, n_rows, _on_host=False):
    src = _f32c(src)
    idx = idx.to(torch.int32).contiguous()
    tail = _tr(_on_host, src, idx)
    out = torch.empty((n_rows, src.shape[1]), dtype=torch.float32, device=src.device)
    check(_lib.load().t4r_train_scatter_rows(ptr(src), ptr(idx), idx.numel(), src.shape[1], n_rows, ptr(out), *tail),
          "t4r_train_scatter_rows")
    return out


def softmax_ce_bwd(z, row_lse, labels, v0, scale, label_smoothing=0.0, V_total=None, _on_host=False):
    """In place on ``z`` [T, Vc]: (softmax - target distribution) * scale; returns ``z``."""
    assert z.dtype == torch.float32 and z.is_contiguous()
    row_lse, labels = _f32c(row_lse), labels.long().contiguous()
    tail = _tr(_on_host, z, row_lse, labels)
    T, Vc = z.shape
    check(_lib.load().t4r_train_softmax_ce_bwd(ptr(z), ptr(row_lse), ptr(labels), T, Vc, int(v0), float(scale),
                                               float(label_smoothing), int(V_total if V_total is not None else Vc), *tail),
          "t4r_train_softmax_ce_bwd")
    return z


def sampled_ce_bwd(z, row_lse, labels, col_bias, col_ids, inv_tau, scale, _on_host=False):
    """In place on ``z`` [T, S] (= x . w_s / tau of the sampled negatives): exp(logit - lse) * scale, 0 at accidental hits."""
    assert z.dtype == torch.float32 and z.is_contiguous()
    row_lse, col_bias = _f32c(row_lse), _f32c(col_bias)
    labels, col_ids = labels.long().contiguous(), col_ids.long().contiguous()
    tail = _tr(_on_host, z, row_lse, labels, col_bias, col_ids)
    T, S = z.shape
    check(_lib.load().t4r_train_sampled_ce_bwd(ptr(z), ptr(row_lse), ptr(labels), ptr(col_bias), ptr(col_ids), T, S,
                                               float(inv_tau), float(scale), *tail), "t4r_train_sampled_ce_bwd")
    return z


def index_add_rows(dst, idx, src, col, width, skip_index=None, _on_host=False):
    assert dst.dtype == torch.float32 and dst.is_contiguous() and dst.shape[1] == width
    src, idx = _f32c(src), idx.long().contiguous()
    if idx.numel() != src.shape[0]:
        raise _lib.T4RError(f"index_add_rows: {idx.numel()} indices for {src.shape[0]} source rows")
    tail = _tr(_on_host, dst, idx, src)
    check(_lib.load().t4r_train_index_add_rows(ptr(dst), ptr(idx), ptr(src), idx.numel(), src.shape[1], int(col), int(width),
                                               -1 if skip_index is None else int(skip_index), *tail),
          "t4r_train_index_add_rows")
    return dst


def soft_emb_fwd(x, w, b, table, _on_host=False):
    """x [M] scalars, w / b [n] (``Linear(1, n)``), table [n, dim] -> (out [M, dim], p [M, n] softmax weights)."""
    x, w, b, table = _f32c(x.reshape(-1)), _f32c(w.reshape(-1)), _f32c(b.reshape(-1)), _f32c(table)
    tail = _tr(_on_host, x, w, b, table)
    M, (n, dim) = x.numel(), table.shape
    p = torch.empty((M, n), dtype=torch.float32, device=x.device)
    out = torch.empty((M, dim), dtype=torch.float32, device=x.device)
    check(_lib.load().t4r_train_soft_emb_fwd(ptr(x), ptr(w), ptr(b), ptr(table), M, n, dim, ptr(p), ptr(out), *tail),
          "t4r_train_soft_emb_fwd")
    return out, p


def soft_emb_bwd(x, table, p, dout, _on_host=False):
    """-> (dlogit [M, n], dlogit * x [M, n])"""
    x, table, p, dout = _f32c(x.reshape(-1)), _f32c(table), _f32c(p), _f32c(dout)
    tail = _tr(_on_host, x, table, p, dout)
    M, (n, dim) = x.numel(), table.shape
    dl = torch.empty((M, n), dtype=torch.float32, device=x.device)
    dlx = torch.empty((M, n), dtype=torch.float32, device=x.device)
    check(_lib.load().t4r_train_soft_emb_bwd(ptr(x), ptr(table), ptr(p), ptr(dout), M, n, dim, ptr(dl), ptr(dlx), *tail),
          "t4r_train_soft_emb_bwd")
    return dl, dlx


def _binary(op, a, b, _on_host):
    a, b = _f32c(a), _f32c(b)
    assert a.shape == b.shape
    tail = _tr(_on_host, a, b)
    out = torch.empty_like(a)
    check(_lib.load().t4r_train_binary(op, ptr(a), ptr(b), ptr(out), a.numel(), *tail), "t4r_train_binary")
    return out


def ew_add(a, b, _on_host=False):
    return _binary(0, a, b, _on_host)


def ew_mul(a, b, _on_host=False):
    return _binary(1, a, b, _on_host)


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, _on_host=False):
    """One AdamW update in place on ``p`` / ``m`` / ``v`` (flat fp32, contiguous)."""
    for t in (p, g, m, v):
        assert t.dtype == torch.float32 and t.is_contiguous()
    tail = _tr(_on_host, p, g, m, v)
    check(_lib.load().t4r_train_adamw(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), float(lr), float(beta1), float(beta2),
                                      float(eps), float(weight_decay), int(step), *tail), "t4r_train_adamw")
    return p


def col_sum(x, _on_host=False):
    x = _f32c(x)
    tail = _tr(_on_host, x)
    out = torch.empty((x.shape[1],), dtype=torch.float32, device=x.device)
    check(_lib.load().t4r_train_col_sum(ptr(x), x.shape[0], x.shape[1], ptr(out), *tail), "t4r_train_col_sum")
    return out


def layer_norm_fwd(x, gamma, beta, eps, _on_host=False):
    x, gamma, beta = _f32c(x), _f32c(gamma), _f32c(beta)
    tail = _tr(_on_host, x, gamma, beta)
    y = torch.empty_like(x)
    check(_lib.load().t4r_train_layer_norm_fwd(ptr(x), ptr(gamma), ptr(beta), x.shape[0], x.shape[1], float(eps), ptr(y),
                                               *tail), "t4r_train_layer_norm_fwd")
    return y


def layer_norm_bwd(x_pre, gamma, eps, dy, add=None, _on_host=False):
    x_pre, gamma, dy = _f32c(x_pre), _f32c(gamma), _f32c(dy)
    add = _f32c(add) if add is not None else None
    tail = _tr(_on_host, x_pre, gamma, dy, add)
    d = x_pre.shape[1]
    dx = torch.empty_like(x_pre)
    tmp = torch.empty_like(x_pre)
    dg = torch.empty((d,), dtype=torch.float32, device=x_pre.device)
    db = torch.empty((d,), dtype=torch.float32, device=x_pre.device)
    check(_lib.load().t4r_train_layer_norm_bwd(ptr(x_pre), ptr(gamma), x_pre.shape[0], d, float(eps), ptr(dy), ptr(add),
                                               ptr(dx), ptr(dg), ptr(db), ptr(tmp), *tail), "t4r_train_layer_norm_bwd")
    return dx, dg, db


def dropout(x, p: float, seed: int, site: int, _on_host=False):
    """x * keep / (1 - p) with the counter-based mask of (seed, site) (t4r_train_dropout); applying it to a gradient
    with the same (p, seed, site) is the backward.  p == 0 returns x itself."""
    if p <= 0.0:
        return x
    x = _f32c(x)
    tail = _tr(_on_host, x)
    y = torch.empty_like(x)
    check(_lib.load().t4r_train_dropout(ptr(x), ptr(y), x.numel(), float(p), int(seed), int(site), *tail), "t4r_train_dropout")
    return y


def attn_drop_fwd(qkv, R, rw, rr, B, L, H, drop, plm_mask=None, _on_host=False):
    """Attention forward of the training graph with dropout of the probabilities: ``drop`` = (p, seed, site).
    R / rw / rr None selects GPT-2's causal form; ``plm_mask`` the two-stream form.  -> fp32 [rows of qkv, d]."""
    qkv = _f32c(qkv)
    R, rw, rr = ((_f32c(t) if t is not None else None) for t in (R, rw, rr))
    plm_mask = plm_mask.to(torch.uint8).contiguous() if plm_mask is not None else None
    tail = _tr(_on_host, qkv, R, rw, rr, plm_mask)
    d = qkv.shape[1] // 3
    out = torch.empty((qkv.shape[0], d), dtype=torch.float32, device=qkv.device)
    p, seed, site = drop
    check(_lib.load().t4r_train_attn_drop_fwd(ptr(qkv), ptr(R), ptr(rw), ptr(rr), B, L, d, H, ptr(plm_mask), float(p),
                                              int(seed), int(site), ptr(out), *tail), "t4r_train_attn_drop_fwd")
    return out


def xlnet_attn_bwd(qkv, R, rw, rr, dout, B, L, H, plm_mask=None, drop=None, _on_host=False):
    """``plm_mask`` [B, L, L] uint8: the two-stream form (qkv / dout hold 2 B L rows: content stream, then query stream).
    ``drop`` = (p, seed, site): the forward dropped its probabilities with that mask."""
    qkv, R, rw, rr, dout = (_f32c(t) for t in (qkv, R, rw, rr, dout))
    plm_mask = plm_mask.to(torch.uint8).contiguous() if plm_mask is not None else None
    tail = _tr(_on_host, qkv, R, rw, rr, dout, plm_mask)
    d = dout.shape[1]
    dev = qkv.device
    dqkv = torch.empty_like(qkv)
    dR = torch.empty((2 * L, d), dtype=torch.float32, device=dev)
    drw = torch.empty((d,), dtype=torch.float32, device=dev)
    drr = torch.empty((d,), dtype=torch.float32, device=dev)
    part = torch.empty((B * (2 * L + 2) * d,), dtype=torch.float32, device=dev)
    if drop is not None and drop[0] > 0.0:
        check(_lib.load().t4r_train_attn_drop_bwd(ptr(qkv), ptr(R), ptr(rw), ptr(rr), ptr(dout), B, L, d, H, ptr(dqkv),
                                                  ptr(dR), ptr(drw), ptr(drr), ptr(part), ptr(plm_mask), float(drop[0]),
                                                  int(drop[1]), int(drop[2]), *tail), "t4r_train_attn_drop_bwd")
    else:
        check(_lib.load().t4r_train_attn_bwd(ptr(qkv), ptr(R), ptr(rw), ptr(rr), ptr(dout), B, L, d, H, ptr(dqkv), ptr(dR),
                                             ptr(drw), ptr(drr), ptr(part), ptr(plm_mask), *tail), "t4r_train_attn_bwd")
    return dqkv, dR, drw, drr


def causal_attn_bwd(qkv, dout, B, L, H, drop=None, _on_host=False):
    qkv, dout = _f32c(qkv), _f32c(dout)
    tail = _tr(_on_host, qkv, dout)
    dqkv = torch.empty_like(qkv)
    if drop is not None and drop[0] > 0.0:
        check(_lib.load().t4r_train_attn_drop_bwd(ptr(qkv), None, None, None, ptr(dout), B, L, dout.shape[1], H, ptr(dqkv),
                                                  None, None, None, None, None, float(drop[0]), int(drop[1]), int(drop[2]),
                                                  *tail), "t4r_train_attn_drop_bwd")
    else:
        check(_lib.load().t4r_train_attn_bwd(ptr(qkv), None, None, None, ptr(dout), B, L, dout.shape[1], H, ptr(dqkv), None,
                                             None, None, None, None, *tail), "t4r_train_attn_bwd")
    return dqkv


def _planes_to_f32(planes, d):
    return planes[0, :, :d].float() + planes[1, :, :d].float()


def xlnet_attn_fwd(qkv, R, rw, rr, B, L, H):
    qkv, R, rw, rr = (_f32c(t) for t in (qkv, R, rw, rr))
    _need_cuda(qkv, R, rw, rr)
    d = qkv.shape[1] // 3
    out = torch.empty((2, B * L, d), dtype=torch.bfloat16, device=qkv.device)
    check(_lib.load().t4r_train_xlnet_attn_fwd(ptr(qkv), ptr(R), ptr(rw), ptr(rr), B, L, d, H, ptr(out), _stream()),
          "t4r_train_xlnet_attn_fwd")
    return _planes_to_f32(out, d)


def xlnet_attn_plm_fwd(qkv, R, rw, rr, B, L, H, plm_mask):
    """Two-stream (PLM) attention forward: qkv fp32 [2 B L, 3d] (content stream rows, then query stream rows)."""
    qkv, R, rw, rr = (_f32c(t) for t in (qkv, R, rw, rr))
    pm = plm_mask.to(torch.uint8).contiguous()
    _need_cuda(qkv, R, rw, rr, pm)
    d = qkv.shape[1] // 3
    if (3 * d) % 64 or d % 64:
        raise _lib.T4RError("xlnet_attn_plm_fwd: d_model must be a multiple of 64 (plane rows are not padded)")
    out = torch.empty((2, 2 * B * L, d), dtype=torch.bfloat16, device=qkv.device)
    qkv_p, r_p = split_planes(qkv), split_planes(R)          # kept alive across the launch
    check(_lib.load().t4r_train_xlnet_attn_plm_fwd(ptr(qkv_p), ptr(r_p), ptr(rw), ptr(rr), B, L, d, H, ptr(pm), ptr(out),
                                                   _stream()), "t4r_train_xlnet_attn_plm_fwd")
    return _planes_to_f32(out, d)


def causal_attn_fwd(qkv, B, L, H):
    qkv = _f32c(qkv)
    _need_cuda(qkv)
    d = qkv.shape[1] // 3
    out = torch.empty((2, B * L, d), dtype=torch.bfloat16, device=qkv.device)
    check(_lib.load().t4r_train_causal_attn_fwd(ptr(qkv), B, L, d, H, ptr(out), _stream()), "t4r_train_causal_attn_fwd")
    return _planes_to_f32(out, d)


def rel_pos_proj(wr_list, L, d):
    """R_l = pos(L, d) @ Wr_l for every layer -> [n_layer, 2L, d] (the inference path's positional kernel)."""
    ws = [_f32c(w) for w in wr_list]
    _need_cuda(*ws)
    out = torch.empty((len(ws), 2 * L, d), dtype=torch.float32, device=ws[0].device)
    arr = (C.c_void_p * len(ws))(*[w.data_ptr() for w in ws])
    check(_lib.load().t4r_train_rel_pos_proj(arr, len(ws), L, d, ptr(out), _stream()), "t4r_train_rel_pos_proj")
    return out


def rel_pos_table(L, d, device=None):
    """XLNet's relative position table [2L, d] (HF:xlnet:930-976, sin || cos of positions L .. -L+1): a constant of
    (L, d), built once with torch."""
    freq = torch.arange(0, d, 2.0, dtype=torch.float32, device=device)
    inv_freq = 1.0 / torch.pow(10000, freq / d)
    pos = torch.arange(L, -L, -1.0, dtype=torch.float32, device=device)
    sinusoid = torch.einsum("i,d->id", pos, inv_freq)
    return torch.cat([torch.sin(sinusoid), torch.cos(sinusoid)], dim=-1)


def host_twin(name: str):
    """The primitive ``name`` with its per-item code run in a host loop on CPU tensors (test infrastructure)."""
    fn = globals()[name]
    return lambda *a, **k: fn(*a, _on_host=True, **k)
