"""Row-sharded item table + sharded tied-weight head (BASELINE configs 4-5).

The reference replicates the table and the logit GEMM on every rank (DDP only,
SURVEY §2.1); at 10-50 M items both stop fitting, and under weight tying they are
the same tensor.  It is sharded row-wise once and used from both ends with one
collective per end:

* lookup (K11): ranks all-gather their item ids (8 B/id), every rank gathers the rows
  it owns for every peer (``t4r_gather_rows_split_i64``) and ONE all-to-all returns
  them to the sessions' owners; a second ``t4r_gather_rows_split`` call un-permutes the
  received rows and emits the split planes the projection GEMM consumes.
* head (K12): label rows are all-gathered, every rank runs the fused logits+LSE
  kernel over its V/world rows (``v_offset``), and ONE all-gather of the per-row
  ``(lse, label-logit)`` pairs (8 B/row) is combined by ``t4r_combine_shard_lse``.

Everything else (projection, masking, encoder) is data parallel with no collective.
``torch.distributed`` (NCCL on GPUs) is the plumbing; the routing arithmetic below is
integer index bookkeeping shared by all backends.  The local compute steps are
injectable (``gather_rows`` / ``head_rows``): the product default is the CUDA kernels;
the world_size-2 ``gloo`` tests on CPU pass test-side stand-ins to exercise the
collective choreography (there is no CPU fallback in the product).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(V: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [lo, hi) of the table owned by ``rank`` (block partition, ceil division)."""
    per = (V + world - 1) // world
    lo = min(V, rank * per)
    return lo, min(V, lo + per)


def owner_of(ids: torch.Tensor, V: int, world: int) -> torch.Tensor:
    per = (V + world - 1) // world
    return torch.div(ids, per, rounding_mode="floor").clamp_(max=world - 1)


def _all_gather(t: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """all_gather_into_tensor with a flat output buffer (the form every backend accepts);
    returns [world, *t.shape]."""
    t = t.contiguous()
    out = torch.empty(world * t.numel(), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.reshape(-1), group=group)
    return out.view((world,) + tuple(t.shape))


class LookupPlan:
    """Index bookkeeping for one sharded lookup step (pure integer work)."""

    def __init__(self, ids_all: torch.Tensor, rank: int, world: int, V: int):
        # ids_all [world, n]: flattened item ids of every rank
        self.world, self.rank = world, rank
        lo, hi = shard_bounds(V, rank, world)
        n = ids_all.shape[1]
        mine = (ids_all >= lo) & (ids_all < hi)                       # [world, n] rows I own, per requester
        self.send_counts: List[int] = mine.sum(dim=1).tolist()         # rows I send to each peer
        # local row index (within my shard) of every row I send, requester-major, position-ascending
        self.send_local_idx = (ids_all[mine] - lo).contiguous()
        own = owner_of(ids_all[rank], V, world)                        # owner of each of MY positions
        self.recv_counts: List[int] = [int((own == q).sum()) for q in range(world)]
        # received rows arrive owner-major, position-ascending: position p sits at perm_inv[p]
        order = torch.argsort(own, stable=True)                        # positions grouped by owner
        inv = torch.empty_like(order)
        inv[order] = torch.arange(n, device=order.device)
        self.unpermute = inv.to(torch.int32)                           # out[p] = recv[unpermute[p]]


def _default_gather_rows(table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    from . import ops
    if idx.numel() == 0:
        return torch.empty((0, table.shape[1]), dtype=torch.float32, device=table.device)
    _, rows = ops.gather_rows_split(table, idx, None, idx.numel(), want_f32=True)
    return rows


def _default_place(recv: torch.Tensor, unpermute: torch.Tensor):
    from . import ops
    planes, rows = ops.gather_rows_split(recv, unpermute, None, unpermute.numel(), want_f32=True)
    return rows, planes


def sharded_embedding_lookup(local_table: torch.Tensor, ids: torch.Tensor, V: int, group=None,
                             gather_rows: Optional[Callable] = None, place: Optional[Callable] = None,
                             plan_out: Optional[list] = None):
    """Embedding rows for ``ids`` [B, L] from a table whose rows are block-sharded over
    the group.  Returns (rows fp32 [B*L, De], planes or None)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    gather_rows = gather_rows or _default_gather_rows
    place = place or _default_place
    flat = ids.reshape(-1).long().contiguous()
    ids_all = _all_gather(flat, world, group)
    plan = LookupPlan(ids_all, rank, world, V)
    send = gather_rows(local_table, plan.send_local_idx)
    De = local_table.shape[1]
    recv = torch.empty((sum(plan.recv_counts), De), dtype=torch.float32, device=flat.device)
    dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=plan.recv_counts,
                           input_split_sizes=plan.send_counts, group=group)
    if plan_out is not None:
        plan_out.append(plan)
    return place(recv, plan.unpermute)


def sharded_embedding_lookup_bwd(d_rows: torch.Tensor, plan: "LookupPlan", local_grad: torch.Tensor, group=None,
                                 skip_local_index: int = -1, scatter_rows: Optional[Callable] = None,
                                 index_add: Optional[Callable] = None):
    """Backward of ``sharded_embedding_lookup`` ("sharded-table grads via the transposed all-to-all"):
    ``d_rows`` [n, De] are the gradients of the rows this rank looked up (position order).  They are put back into
    arrival order (the inverse of the un-permutation), ONE all-to-all with the forward's split sizes swapped returns them
    to the rows' owners, and each owner scatter-adds what it receives into ``local_grad`` [rows of its shard, De] at the
    local indices it had served.  ``skip_local_index``: local index of the padding row (no gradient), -1 if not here."""
    if scatter_rows is None or index_add is None:
        from . import ops
        scatter_rows = scatter_rows or (lambda src, idx, n: ops.scatter_rows(src, idx, n))
        index_add = index_add or (lambda dst, idx, src, skip: ops.index_add_rows(dst, idx, src, 0, dst.shape[1],
                                                                                skip_index=skip if skip >= 0 else None))
    n, De = d_rows.shape
    d_recv = scatter_rows(d_rows.contiguous(), plan.unpermute, n)              # d_recv[unpermute[p]] = d_rows[p]
    d_send = torch.empty((sum(plan.send_counts), De), dtype=torch.float32, device=d_rows.device)
    dist.all_to_all_single(d_send, d_recv.contiguous(), output_split_sizes=plan.send_counts,
                           input_split_sizes=plan.recv_counts, group=group)
    if d_send.shape[0]:
        index_add(local_grad, plan.send_local_idx, d_send, skip_local_index)
    return local_grad


def _default_head_rows(xt: torch.Tensor, labels: torch.Tensor, local_table: torch.Tensor, w_planes, v_offset: int,
                       inv_tau: float, rank_tgt: Optional[torch.Tensor] = None):
    """-> [T, 2] (lse over this shard, label logit if the label lives here else 0); with ``rank_tgt`` (the
    label's logit over the whole table) also the per-row count of this shard's classes scoring above it."""
    from . import ops
    if w_planes is None:             # training: the table changes every step, its planes are made on the spot
        w_planes = ops.split_planes(local_table)
    if isinstance(w_planes, tuple):  # (mixed planes, inverse row scales): the 2-unit product (ops.split_planes_mixed)
        xp, xi = ops.split_planes_mixed(xt)
        res = ops.head_softmax_ce(xp, xt, labels, w_planes[0], local_table, inv_temperature=inv_tau, v_offset=v_offset,
                                  want_loss=False, want_rank=rank_tgt is not None, rank_tgt=rank_tgt, nprod=2,
                                  xt_inv_scale=xi, w_inv_scale=w_planes[1])
    else:
        xp = ops.split_planes(xt)
        res = ops.head_softmax_ce(xp, xt, labels, w_planes, local_table, inv_temperature=inv_tau, v_offset=v_offset,
                                  want_loss=False, want_rank=rank_tgt is not None, rank_tgt=rank_tgt)
    part = torch.stack([res["row_lse"], res["row_tgt"]], dim=1)
    return part if rank_tgt is None else (part, res["row_rank"])


def _default_label_logit(xt: torch.Tensor, labels: torch.Tensor, local_table: torch.Tensor, v_offset: int, inv_tau: float):
    from . import ops
    return ops.label_logit(xt, local_table, labels, inv_temperature=inv_tau, v_offset=v_offset)


def _default_combine(parts: torch.Tensor):
    from . import ops
    return ops.combine_shard_lse(parts)


def sharded_softmax_ce(xt: torch.Tensor, labels: torch.Tensor, local_table: torch.Tensor, V: int, group=None,
                       w_planes=None, inv_tau: float = 1.0, head_rows: Optional[Callable] = None,
                       combine: Optional[Callable] = None, want_rank: bool = False,
                       label_logit: Optional[Callable] = None):
    """Full-softmax CE of the label rows of ALL ranks against a row-sharded output
    table.  ``xt`` [T_local, De] / ``labels`` [T_local] are this rank's label rows.
    Returns (row_loss for this rank's rows, global mean loss, T_total) and, with ``want_rank``
    (evaluation), a 4th item: the rank of each of this rank's labels over the WHOLE table
    (two more small collectives: the label logits are summed over shards first -- only the
    owner contributes -- then the per-shard counts of higher-scoring classes are summed)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    head_rows = head_rows or _default_head_rows
    combine = combine or _default_combine
    dev = xt.device
    counts_l = _all_gather(torch.tensor([xt.shape[0]], dtype=torch.int64, device=dev), world, group).reshape(-1).tolist()
    T_max = max(counts_l)
    De = xt.shape[1]
    pad_x = torch.zeros((T_max, De), dtype=torch.float32, device=dev)
    pad_x[: xt.shape[0]] = xt
    pad_y = torch.zeros((T_max,), dtype=torch.int64, device=dev)
    pad_y[: xt.shape[0]] = labels
    all_x = _all_gather(pad_x, world, group)
    all_y = _all_gather(pad_y, world, group)
    xg = torch.cat([all_x[r, : counts_l[r]] for r in range(world)], dim=0).contiguous()
    yg = torch.cat([all_y[r, : counts_l[r]] for r in range(world)], dim=0).contiguous()
    lo, _ = shard_bounds(V, rank, world)
    start = sum(counts_l[:rank])
    mine = slice(start, start + counts_l[rank])
    if want_rank:
        label_logit = label_logit or _default_label_logit
        tgt = label_logit(xg, yg, local_table, lo, inv_tau).contiguous()  # 0 where the label is not mine
        dist.all_reduce(tgt, group=group)
        part, cnt = head_rows(xg, yg, local_table, w_planes, lo, inv_tau, tgt)
        cnt = cnt.contiguous()
        dist.all_reduce(cnt, group=group)
    else:
        part = head_rows(xg, yg, local_table, w_planes, lo, inv_tau)      # [T_total, 2]
    parts = _all_gather(part, world, group)                               # the one head collective
    row_loss, loss = combine(parts)
    if want_rank:
        return row_loss[mine], loss, int(sum(counts_l)), cnt[mine]
    return row_loss[mine], loss, int(sum(counts_l))


def _default_local_topk(x_all: torch.Tensor, local_table: torch.Tensor, w_planes, inv_tau: float, k: int,
                        max_score_bytes: int = 4 << 30):
    """(scores, local row ids) of the k best rows of THIS shard for every session row; the [rows, V_local]
    scores are materialised a block of session rows at a time (<= max_score_bytes) by the tensor-core GEMM."""
    from . import ops
    if w_planes is None:
        w_planes = ops.split_planes(local_table)
    n, De = x_all.shape
    v_loc = local_table.shape[0]
    step = max(1, min(n, max_score_bytes // (4 * max(v_loc, 1))))
    sc, ids = [], []
    for r0 in range(0, n, step):
        xp = ops.split_planes(x_all[r0:r0 + step])
        scores = ops.head_logits(xp, w_planes, De, inv_temperature=inv_tau)
        a, b = ops.topk(scores, k)
        sc.append(a)
        ids.append(b)
    return torch.cat(sc), torch.cat(ids)


def _default_merge_topk(cand_scores: torch.Tensor, k: int):
    from . import ops
    return ops.topk(cand_scores, k)


def sharded_topk(xs: torch.Tensor, local_table: torch.Tensor, V: int, k: int, group=None, w_planes=None,
                 inv_tau: float = 1.0, local_topk: Optional[Callable] = None, merge: Optional[Callable] = None):
    """Top-k items of every one of this rank's sessions over a row-sharded output table (serving over
    BASELINE configs 4-5; the reference's ``top_k`` path, model/prediction_task.py:452-470, on a replicated table).

    ``xs`` [B, De]: hidden row at the next-item position of each of THIS rank's sessions (same B on every rank).
    One all-gather of the rows (B*De*4 bytes per rank), every rank scores all sessions against its V/world
    rows and keeps its k best per session, ONE all-to-all returns the (score, global id) candidates to the
    sessions' owner, which merges world*k candidates.  Order: score descending, ties by lower item id -- the
    order of ``t4r_topk`` on a replicated table (candidates arrive shard-major, i.e. id-ascending among ties).
    Returns (scores [B, k] fp32, ids [B, k] int64)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    local_topk = local_topk or _default_local_topk
    merge = merge or _default_merge_topk
    B, De = xs.shape
    lo, hi = shard_bounds(V, rank, world)
    if not 1 <= k <= V:
        raise ValueError(f"top_k={k} must be in [1, {V}]")
    x_all = _all_gather(xs.float().contiguous(), world, group).reshape(world * B, De)
    k_loc = min(k, hi - lo)
    neg_inf = float("-inf")
    if k_loc > 0:
        sc, ids = local_topk(x_all, local_table, w_planes, inv_tau, k_loc)
        ids = ids.long() + lo
    else:  # more ranks than rows: this shard is empty
        sc = torch.empty((world * B, 0), dtype=torch.float32, device=xs.device)
        ids = torch.empty((world * B, 0), dtype=torch.int64, device=xs.device)
    if k_loc < k:  # pad so that every shard sends k candidates (-inf never wins against a real score)
        sc = torch.cat([sc, sc.new_full((world * B, k - k_loc), neg_inf)], dim=1)
        ids = torch.cat([ids, ids.new_full((world * B, k - k_loc), -1)], dim=1)
    # one exchange: score bits and ids travel in the same int64 buffer
    send = torch.stack([sc.contiguous().view(torch.int32).long(), ids], dim=-1).contiguous()  # [world*B, k, 2]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)                                         # recv[s*B + b] = shard s, my session b
    recv = recv.view(world, B, k, 2)
    cand_sc = recv[..., 0].to(torch.int32).view(torch.float32).permute(1, 0, 2).reshape(B, world * k).contiguous()
    cand_id = recv[..., 1].permute(1, 0, 2).reshape(B, world * k).contiguous()
    top_sc, pos = merge(cand_sc, k)
    return top_sc, cand_id.gather(1, pos.long())


def _default_train_head(xg, yg, local_table, lo, inv_tau, lse_global, scale, head_chunk):
    """This shard's part of the head backward for ALL label rows: (dX partial [T, De], dW_local [rows, De])."""
    from . import ops
    from .training import gemm_nt
    Vl, De = local_table.shape
    dx = torch.zeros_like(xg)
    dW = torch.zeros_like(local_table)
    xg_planes, xg_t = ops.split_planes(xg), ops.transpose(xg)
    for v0 in range(0, Vl, head_chunk):
        v1 = min(Vl, v0 + head_chunk)
        Wc = local_table[v0:v1].contiguous()
        z = ops.head_logits(xg_planes, ops.split_planes(Wc), De, inv_temperature=inv_tau)
        P = ops.softmax_ce_bwd(z, lse_global, yg, lo + v0, scale * inv_tau)
        dx = gemm_nt(P, ops.transpose(Wc), residual=dx)
        dW[v0:v1] = gemm_nt(ops.transpose(P), xg_t)
    return dx, dW


def sharded_softmax_ce_train(xt: torch.Tensor, labels: torch.Tensor, local_table: torch.Tensor, V: int, group=None,
                             w_planes=None, inv_tau: float = 1.0, head_chunk: int = 32768,
                             head_rows: Optional[Callable] = None, train_head: Optional[Callable] = None):
    """Forward AND backward of the row-sharded full-softmax head for a training step (the loss is a leaf: its upstream
    gradient is 1).  Returns (global mean loss, dX_t for THIS rank's label rows, dW for this rank's table rows).
    Collectives: the forward's all-gathers of the label rows and of the per-shard (lse, label logit) pairs, plus ONE
    all-reduce of the partial dX [T_total, De] (every shard contributes P_s W_s); dW needs none -- each rank holds all
    label rows and only its own table rows."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    head_rows = head_rows or _default_head_rows
    train_head = train_head or _default_train_head
    dev = xt.device
    counts_l = _all_gather(torch.tensor([xt.shape[0]], dtype=torch.int64, device=dev), world, group).reshape(-1).tolist()
    T_max, De = max(counts_l), xt.shape[1]
    pad_x = torch.zeros((T_max, De), dtype=torch.float32, device=dev)
    pad_x[: xt.shape[0]] = xt
    pad_y = torch.zeros((T_max,), dtype=torch.int64, device=dev)
    pad_y[: xt.shape[0]] = labels
    all_x, all_y = _all_gather(pad_x, world, group), _all_gather(pad_y, world, group)
    xg = torch.cat([all_x[r, : counts_l[r]] for r in range(world)], dim=0).contiguous()
    yg = torch.cat([all_y[r, : counts_l[r]] for r in range(world)], dim=0).contiguous()
    lo, _ = shard_bounds(V, rank, world)
    part = head_rows(xg, yg, local_table, w_planes, lo, inv_tau)           # [T_total, 2] (lse_s, label logit_s)
    parts = _all_gather(part, world, group)
    lse_global = torch.logsumexp(parts[:, :, 0], dim=0)
    tgt = parts[:, :, 1].sum(dim=0)
    T_total = xg.shape[0]
    loss = ((lse_global - tgt).sum() / max(T_total, 1)).reshape(())
    dx, dW = train_head(xg, yg, local_table, lo, inv_tau, lse_global, 1.0 / max(T_total, 1), head_chunk)
    dx = dx.contiguous()
    dist.all_reduce(dx, group=group)
    start = sum(counts_l[:rank])
    return loss, dx[start:start + counts_l[rank]].contiguous(), dW


# ---------------------------------------------------------------------------------------------------------------
# NVLink peer memory (csrc/t4r_peer.cu): the product path on GPUs.  One process per GPU; every rank maps its peers'
# buffers through CUDA IPC once, after which the kernels read remote rows with plain loads -- no bulk collective.
# ---------------------------------------------------------------------------------------------------------------
_OPENED: dict = {}   # exported-allocation handle (bytes) -> base address of its mapping in this process


def peer_memory_available(t: torch.Tensor, group=None) -> bool:
    """CUDA tensor + NCCL backend + not switched off (``T4R_PEER=0`` keeps the NCCL all-to-all formulation)."""
    import os
    return bool(t.is_cuda and dist.is_initialized() and dist.get_backend(group) == "nccl"
                and os.environ.get("T4R_PEER", "1") != "0")


class PeerView:
    """The same buffer on every rank of ``group`` as addressable from THIS process: ``local`` is this rank's tensor
    (any cudaMalloc'ed torch tensor; it must stay alive and in place), the others are CUDA-IPC mappings of the peers'.
    Construction is a collective (one all-gather of 72 bytes per rank) and synchronises with the host; it happens once
    per buffer, not per step.  ``struct`` is the ``t4r_peer_ptrs`` the kernels take."""

    def __init__(self, local: torch.Tensor, group=None):
        import ctypes as C

        from . import _lib
        if not local.is_cuda or not local.is_contiguous():
            raise _lib.T4RError("PeerView needs a contiguous CUDA tensor")
        lib = _lib.load()
        self.local, self.group = local, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _lib.T4R_MAX_PEERS:
            raise _lib.T4RError(f"peer windows support up to {_lib.T4R_MAX_PEERS} ranks per group")
        hb = _lib.T4R_PEER_HANDLE_BYTES
        handle = (C.c_ubyte * hb)()
        off = C.c_int64(0)
        # every rank runs the SAME two collectives whatever fails locally (a rank that skipped one would desynchronise
        # the communicator): byte hb + 8 of the exchanged record says whether this rank's export worked
        rc = lib.t4r_peer_export(local.data_ptr(), handle, C.byref(off))
        err = "" if rc == 0 else (lib.t4r_last_error() or b"t4r_peer_export failed").decode()
        mine = torch.tensor(list(bytes(handle)) + list(int(off.value if rc == 0 else 0).to_bytes(8, "little")) +
                            [1 if rc == 0 else 0], dtype=torch.uint8, device=local.device)
        allh = _all_gather(mine, self.world, group).cpu()
        self.ptr_of_rank = []
        for r in range(self.world):
            if r == self.rank:
                self.ptr_of_rank.append(local.data_ptr())
                continue
            raw = bytes(allh[r].tolist())
            h, o, ok = raw[:hb], int.from_bytes(raw[hb:hb + 8], "little"), raw[hb + 8]
            if not ok:
                err = err or f"rank {r} could not export its buffer"
                continue
            if err:
                continue
            base = _OPENED.get(h)
            if base is None:   # an allocation can be opened once per process: keep the mapping for later views
                out = C.c_void_p()
                if lib.t4r_peer_open((C.c_ubyte * hb).from_buffer_copy(h), 0, C.byref(out)) != 0:
                    err = (lib.t4r_last_error() or b"t4r_peer_open failed").decode()
                    continue
                base = _OPENED[h] = int(out.value)
            self.ptr_of_rank.append(base + o)
        ok_t = torch.tensor([0 if err else 1], dtype=torch.int32, device=local.device)
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN, group=group)
        if int(ok_t.item()) == 0:
            raise _lib.T4RError("peer memory is not available: " + (err or "a peer rank failed to map the buffers"))
        self.struct = _lib.PeerPtrs()
        self.struct.world, self.struct.rank = self.world, self.rank
        for r, p in enumerate(self.ptr_of_rank):
            self.struct.base[r] = p
        self.local_ptr = local.data_ptr()

    def still_valid(self) -> bool:
        return self.local.data_ptr() == self.local_ptr


def _try_peer_view(local: torch.Tensor, group, what: str):
    """PeerView or None -- the SAME answer on every rank (a failed mapping anywhere turns the feature off everywhere)."""
    import logging

    from . import _lib
    try:
        return PeerView(local, group)   # raises on EVERY rank or on none (its last step is an all-reduce of the outcome)
    except _lib.T4RError as exc:
        logging.getLogger("transformers4rec_b200").warning("%s for %s: using the NCCL all-to-all formulation", exc, what)
        return None


class PeerHead:
    """Per-step windows of the row-sharded full-softmax head over peer memory.

    ``mail_x`` fp32 [cap, De] / ``mail_y`` int64 [cap]: this rank's label rows for the peers to pull;
    ``stats`` fp32 [3, world*cap]: this shard's (lse, label logit, rank count) for ALL label rows, for the peers to
    read.  One step on every rank, in stream order:

        write mail            (gather_rows_split / compact_targets write straight into the window)
        B1: all-gather of the 4-byte counts      -> every peer's mail is complete, counts on the device
        pull                  (peer_pull_rows: exactly counts[r] rows from rank r, planes emitted in the same pass)
        head                  (logits + online LSE over the local V/world rows; statistics into ``stats``)
        B2: 4-byte all-reduce                    -> every peer's statistics are complete
        combine               (peer_combine_lse reads all shards' statistics)

    Write-after-read safety: a rank overwrites its mail for step i+1 only after its own B2(i), which cannot complete
    before every peer has entered B2(i), i.e. after every peer's pull(i); it overwrites its statistics in head(i+1)
    only after its own B1(i+1), which every peer enters after its combine(i).  No host synchronisation anywhere."""

    def __init__(self, cap: int, De: int, device, group=None):
        self.cap, self.De, self.group = int(cap), int(De), group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.cap_g = self.world * self.cap
        self.mail_x = torch.zeros((self.cap, De), dtype=torch.float32, device=device)
        self.mail_y = torch.zeros((self.cap,), dtype=torch.int64, device=device)
        self.stats = torch.zeros((3, self.cap_g), dtype=torch.float32, device=device)
        self.views = None
        caps = _all_gather(torch.tensor([self.cap, De], dtype=torch.int64, device=device), self.world, group).cpu()
        if not bool((caps == caps[0]).all()):
            raise RuntimeError(f"the row-sharded head needs the same batch x length capacity on every rank, got {caps.tolist()}")
        vx = _try_peer_view(self.mail_x, group, "the head's label-row window")
        vy = _try_peer_view(self.mail_y, group, "the head's label window") if vx is not None else None
        vs = _try_peer_view(self.stats, group, "the head's statistics window") if vy is not None else None
        if vs is not None:
            self.views = (vx, vy, vs)
        self.counts = torch.zeros(self.world, dtype=torch.int32, device=device)
        self.token = torch.zeros(1, dtype=torch.int32, device=device)

    @property
    def ok(self) -> bool:
        return self.views is not None


def peer_softmax_ce(ph: PeerHead, count: torch.Tensor, local_table: torch.Tensor, V: int, w_planes=None,
                    inv_tau: float = 1.0, want_rank: bool = False, rank_tgt_fn: Optional[Callable] = None,
                    head_rows: Optional[Callable] = None):
    """Full-softmax CE of the label rows of ALL ranks against the row-sharded output table, over peer memory.
    The caller has written this rank's label rows / labels into ``ph.mail_x`` / ``ph.mail_y`` and holds their count
    on the device (``count`` int32 [1]).  ``rank_tgt_fn(x_all, y_all, t_total)`` (evaluation) returns the labels'
    logits over the WHOLE table.  Returns dict(loss [1], row_loss [cap_g], row_rank or None, t_total, my_start):
    rows of all ranks, rank-major; this rank's rows start at ``my_start`` (device scalars -- nothing syncs)."""
    from . import ops
    vx, vy, vs = ph.views
    rank, world = ph.rank, ph.world
    dist.all_gather_into_tensor(ph.counts, count.reshape(1).to(torch.int32), group=ph.group)          # B1
    pulled = ops.peer_pull_rows(vx, vy, ph.counts, ph.cap, ph.De, want_f32=True)
    xg, yg, t_total = pulled["x"], pulled["labels"], pulled["t_total"]
    lo, _ = shard_bounds(V, rank, world)
    if w_planes is None:
        w_planes = ops.split_planes(local_table)
    rank_tgt = rank_tgt_fn(xg, yg, t_total) if want_rank else None
    if head_rows is not None:
        head_rows(pulled, w_planes, lo, inv_tau, rank_tgt, ph.stats)
    elif isinstance(w_planes, tuple):   # the 2-unit product: (mixed planes, inverse row scales)
        xm, xi = ops.split_planes_mixed(xg, count=t_total)
        ops.head_softmax_ce(xm, xg, yg, w_planes[0], local_table, t_dev=t_total, inv_temperature=inv_tau, v_offset=lo,
                            want_loss=False, want_rank=want_rank, rank_tgt=rank_tgt, nprod=2, xt_inv_scale=xi,
                            w_inv_scale=w_planes[1], out_stats=ph.stats)
    else:
        ops.head_softmax_ce(pulled["planes"], xg, yg, w_planes, local_table, t_dev=t_total, inv_temperature=inv_tau,
                            v_offset=lo, want_loss=False, want_rank=want_rank, rank_tgt=rank_tgt, out_stats=ph.stats)
    dist.all_reduce(ph.token, group=ph.group)                                                           # B2
    row_loss, loss, row_rank = ops.peer_combine_lse(vs, ph.cap_g, t_total, with_rank=want_rank)
    return {"loss": loss, "row_loss": row_loss, "row_rank": row_rank, "t_total": t_total, "my_start": pulled["my_start"]}


class ShardedEmbedding(torch.nn.Module):
    """Rows [lo, hi) of an item table of ``num_embeddings`` rows, block-partitioned over the process
    group: the drop-in for the item feature's ``nn.Embedding`` (and, under weight tying, for the output
    layer of the head) in BASELINE configs 4-5.  ``weight`` is the LOCAL shard."""

    def __init__(self, num_embeddings: int, embedding_dim: int, padding_idx: int = 0, group=None,
                 initializer: Optional[Callable] = None, device=None):
        super().__init__()
        if not dist.is_initialized():
            raise RuntimeError("ShardedEmbedding needs an initialised torch.distributed process group")
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.num_embeddings, self.embedding_dim, self.padding_idx = int(num_embeddings), int(embedding_dim), padding_idx
        self.lo, self.hi = shard_bounds(self.num_embeddings, self.rank, self.world)
        self.weight = torch.nn.Parameter(torch.empty((self.hi - self.lo, embedding_dim), device=device))
        (initializer or (lambda w: torch.nn.init.normal_(w, mean=0.0, std=0.05)))(self.weight)
        # local compute steps of the exchange; None = the CUDA kernels (the gloo tests inject stand-ins)
        self.gather_rows: Optional[Callable] = None
        self.place: Optional[Callable] = None
        # peer-memory view of all shards (set up lazily at the first lookup on GPUs; None = NCCL all-to-all)
        self._peer: Optional[PeerView] = None
        self._peer_tried = False
        self._peer_version = -1

    @classmethod
    def from_full(cls, full_weight: torch.Tensor, group=None, padding_idx: int = 0) -> "ShardedEmbedding":
        """Take this rank's block of a replicated table (tests / converting a trained model)."""
        m = cls(full_weight.shape[0], full_weight.shape[1], padding_idx, group, initializer=lambda w: None,
                device=full_weight.device)
        with torch.no_grad():
            m.weight.copy_(full_weight[m.lo:m.hi])
        return m

    @property
    def rows_per_shard(self) -> int:
        return (self.num_embeddings + self.world - 1) // self.world

    def peer_view(self) -> Optional[PeerView]:
        """The shards of all ranks as peer memory, or None when that is not available (CPU / gloo tests, T4R_PEER=0,
        a failed mapping).  The first call on a CUDA weight is a collective."""
        w = self.weight
        if self._peer is not None and w.data_ptr() != self._peer.local_ptr:
            self._peer, self._peer_tried = None, False          # the parameter was re-allocated (.to(), load): map again
        if self._peer is None and not self._peer_tried:
            self._peer_tried = True
            if peer_memory_available(w, self.group) and self.embedding_dim % 4 == 0 and self.embedding_dim <= 1024 \
                    and self.gather_rows is None and self.place is None:
                self._peer = _try_peer_view(w.detach(), self.group, "the row-sharded item table")
                self._peer_version = w._version
        return self._peer

    def lookup(self, ids: torch.Tensor, ragged: bool = False, plan_out: Optional[list] = None,
               count: Optional[torch.Tensor] = None):
        """-> (rows fp32 [ids.numel(), dim], split planes) for arbitrary global ids.

        On GPUs the rows are read straight from their owners' shards over NVLink peer memory (``t4r_peer_gather_rows``:
        no collective, no host synchronisation; ``count`` = optional device-side number of valid ids, the rest give
        zero rows).  Otherwise -- and whenever the caller needs the routing plan for the backward (``plan_out``) --
        one all-gather of the ids + one all-to-all of the rows.  ``ragged``: the ranks pass different numbers of ids
        (label rows); for the all-to-all they are padded to the longest with the padding id."""
        flat = ids.reshape(-1)
        n = flat.numel()
        peer = self.peer_view() if plan_out is None else None
        if peer is not None:
            from . import ops
            if self.weight._version != self._peer_version:
                # the shards moved (an optimizer step): every rank must be done writing before anyone reads remotely
                dist.barrier(group=self.group)
                self._peer_version = self.weight._version
            return ops.peer_gather_rows(peer, self.num_embeddings, self.rows_per_shard, self.embedding_dim, flat, count,
                                        pad_id=self.padding_idx if self.padding_idx is not None else -1)
        if count is not None:
            raise NotImplementedError("a device-side id count needs the peer-memory lookup")
        if ragged:
            counts = _all_gather(torch.tensor([n], dtype=torch.int64, device=flat.device), self.world, self.group)
            n_max = int(counts.max())
            if n_max > n:
                flat = torch.cat([flat, flat.new_full((n_max - n,), self.padding_idx)])
        rows, planes = sharded_embedding_lookup(self.weight.detach(), flat, self.num_embeddings, self.group,
                                                gather_rows=self.gather_rows, place=self.place, plan_out=plan_out)
        if ragged and rows.shape[0] != n:
            rows, planes = rows[:n], (planes[:, :n] if planes is not None else None)
        return rows, planes

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        rows, _ = self.lookup(ids)
        return rows.view(*ids.shape, self.embedding_dim)

    def extra_repr(self) -> str:
        return f"{self.num_embeddings}, {self.embedding_dim}, rows [{self.lo}, {self.hi}) of rank {self.rank}/{self.world}"
