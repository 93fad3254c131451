"""TEST DOUBLES for ``transformers4rec_b200.ops`` (plain torch, CPU).

The product has no CPU path: every op in ``ops.py`` hands device pointers to ``libt4r_b200.so``.  To exercise the
HOST logic of the package (module wiring, state side channels, lazy outputs, masking modes, the sharded
choreography through the model, bench helpers) in the ``-m "not gpu"`` suite, the tests -- and only the tests --
swap those ops for the stand-ins below with ``install(monkeypatch)``.  Each stand-in mirrors the signature, the output
shapes / dtypes and the padding conventions of its kernel; the arithmetic is the oracle's.  The 2-unit product
(``nprod=2``) is emulated bit-for-bit from the packed operands (tests/_mixed_ref.py), so the control flow that
selects it sees what the tensor core would produce.  Nothing here is imported by the package.
"""
import torch
import torch.nn.functional as F

import _mixed_ref as R
import t4r_oracle as O
from transformers4rec_b200 import _lib
from transformers4rec_b200.ops import dropout as _REAL_DROPOUT   # bound before install() swaps ops.dropout


def _r64(k):
    return (k + 63) // 64 * 64


def make_planes(x, Kp=None):
    rows, K = x.shape
    Kp = Kp or _r64(K)
    xp = torch.zeros((rows, Kp), dtype=torch.float32)
    xp[:, :K] = x
    hi = xp.to(torch.bfloat16)
    lo = (xp - hi.float()).to(torch.bfloat16)
    return torch.stack([hi, lo])


def from_planes(p, K=None):
    x = p[0].float() + p[1].float()
    return x if K is None else x[:, :K]


def _count(t_dev, cap):
    return cap if t_dev is None else min(int(t_dev.reshape(-1)[0]), cap)


# ---------------------------------------------------------------------------------------------------- packing
def split_planes(x, row_code=None, mask_vec=None, want_f32=False, out=None):
    v = x.float().clone()
    if row_code is not None:
        rc = row_code.reshape(-1)
        v[rc == 1] = mask_vec.float()
        v[rc == 2] = 0.0
    planes = make_planes(v)
    if out is not None:
        out.copy_(planes)
        planes = out
    return (planes, v) if want_f32 else planes


def split_planes_mixed(x, count=None):
    from transformers4rec_b200 import ops
    return ops.split_planes_mixed_host(x)  # the kernel's own packing code, compiled for the host


def gather_rows_split(x2d, idx, count, cap, want_f32=True):
    n = _count(count, cap) if idx.dtype == torch.int32 else cap
    K = x2d.shape[1]
    of = torch.zeros((cap, K), dtype=torch.float32)
    of[:n] = x2d.float()[idx[:n].long()]
    return make_planes(of), (of if want_f32 else None)


def compact_targets(masked_targets, padding_idx=0):
    mt = masked_targets.long().reshape(-1)
    nz = (mt != padding_idx).nonzero().squeeze(1)
    rows = torch.zeros(mt.numel(), dtype=torch.int32)
    labels = torch.zeros(mt.numel(), dtype=torch.int64)
    rows[: nz.numel()] = nz.int()
    labels[: nz.numel()] = mt[nz]
    return rows, labels, torch.tensor([nz.numel()], dtype=torch.int32)


# ---------------------------------------------------------------------------------------------------- K1 / K3
def embed_concat(cats, conts, M, C_width, want_f32, want_planes):
    out = torch.zeros((M, C_width), dtype=torch.float32)
    err = torch.zeros(1, dtype=torch.int32)
    for table, ids, col in cats:
        ids = ids.reshape(-1).long()
        bad = (ids < 0) | (ids >= table.shape[0])
        if bad.any():
            err[0] = 1
        out[:, col:col + table.shape[1]] = table.detach().float()[torch.where(bad, torch.zeros_like(ids), ids)]
    for vals, col in conts:
        out[:, col] = vals.reshape(-1).float()
    return (out if want_f32 else None), (make_planes(out) if want_planes else None), err


def input_block(feats, M, L, C_width, agg=_lib.AGG_CONCAT, item_feature=-1, ln_eps=1e-5, want_f32=True, want_planes=False):
    err = torch.zeros(1, dtype=torch.int32)
    vals = []
    for f in feats:
        kind, dim, x = int(f["kind"]), int(f["dim"]), f["input"]
        per_session = bool(f.get("per_session"))

        def rows(t):  # [B(, ...)] per-session values are repeated for every position of the sequence
            return t.repeat_interleave(L, dim=0) if per_session else t
        if kind == _lib.FEAT_CAT:
            ids = rows(x.reshape(-1).long())
            table = f["table"].detach().float()
            bad = (ids < 0) | (ids >= table.shape[0])
            if bad.any():
                err[0] = 1
            v = table[torch.where(bad, torch.zeros_like(ids), ids)]
        elif kind == _lib.FEAT_CONT:
            v = rows(x.reshape(-1).float()).unsqueeze(-1)
        elif kind == _lib.FEAT_SOFT:
            v = O.soft_embedding(rows(x.reshape(-1).float()), f["soft_w"].detach().reshape(-1, 1).float(),
                                 f["soft_b"].detach().reshape(-1).float(), f["table"].detach().float())
        else:
            v = rows(x.reshape(-1, dim).float())
        assert v.shape == (M, dim), (v.shape, M, dim)
        if f.get("ln") is not None:
            v = F.layer_norm(v, (dim,), f["ln"][0].detach().float(), f["ln"][1].detach().float(), ln_eps)
        vals.append((int(f.get("col", 0)), dim, v))
    if agg == _lib.AGG_CONCAT:
        out = torch.zeros((M, C_width), dtype=torch.float32)
        for col, dim, v in vals:
            out[:, col:col + dim] = v
    elif agg == _lib.AGG_SUM:
        out = torch.stack([v for _, _, v in vals]).sum(0)
    else:
        others = torch.stack([v for i, (_, _, v) in enumerate(vals) if i != item_feature]).sum(0)
        out = vals[item_feature][2] * others
    return (out if want_f32 else None), (make_planes(out) if want_planes else None), err


def swap_noise(values, keep_mask, u, perm, replacement_prob):
    return O.stochastic_swap_noise(values, keep_mask, u, perm.long(), replacement_prob)


def mask_mlm(item_ids, mode, padding_idx=0, mlm_probability=0.15, u=None):
    ids = item_ids.long()
    B, L = ids.shape
    if mode == _lib.MLM_TRAIN:
        if u is None:
            u = torch.rand((B, L + 2))
        m, l = O.mlm_compute_masked_targets(ids, True, False, mlm_probability=mlm_probability, padding_idx=padding_idx,
                                            u_bern=u[:, :L], u_force=u[:, L], u_unmask=u[:, L + 1])
    elif mode == _lib.MLM_INFERENCE:
        m, l = O.mlm_compute_masked_targets(ids, False, False, padding_idx=padding_idx)
    else:
        m, l = O.mlm_compute_masked_targets(ids, False, True, padding_idx=padding_idx,
                                            eval_on_last_item_seq_only=(mode == _lib.MLM_EVAL_LAST))
    return m, l, m.to(torch.uint8)


def mask_plm(item_ids, mode, padding_idx=0, max_span_length=5, plm_probability=1 / 6, draws=None):
    ids = item_ids.long()
    B, L = ids.shape
    if mode == _lib.PLM_TRAIN and draws is None:
        draws = {"u_span": torch.rand((B, L)), "u_start": torch.rand((B, L)), "u_force": torch.rand((B,)),
                 "u_unmask": torch.rand((B,)), "perm": torch.argsort(torch.rand((B, L)), dim=1)}
    m, l, _, pm, _ = O.plm_compute_masked_targets(ids, mode == _lib.PLM_TRAIN, padding_idx=padding_idx,
                                                  eval_on_last_item_seq_only=(mode != _lib.PLM_EVAL_ALL),
                                                  plm_probability=plm_probability, max_span_length=max_span_length,
                                                  draws=draws)
    return m, l, pm.to(torch.uint8)


def mask_clm(item_ids, mode, padding_idx=0):
    ids = item_ids.long()
    training, testing = (mode == _lib.CLM_ALL), (mode == _lib.CLM_LAST)
    m, l = O.clm_compute_masked_targets(ids, training, testing, padding_idx=padding_idx) if mode != _lib.CLM_INFERENCE \
        else O.clm_compute_masked_targets(ids, False, False, padding_idx=padding_idx)
    probe = O.clm_apply_mask_to_inputs(torch.ones(ids.shape + (1,)), m, torch.tensor([2.0]), training=training,
                                       testing=testing)[..., 0]
    code = torch.zeros(ids.shape, dtype=torch.uint8)
    code[probe == 2.0] = 1
    code[probe == 0.0] = 2
    return m, l, code


# ---------------------------------------------------------------------------------------------------- K2
def linear(x_planes, w_planes, K, *, bias=None, act=_lib.ACT_NONE, row_code=None, mask_vec=None, residual=None, ln=None,
           ln_eps=0.0, want_f32=True, want_planes=True, want_pre_ln=False, m_dev=None, nprod=3):
    x, w = from_planes(x_planes, K), from_planes(w_planes, K)
    M = x.shape[0]
    n = _count(m_dev, M)
    y = x @ w.t()
    if bias is not None:
        y = y + bias.detach().float()
    if act == _lib.ACT_RELU:
        y = torch.relu(y)
    elif act == _lib.ACT_GELU:
        y = F.gelu(y)
    if row_code is not None:
        rc = row_code.reshape(-1)
        y[rc == 1] = mask_vec.detach().float()
        y[rc == 2] = 0.0
    if residual is not None:
        y = y + residual.float()
    pre = y.clone()
    if ln is not None:
        y = F.layer_norm(y, (y.shape[1],), ln[0].detach().float(), ln[1].detach().float(), ln_eps)
    y[n:] = 0.0
    return (y if want_f32 else None), (make_planes(y) if want_planes else None), (pre if want_pre_ln else None)


# ---------------------------------------------------------------------------------------------------- head
def _logits(xt_planes, w_planes, De, nprod, xt_inv, w_inv):
    if nprod == 2:
        ha, h8a, l8a = R.unpack_planes(xt_planes)
        hb, h8b, l8b = R.unpack_planes(w_planes)
        pa = {"h16": ha, "hi8": h8a, "lo8": l8a, "inv_scale": xt_inv}
        pb = {"h16": hb, "hi8": h8b, "lo8": l8b, "inv_scale": w_inv}
        return R.product(pa, pb).float()
    return from_planes(xt_planes, De) @ from_planes(w_planes, De).t()


def label_logit(xt_f32, w_f32, labels, *, t_dev=None, class_bias=None, inv_temperature=1.0, v_offset=0):
    T_cap = xt_f32.shape[0]
    n = _count(t_dev, T_cap)
    loc = labels.long() - v_offset
    mine = (loc >= 0) & (loc < w_f32.shape[0])
    rows = w_f32.detach().float()[loc.clamp(0, w_f32.shape[0] - 1)]
    val = (xt_f32.float() * rows).sum(1)
    if class_bias is not None:
        val = val + class_bias[loc.clamp(0, w_f32.shape[0] - 1)]
    out = torch.where(mine, val * inv_temperature, torch.zeros_like(val))
    out[n:] = 0.0
    return out


def head_softmax_ce(xt_planes, xt_f32, labels, w_planes, w_f32, *, t_dev=None, inv_temperature=1.0, col_bias=None,
                    col_ids=None, hit_value=0.0, pos_logit=None, v_offset=0, want_rank=False, want_loss=True, nprod=3,
                    events=None, label_smoothing=0.0, rank_tgt=None, xt_inv_scale=None, w_inv_scale=None, out_stats=None,
                    col_ids_sorted_unique=False):
    if col_ids is not None and col_ids_sorted_unique:
        assert bool((col_ids[1:] > col_ids[:-1]).all()), "col_ids_sorted_unique promised, but the ids are not ascending"
    if nprod == 2 and (xt_inv_scale is None or w_inv_scale is None):
        raise _lib.T4RError("head_softmax_ce: nprod=2 needs the mixed planes' inverse row scales")
    T_cap, V = xt_planes.shape[1], w_planes.shape[1]
    De = xt_f32.shape[1] if xt_f32 is not None else w_f32.shape[1]
    n = _count(t_dev, T_cap)
    z = _logits(xt_planes, w_planes, De, nprod, xt_inv_scale, w_inv_scale)
    if col_bias is not None:
        z = z + col_bias.unsqueeze(0)
    y = labels.long() if labels is not None else None
    if col_ids is not None:
        z = torch.where(col_ids.unsqueeze(0) == y.unsqueeze(1), torch.full_like(z, hit_value), z)
    z = z * inv_temperature
    row_tgt = torch.zeros(T_cap)
    if pos_logit is not None:  # sampled softmax: class 0 = the positive
        full = torch.cat([pos_logit.reshape(-1, 1).float(), z], dim=1)
        row_lse = torch.logsumexp(full, dim=1)
        row_tgt = pos_logit.float().clone()
    else:
        row_lse = torch.logsumexp(z, dim=1)
        row_tgt = label_logit(xt_f32, w_f32, y, inv_temperature=inv_temperature, v_offset=v_offset)
    row_loss = row_lse - row_tgt
    if label_smoothing:
        row_loss = row_lse - (1 - label_smoothing) * row_tgt - (label_smoothing / V) * z.sum(1)
    row_rank = None
    if want_rank:
        tgt = (rank_tgt if rank_tgt is not None else row_tgt).unsqueeze(1)
        col = torch.arange(V).unsqueeze(0) + v_offset
        above = ((z > tgt) | ((z == tgt) & (col < y.unsqueeze(1)))) & (col != y.unsqueeze(1))
        row_rank = above.sum(1).to(torch.int32)
        row_rank[n:] = 0
    for t in (row_lse, row_tgt, row_loss):
        t[n:] = 0.0
    loss = (row_loss[:n].sum() / max(n, 1)).reshape(1) if want_loss else None
    return {"row_lse": row_lse, "row_tgt": row_tgt, "row_loss": row_loss, "loss": loss, "row_rank": row_rank}


def head_logits(xt_planes, w_planes, De, *, t_dev=None, inv_temperature=1.0, nprod=3):
    if nprod == 2:
        raise _lib.T4RError("head_logits: nprod = 2 operands go through t4r_head_logits_mixed")
    out = _logits(xt_planes, w_planes, De, nprod, None, None) * inv_temperature
    out[_count(t_dev, out.shape[0]):] = 0.0
    return out


def head_logits_mixed(xt_planes, xt_inv_scale, w_planes, w_inv_scale, De, *, t_dev=None, inv_temperature=1.0):
    out = _logits(xt_planes, w_planes, De, 2, xt_inv_scale, w_inv_scale) * inv_temperature
    out[_count(t_dev, out.shape[0]):] = 0.0
    return out


def topk(logits, k):
    order = torch.sort(-logits.float(), dim=1, stable=True).indices[:, :k]  # value desc, index asc: t4r_topk's order
    return logits.float().gather(1, order), order


def recall_from_ranks(row_rank, ks, t_dev=None):
    n = _count(t_dev, row_rank.numel())
    r = row_rank[:n]
    return torch.tensor([float((r < k).float().mean()) if n else 0.0 for k in ks])


def metrics_from_ranks(row_rank, ks, kind, t_dev=None):
    n = _count(t_dev, row_rank.numel())
    r = row_rank[:n].float()
    outs = []
    for k in ks:
        hit = (r < k).float()
        if kind == _lib.METRIC_RECALL:
            v = hit
        elif kind == _lib.METRIC_PRECISION:
            v = hit / k
        elif kind == _lib.METRIC_RR:
            v = hit / (r + 1)
        else:
            v = hit / torch.log2(r + 2)
        outs.append(v.mean() if n else torch.tensor(0.0))
    return torch.stack(outs)


def combine_shard_lse(parts, t_dev=None):
    n = _count(t_dev, parts.shape[1])
    row = torch.logsumexp(parts[:, :, 0], dim=0) - parts[:, :, 1].sum(0)
    row[n:] = 0.0
    return row, (row[:n].sum() / max(n, 1)).reshape(1)


def _pad(values, offsets, rows, in_len, pad_len):
    """padding._pad: ragged (values, offsets) or dense [rows, in_len] -> [rows, pad_len], zeros on the right."""
    if values.dtype in (torch.int32, torch.int16, torch.uint8, torch.bool):
        values = values.long()
    elif values.dtype not in (torch.int64, torch.float32):
        values = values.float()
    if offsets is not None:
        return O.pad_ragged(values, offsets.long(), pad_len)
    out = torch.zeros((rows, pad_len), dtype=values.dtype)
    n = min(in_len, pad_len)
    out[:, :n] = values.reshape(rows, in_len)[:, :n]
    return out


# ---------------------------------------------------------------------------------------------------- training (N3)
def transpose(x):
    return x.t().contiguous()


def col_sum(x):
    return x.sum(0)


def rel_pos_table(L, d, device=None):
    return O.xlnet_relative_positions(L, d)


def rel_pos_proj(wr_list, L, d):
    pos = O.xlnet_relative_positions(L, d)
    return torch.stack([pos @ w for w in wr_list])


def _xl_scores(qkv, R, rw, rr, B, L, H):
    d = qkv.shape[1] // 3
    dh = d // H
    q, k, v = (t.reshape(B, L, H, dh) for t in qkv.split(d, dim=1))
    r = R.reshape(2 * L, H, dh)
    ac = torch.einsum("bihd,bjhd->bhij", q + rw.view(H, dh), k)
    bd_full = torch.einsum("bihd,mhd->bhim", q + rr.view(H, dh), r)
    idx = (torch.arange(L).view(1, L) + L - torch.arange(L).view(L, 1))
    bd = torch.gather(bd_full, 3, idx.view(1, 1, L, L).expand(B, H, L, L))
    return (ac + bd) / dh ** 0.5, v


def dropout(x, p, seed, site):
    """The mask IS the kernel's definition (Philox keyed on (seed, site, flat index)): its host twin, on any shape."""
    if p <= 0.0:
        return x
    mask = _REAL_DROPOUT(torch.ones(x.numel()), p, seed, site, _on_host=True).reshape(x.shape)
    return x * mask


def _prob_mask(shape, drop):
    return None if drop is None or drop[0] <= 0.0 else dropout(torch.ones(shape), *drop)


def xlnet_attn_fwd(qkv, R, rw, rr, B, L, H, drop=None):
    s, v = _xl_scores(qkv, R, rw, rr, B, L, H)
    p = torch.softmax(s, -1)
    mk = _prob_mask((B, H, L, L), drop)
    return torch.einsum("bhij,bjhd->bihd", p if mk is None else p * mk, v).reshape(B * L, -1)


def attn_drop_fwd(qkv, R, rw, rr, B, L, H, drop, plm_mask=None):
    if R is None:
        return causal_attn_fwd(qkv, B, L, H, drop=drop)
    if plm_mask is not None:
        return xlnet_attn_plm_fwd(qkv, R, rw, rr, B, L, H, plm_mask, drop=drop)
    return xlnet_attn_fwd(qkv, R, rw, rr, B, L, H, drop=drop)


def xlnet_attn_bwd(qkv, R, rw, rr, dout, B, L, H, plm_mask=None, drop=None):
    with torch.enable_grad():
        a, b, c, e = (t.detach().clone().requires_grad_(True) for t in (qkv, R, rw, rr))
        out = (xlnet_attn_fwd(a, b, c, e, B, L, H, drop=drop) if plm_mask is None else
               xlnet_attn_plm_fwd(a, b, c, e, B, L, H, plm_mask, drop=drop))
        out.backward(dout)
    return a.grad, b.grad, c.grad, e.grad


def xlnet_attn_plm_fwd(qkv, R, rw, rr, B, L, H, plm_mask, drop=None):
    """HF:xlnet two-stream rel_attn_core with target_mapping = identity: rows [0, M) = content stream h, [M, 2M) = query
    stream g; K / V from h; score - 1e30 * mask (h: diagonal exempt)."""
    M = B * L
    d = qkv.shape[1] // 3
    dh = d // H
    kv = qkv[:M]
    outs = []
    mk_all = _prob_mask((2, B, H, L, L), drop)
    for st in range(2):
        q_rows = qkv[st * M:(st + 1) * M, :d]
        fake = torch.cat([q_rows, kv[:, d:]], dim=1)
        s, v = _xl_scores(fake, R, rw, rr, B, L, H)
        m = plm_mask.float().view(B, 1, L, L)
        if st == 0:
            m = m * (1.0 - torch.eye(L)).view(1, 1, L, L)
        s = s - 1e30 * m
        pr = torch.softmax(s, -1)
        if mk_all is not None:
            pr = pr * mk_all[st]
        outs.append(torch.einsum("bhij,bjhd->bihd", pr, v).reshape(M, d))
    return torch.cat(outs, dim=0)


def causal_attn_fwd(qkv, B, L, H, drop=None):
    d = qkv.shape[1] // 3
    dh = d // H
    q, k, v = (t.reshape(B, L, H, dh).transpose(1, 2) for t in qkv.split(d, dim=1))
    s = (q @ k.transpose(-1, -2)) / dh ** 0.5
    s = s.masked_fill(~torch.tril(torch.ones(L, L, dtype=torch.bool)), float("-inf"))
    pr = torch.softmax(s, -1)
    mk = _prob_mask((B, H, L, L), drop)
    return ((pr if mk is None else pr * mk) @ v).transpose(1, 2).reshape(B * L, d)


def causal_attn_bwd(qkv, dout, B, L, H, drop=None):
    with torch.enable_grad():
        a = qkv.detach().clone().requires_grad_(True)
        causal_attn_fwd(a, B, L, H, drop=drop).backward(dout)
    return a.grad


def layer_norm_fwd(x, gamma, beta, eps):
    return F.layer_norm(x, (x.shape[1],), gamma.float(), beta.float(), eps)


def layer_norm_bwd(x_pre, gamma, eps, dy, add=None):
    with torch.enable_grad():
        a = x_pre.detach().clone().requires_grad_(True)
        g = gamma.detach().clone().float().requires_grad_(True)
        b = torch.zeros_like(g, requires_grad=True)
        F.layer_norm(a, (a.shape[1],), g, b, eps).backward(dy)
    return (a.grad if add is None else a.grad + add), g.grad, b.grad


def act_fwd(kind, x):
    return F.gelu(x) if kind == _lib.ACT_GELU else (torch.relu(x) if kind == _lib.ACT_RELU else x)


def act_bwd(kind, pre, dy):
    with torch.enable_grad():
        a = pre.detach().clone().requires_grad_(True)
        act_fwd(kind, a).backward(dy)
    return a.grad


def add_positions(x, wpe, B, L):
    return (x.reshape(B, L, -1) + wpe[:L].float().unsqueeze(0)).reshape(B * L, -1)


def sum_over_sessions(x, B, L):
    return x.reshape(B, L, -1).sum(0)


def apply_row_codes(y, code, mask_vec):
    out = y.clone()
    out[code == 1] = mask_vec
    out[code == 2] = 0.0
    return out


def row_codes_bwd(dx, code):
    dmask = dx[code == 1].sum(0)
    dy = dx.clone()
    dy[code != 0] = 0.0
    return dmask, dy


def gather_rows(x, idx):
    return x[idx.long()].contiguous()


def scatter_rows(src, idx, n_rows):
    out = torch.zeros((n_rows, src.shape[1]), dtype=src.dtype)
    out[idx.long()] = src
    return out


def softmax_ce_bwd(z, row_lse, labels, v0, scale, label_smoothing=0.0, V_total=None):
    V_total = V_total if V_total is not None else z.shape[1]
    P = (torch.exp(z - row_lse[: z.shape[0]].unsqueeze(1)) - label_smoothing / V_total) * scale
    loc = labels.long() - v0
    mine = (loc >= 0) & (loc < z.shape[1])
    rows = torch.arange(z.shape[0])[mine]
    P[rows, loc[mine]] -= (1.0 - label_smoothing) * scale
    return P


def sampled_ce_bwd(z, row_lse, labels, col_bias, col_ids, inv_tau, scale):
    P = torch.exp(z + col_bias.unsqueeze(0) * inv_tau - row_lse[: z.shape[0]].unsqueeze(1)) * scale
    P[labels.long().unsqueeze(1) == col_ids.long().unsqueeze(0)] = 0.0
    return P


def soft_emb_fwd(x, w, b, table):
    p = torch.softmax(x.reshape(-1, 1).float() * w.reshape(1, -1).float() + b.reshape(1, -1).float(), dim=-1)
    return p @ table.float(), p


def soft_emb_bwd(x, table, p, dout):
    dp = dout @ table.float().t()
    dl = p * (dp - (p * dp).sum(-1, keepdim=True))
    return dl, dl * x.reshape(-1, 1)


def ew_add(a, b):
    return a + b


def ew_mul(a, b):
    return a * b


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step):
    p.mul_(1.0 - lr * weight_decay)
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    bc1, bc2 = 1.0 - beta1 ** step, 1.0 - beta2 ** step
    p.addcdiv_(m, v.sqrt() / bc2 ** 0.5 + eps, value=-lr / bc1)
    return p


def index_add_rows(dst, idx, src, col, width, skip_index=None):
    idx = idx.long()
    keep = torch.ones_like(idx, dtype=torch.bool) if skip_index is None else idx != skip_index
    dst.index_add_(0, idx[keep], src[keep][:, col:col + width].to(dst.dtype))
    return dst


TRAIN_OPS = dict(transpose=transpose, col_sum=col_sum, rel_pos_table=rel_pos_table, rel_pos_proj=rel_pos_proj,
                 xlnet_attn_fwd=xlnet_attn_fwd, xlnet_attn_plm_fwd=xlnet_attn_plm_fwd, xlnet_attn_bwd=xlnet_attn_bwd,
                 causal_attn_fwd=causal_attn_fwd, dropout=dropout, attn_drop_fwd=attn_drop_fwd,
                 causal_attn_bwd=causal_attn_bwd, layer_norm_fwd=layer_norm_fwd, layer_norm_bwd=layer_norm_bwd,
                 act_fwd=act_fwd, act_bwd=act_bwd, add_positions=add_positions, sum_over_sessions=sum_over_sessions,
                 apply_row_codes=apply_row_codes, row_codes_bwd=row_codes_bwd, gather_rows=gather_rows,
                 scatter_rows=scatter_rows, softmax_ce_bwd=softmax_ce_bwd, sampled_ce_bwd=sampled_ce_bwd, adamw_step=adamw_step,
                 soft_emb_fwd=soft_emb_fwd, soft_emb_bwd=soft_emb_bwd, ew_add=ew_add, ew_mul=ew_mul,
                 index_add_rows=index_add_rows)


# ---------------------------------------------------------------------------------------------------- encoders
def _xlnet_forward(self, inputs_embeds, perm_mask=None, target_mapping=None, **kwargs):
    hf = O.build_hf_xlnet(self.config.d_model, self.config.n_head, self.config.n_layer).eval()
    missing, _ = hf.load_state_dict(self.state_dict(), strict=False)
    assert not [m for m in missing if "word_embedding" not in m], missing
    with torch.no_grad():
        if perm_mask is not None:   # permutation language modeling: HF's two-stream forward, output[0] = g
            L = inputs_embeds.shape[1]
            tm = torch.eye(L).expand(inputs_embeds.shape[0], L, L)
            return (O.hf_encoder_forward_plm(hf, inputs_embeds.float(), perm_mask, tm),)
        return (O.hf_encoder_forward(hf, inputs_embeds.float()),)


def _gpt2_forward(self, inputs_embeds, **kwargs):
    B, L, d = inputs_embeds.shape
    if L > self.config.n_positions:
        raise ValueError(f"sequence length {L} exceeds n_positions {self.config.n_positions}")
    hf = O.build_hf_gpt2(self.config.n_embd, self.config.n_head, self.config.n_layer, self.config.n_positions).eval()
    missing, _ = hf.load_state_dict(self.state_dict(), strict=False)
    assert not [m for m in missing if "wte" not in m], missing
    with torch.no_grad():
        return (O.hf_encoder_forward(hf, inputs_embeds.float()),)


OPS = dict(mask_plm=mask_plm, input_block=input_block, swap_noise=swap_noise, split_planes=split_planes, split_planes_mixed=split_planes_mixed, gather_rows_split=gather_rows_split,
           compact_targets=compact_targets, embed_concat=embed_concat, mask_mlm=mask_mlm, mask_clm=mask_clm, linear=linear,
           label_logit=label_logit, head_softmax_ce=head_softmax_ce, head_logits=head_logits,
           head_logits_mixed=head_logits_mixed, topk=topk, recall_from_ranks=recall_from_ranks,
           metrics_from_ranks=metrics_from_ranks, combine_shard_lse=combine_shard_lse)


def install(monkeypatch):
    """Swap the kernels for the stand-ins for the duration of one test."""
    from transformers4rec_b200 import block, ops
    for name, fn in list(OPS.items()) + list(TRAIN_OPS.items()):
        monkeypatch.setattr(ops, name, fn, raising=False)
    monkeypatch.setattr(block.XLNetEncoder, "forward", _xlnet_forward)
    monkeypatch.setattr(block.GPT2Encoder, "forward", _gpt2_forward)
    from transformers4rec_b200 import padding
    monkeypatch.setattr(padding, "_pad", _pad)


def install_plain():
    """Same without pytest (spawned gloo workers): returns nothing, patches for the life of the process."""
    from transformers4rec_b200 import block, ops
    for name, fn in list(OPS.items()) + list(TRAIN_OPS.items()):
        setattr(ops, name, fn)
    block.XLNetEncoder.forward = _xlnet_forward
    block.GPT2Encoder.forward = _gpt2_forward
