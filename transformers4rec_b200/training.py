"""N3: backward of the fused path, so that a training step can run on it.

``FusedTrainingStep(model)`` runs forward + backward of

    TabularSequenceFeatures (concat + Linear/ReLU projection + MLM / CLM masking)
      -> XLNet / GPT-2 encoder -> NextItemPredictionTask (tied weights, full softmax, optional task_block)

with the t4r kernels and leaves the gradients in ``param.grad`` (accumulating, like autograd), so any
``torch.optim`` optimizer -- plumbing, not the product -- can step.  ``training_loss(model, batch)`` wraps it
in one ``torch.autograd.Function`` whose ``backward`` hands those gradients to autograd: ``loss.backward()``
(HF Trainer's ``training_step``, trainer.py:315-338 in the reference) then works unchanged.

How it is built: the training forward is the same math as the inference kernels' but written as a sequence of
primitive ops that keep what the backward needs (layer inputs, q|k|v, attention output, pre-LayerNorm sums,
pre-GELU activations); every matmul of the backward is the existing tcgen05 split-bf16 GEMM on transposed
operands (``dX = dY W``, ``dW = dY^T X``), the rest are small element / row kernels (t4r_train.cu): transpose,
GELU / ReLU derivative, LayerNorm backward, column sums, attention backward (XLNet relative and GPT-2 causal),
softmax-cross-entropy backward, row scatter-add.  The head never materialises [T, V]: the table is walked in
column chunks, logits are recomputed per chunk from the saved log-sum-exp, turned into ``(softmax - onehot) / T``
in place and consumed by two GEMMs (``dX_t += P W_c``, ``dW_c = P^T X_t``).

Status: written without access to a GPU.  The COMPOSITION (every formula, every transposition, the bookkeeping of
masks / codes / tied weights) is verified on the CPU against torch autograd of the oracle graph with the kernels
replaced by test doubles (tests/test_host_training_cpu.py); each new kernel is a one-thread-per-row/element call of
a ``__host__ __device__`` function whose host twin is checked against torch on the CPU.  Not yet run on hardware;
nothing in the inference path uses this module.  Covered: full softmax (replicated or row-sharded table) and sampled
softmax (replicated table), label smoothing (replicated full softmax), MLM / CLM / PLM masking, the widened input block
(per-feature LayerNorm, soft embeddings, continuous projection, element-wise aggregations; replicated tables),
StochasticSwapNoise as the input block's pre-transform.  Not covered: dropout.
"""
from __future__ import annotations

import math
import logging
from typing import Dict, List, Optional

import torch

from . import _lib, ops
from .block import GPT2Encoder, XLNetEncoder
from .masking import CausalLanguageModeling, MaskedLanguageModeling, PermutationLanguageModeling

LOG = logging.getLogger("transformers4rec_b200")


# --------------------------------------------------------------------------------------------------------------
# matmul helpers on top of the tcgen05 GEMM:  C = A B^T with A [M, K], B [N, K] fp32
# --------------------------------------------------------------------------------------------------------------
def gemm_nt(a: torch.Tensor, b: torch.Tensor, *, bias=None, residual=None) -> torch.Tensor:
    K = a.shape[1]
    fused = residual is None or b.shape[0] % 32 == 0   # the GEMM epilogue adds a residual only for N % 32 == 0
    y, _, _ = ops.linear(ops.split_planes(a), ops.split_planes(b), K, bias=bias, residual=residual if fused else None,
                         want_planes=False)
    return y if fused else ops.ew_add(y, residual)


def _acc(param: torch.nn.Parameter, grad: torch.Tensor):
    grad = grad.reshape(param.shape).to(param.dtype)
    param.grad = grad.clone() if param.grad is None else param.grad + grad


class _Linear:
    """y = x W^T (+ b), W in nn.Linear layout [N, K]; saves x for dW."""

    def __init__(self, w: torch.Tensor, b: Optional[torch.Tensor] = None):
        self.w, self.b = w.detach().float().contiguous(), (b.detach().float() if b is not None else None)

    def fwd(self, x, residual=None):
        self.x = x
        return gemm_nt(x, self.w, bias=self.b, residual=residual)

    def bwd(self, dy, add_to_dx=None):
        """-> (dx (+ add_to_dx), dW [N, K], db or None)"""
        dx = gemm_nt(dy, ops.transpose(self.w), residual=add_to_dx)
        dw = gemm_nt(ops.transpose(dy), ops.transpose(self.x))
        db = ops.col_sum(dy) if self.b is not None else None
        return dx, dw, db


# --------------------------------------------------------------------------------------------------------------
# encoders
# --------------------------------------------------------------------------------------------------------------
class _XLNetGraph:
    """HF:xlnet:979-1205 as exercised by the reference (oracle: xlnet_forward_restated)."""

    # dropout sites (HF:xlnet): 0 = the input rows (:1085 word_emb_k, :1090 word_emb_q), 5 = the returned rows (:1180);
    # per layer l, base 16 (l + 1): +0 the relative-position rows R_l (HF drops pos_emb per batch element BEFORE its
    # projection, :1159 -- here the projected table shared by the whole batch is dropped instead, which keeps the
    # one-R-per-layer precomputation; a different draw of the same regulariser), +1 the attention probabilities
    # (:129), +2 the output projection before the residual (:147), +3 after the activation (:300), +4 after layer_2
    # (:302).  Active when the encoder module is in train() mode and its config carries a rate.
    def __init__(self, enc: XLNetEncoder):
        self.enc = enc
        cfg = enc.config
        self.d, self.H, self.eps = cfg.d_model, cfg.n_head, float(cfg.layer_norm_eps)
        self.seed = 0

    def _rate(self) -> float:
        return float(getattr(self.enc.config, "dropout", 0.0) or 0.0) if self.enc.training else 0.0

    def fwd(self, x: torch.Tensor, B: int, L: int, plm_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``plm_mask`` [B, L, L]: permutation language modeling -- the content stream (x) and the query stream (mask_emb
        in every row) go through the layers stacked as [2 B L, d] rows; the query stream's rows are returned."""
        d, H = self.d, self.H
        self.B, self.L, self.plm_mask = B, L, plm_mask
        self.tape = []
        p = self.p = self._rate()
        D = (lambda t, site: ops.dropout(t, p, self.seed, site))
        R_all = ops.rel_pos_proj([lyr.rel_attn.r.detach().reshape(d, d).contiguous() for lyr in self.enc.layer], L, d)
        h = x
        if plm_mask is not None:
            h = torch.cat([x, self.enc.mask_emb.detach().reshape(1, d).float().expand(B * L, d)], dim=0).contiguous()
        if p:
            h = D(h, 0)
        for li, lyr in enumerate(self.enc.layer):
            ra, ff = lyr.rel_attn, lyr.ff
            s0 = 16 * (li + 1)
            t = {}
            wqkv = torch.cat([q.detach().reshape(d, d).t() for q in (ra.q, ra.k, ra.v)], dim=0)   # [3d, d]
            t["qkv_lin"] = _Linear(wqkv)
            qkv = t["qkv_lin"].fwd(h)
            t["qkv"], t["R"] = qkv, (D(R_all[li], s0) if p else R_all[li])
            t["rw"], t["rr"] = ra.r_w_bias.detach().reshape(-1).float(), ra.r_r_bias.detach().reshape(-1).float()
            if p:
                a = ops.attn_drop_fwd(qkv, t["R"], t["rw"], t["rr"], B, L, H, (p, self.seed, s0 + 1), plm_mask=plm_mask)
            else:
                a = (ops.xlnet_attn_fwd(qkv, t["R"], t["rw"], t["rr"], B, L, H) if plm_mask is None else
                     ops.xlnet_attn_plm_fwd(qkv, t["R"], t["rw"], t["rr"], B, L, H, plm_mask))
            t["o_lin"] = _Linear(ra.o.detach().reshape(d, d))
            h1_pre = ops.ew_add(D(t["o_lin"].fwd(a), s0 + 2), h) if p else t["o_lin"].fwd(a, residual=h)
            t["h1_pre"] = h1_pre
            h1 = ops.layer_norm_fwd(h1_pre, ra.layer_norm.weight.detach(), ra.layer_norm.bias.detach(), self.eps)
            t["w1"] = _Linear(ff.layer_1.weight, ff.layer_1.bias)
            ffp = t["w1"].fwd(h1)
            t["ffp"] = ffp
            g = ops.act_fwd(_lib.ACT_GELU, ffp)
            if p:
                g = D(g, s0 + 3)
            t["w2"] = _Linear(ff.layer_2.weight, ff.layer_2.bias)
            y_pre = ops.ew_add(D(t["w2"].fwd(g), s0 + 4), h1) if p else t["w2"].fwd(g, residual=h1)
            t["y_pre"] = y_pre
            h = ops.layer_norm_fwd(y_pre, ff.layer_norm.weight.detach(), ff.layer_norm.bias.detach(), self.eps)
            self.tape.append(t)
        out = h if plm_mask is None else h[B * L:].contiguous()
        return D(out, 5) if p else out

    def bwd(self, dh: torch.Tensor) -> torch.Tensor:
        d, H, B, L = self.d, self.H, self.B, self.L
        p = self.p
        D = (lambda t, site: ops.dropout(t, p, self.seed, site))   # the same mask on the gradient = the backward
        if p:
            dh = D(dh, 5)
        if self.plm_mask is not None:      # only the query stream's output was used
            dh = torch.cat([torch.zeros_like(dh), dh], dim=0).contiguous()
        for li in reversed(range(len(self.enc.layer))):
            lyr, t = self.enc.layer[li], self.tape[li]
            ra, ff = lyr.rel_attn, lyr.ff
            s0 = 16 * (li + 1)
            dy_pre, dg2, db2 = ops.layer_norm_bwd(t["y_pre"], ff.layer_norm.weight.detach(), self.eps, dh)
            _acc(ff.layer_norm.weight, dg2); _acc(ff.layer_norm.bias, db2)
            dgel, dw2, dbias2 = t["w2"].bwd(D(dy_pre, s0 + 4) if p else dy_pre)
            _acc(ff.layer_2.weight, dw2); _acc(ff.layer_2.bias, dbias2)
            dffp = ops.act_bwd(_lib.ACT_GELU, t["ffp"], D(dgel, s0 + 3) if p else dgel)
            dh1, dw1, dbias1 = t["w1"].bwd(dffp, add_to_dx=dy_pre)   # + the residual branch of the feed-forward block
            _acc(ff.layer_1.weight, dw1); _acc(ff.layer_1.bias, dbias1)
            dh1_pre, dg1, db1 = ops.layer_norm_bwd(t["h1_pre"], ra.layer_norm.weight.detach(), self.eps, dh1)
            _acc(ra.layer_norm.weight, dg1); _acc(ra.layer_norm.bias, db1)
            da, dwo, _ = t["o_lin"].bwd(D(dh1_pre, s0 + 2) if p else dh1_pre)
            _acc(ra.o, dwo)                                       # o: [d_model, H, dh] == Linear weight [d, HD]
            dqkv, dR, drw, drr = ops.xlnet_attn_bwd(t["qkv"], t["R"], t["rw"], t["rr"], da, B, L, H, plm_mask=self.plm_mask,
                                                    drop=(p, self.seed, s0 + 1) if p else None)
            _acc(ra.r_w_bias, drw); _acc(ra.r_r_bias, drr)
            if p:
                dR = D(dR, s0)
            # R = pos @ Wr  (Wr = r.reshape(d, HD)):  dWr = pos^T dR
            pos = ops.rel_pos_table(L, d, dh.device)
            _acc(ra.r, gemm_nt(ops.transpose(pos), ops.transpose(dR)))
            dh, dwqkv, _ = t["qkv_lin"].bwd(dqkv, add_to_dx=dh1_pre)  # + the residual branch of the attention block
            for j, prm in enumerate((ra.q, ra.k, ra.v)):          # rows [j d, (j+1) d) of the fused weight = W_j^T
                _acc(prm, dwqkv[j * d:(j + 1) * d].t().contiguous())
        if p:
            dh = D(dh, 0)
        if self.plm_mask is not None:      # the query stream started from mask_emb in every row
            M = B * L
            _acc(self.enc.mask_emb, ops.col_sum(dh[M:].contiguous()))
            dh = dh[:M].contiguous()
        return dh


class _GPT2Graph:
    """HF:gpt2:522-636 as exercised by the reference (oracle: gpt2_forward_restated); Conv1D weights are [in, out]."""

    # dropout sites (HF:gpt2; GPT2Config.build sets embd / attn / resid_pdrop to one rate): 0 = after the position
    # embeddings are added (:584 self.drop), per layer l, base 16 (l + 1): +1 the attention probabilities (:66),
    # +2 the attention output projection (:225 resid_dropout), +4 the MLP output projection (:241).
    def __init__(self, enc: GPT2Encoder):
        self.enc = enc
        cfg = enc.config
        self.d, self.H, self.eps = cfg.n_embd, cfg.n_head, float(cfg.layer_norm_epsilon)
        self.seed = 0

    def _rates(self):
        cfg = self.enc.config
        if not self.enc.training:
            return 0.0, 0.0, 0.0
        return tuple(float(getattr(cfg, k, 0.0) or 0.0) for k in ("embd_pdrop", "attn_pdrop", "resid_pdrop"))

    def fwd(self, x: torch.Tensor, B: int, L: int) -> torch.Tensor:
        d, H, enc = self.d, self.H, self.enc
        self.B, self.L = B, L
        self.tape = []
        pe, pa, pr = self.rates = self._rates()
        D = (lambda t, rate, site: ops.dropout(t, rate, self.seed, site))
        h = ops.add_positions(x, enc.wpe.weight.detach(), B, L)
        h = D(h, pe, 0)
        for li, blk in enumerate(enc.h):
            s0 = 16 * (li + 1)
            t = {"h_in": h}
            a = ops.layer_norm_fwd(h, blk.ln_1.weight.detach(), blk.ln_1.bias.detach(), self.eps)
            t["qkv_lin"] = _Linear(blk.attn.c_attn.weight.detach().t(), blk.attn.c_attn.bias)
            qkv = t["qkv_lin"].fwd(a)
            t["qkv"] = qkv
            o = (ops.attn_drop_fwd(qkv, None, None, None, B, L, H, (pa, self.seed, s0 + 1)) if pa else
                 ops.causal_attn_fwd(qkv, B, L, H))
            t["o_lin"] = _Linear(blk.attn.c_proj.weight.detach().t(), blk.attn.c_proj.bias)
            h = ops.ew_add(D(t["o_lin"].fwd(o), pr, s0 + 2), h) if pr else t["o_lin"].fwd(o, residual=h)
            t["h_mid"] = h
            m = ops.layer_norm_fwd(h, blk.ln_2.weight.detach(), blk.ln_2.bias.detach(), self.eps)
            t["fc"] = _Linear(blk.mlp.c_fc.weight.detach().t(), blk.mlp.c_fc.bias)
            fp = t["fc"].fwd(m)
            t["fp"] = fp
            g = ops.act_fwd(_lib.ACT_GELU, fp)
            t["pr"] = _Linear(blk.mlp.c_proj.weight.detach().t(), blk.mlp.c_proj.bias)
            h = ops.ew_add(D(t["pr"].fwd(g), pr, s0 + 4), h) if pr else t["pr"].fwd(g, residual=h)
            self.tape.append(t)
        self.h_last = h
        return ops.layer_norm_fwd(h, enc.ln_f.weight.detach(), enc.ln_f.bias.detach(), self.eps)

    def bwd(self, dout: torch.Tensor) -> torch.Tensor:
        enc, H, B, L = self.enc, self.H, self.B, self.L
        pe, pa, pr = self.rates
        D = (lambda t, rate, site: ops.dropout(t, rate, self.seed, site))
        dh, dg, db = ops.layer_norm_bwd(self.h_last, enc.ln_f.weight.detach(), self.eps, dout)
        _acc(enc.ln_f.weight, dg); _acc(enc.ln_f.bias, db)
        for li in reversed(range(len(enc.h))):
            blk, t = enc.h[li], self.tape[li]
            s0 = 16 * (li + 1)
            dgel, dwp, dbp = t["pr"].bwd(D(dh, pr, s0 + 4))
            _acc(blk.mlp.c_proj.weight, dwp.t()); _acc(blk.mlp.c_proj.bias, dbp)
            dfp = ops.act_bwd(_lib.ACT_GELU, t["fp"], dgel)
            dm, dwf, dbf = t["fc"].bwd(dfp)
            _acc(blk.mlp.c_fc.weight, dwf.t()); _acc(blk.mlp.c_fc.bias, dbf)
            dh, dg2, db2 = ops.layer_norm_bwd(t["h_mid"], blk.ln_2.weight.detach(), self.eps, dm, add=dh)
            _acc(blk.ln_2.weight, dg2); _acc(blk.ln_2.bias, db2)
            do, dwo, dbo = t["o_lin"].bwd(D(dh, pr, s0 + 2))
            _acc(blk.attn.c_proj.weight, dwo.t()); _acc(blk.attn.c_proj.bias, dbo)
            dqkv = ops.causal_attn_bwd(t["qkv"], do, B, L, H, drop=(pa, self.seed, s0 + 1) if pa else None)
            da, dwq, dbq = t["qkv_lin"].bwd(dqkv)
            _acc(blk.attn.c_attn.weight, dwq.t()); _acc(blk.attn.c_attn.bias, dbq)
            dh, dg1, db1 = ops.layer_norm_bwd(t["h_in"], blk.ln_1.weight.detach(), self.eps, da, add=dh)
            _acc(blk.ln_1.weight, dg1); _acc(blk.ln_1.bias, db1)
        # h0 = dropout(x + wpe[:L])
        dh = D(dh, pe, 0)
        _acc_rows(enc.wpe.weight, ops.sum_over_sessions(dh, B, L), L)
        return dh


def _acc_rows(param, rows, n):
    g = torch.zeros_like(param)
    g[:n] = rows
    param.grad = g if param.grad is None else param.grad + g


# --------------------------------------------------------------------------------------------------------------
# the widened input block (SURVEY §8f N4 in training): per-feature LayerNorm, soft embeddings, continuous
# projection, element-wise aggregations
# --------------------------------------------------------------------------------------------------------------
class _MLPGraph:
    """A built MLPBlock (block/mlp.py:30-87): DenseBlocks of Linear (+ ReLU / GELU)."""

    def __init__(self, built):
        self.blocks = list(built)

    def fwd(self, x):
        self.tape = []
        for blk in self.blocks:
            lin = _Linear(blk[0].weight, blk[0].bias)
            pre = lin.fwd(x)
            act = blk.act_code()
            x = ops.act_fwd(act, pre) if act != _lib.ACT_NONE else pre
            self.tape.append((blk[0], lin, pre, act))
        return x

    def bwd(self, dy):
        for mod, lin, pre, act in reversed(self.tape):
            dpre = ops.act_bwd(act, pre, dy) if act != _lib.ACT_NONE else dy
            dy, dw, db = lin.bwd(dpre)
            _acc(mod.weight, dw)
            if mod.bias is not None:
                _acc(mod.bias, db)
        return dy


class _WideInput:
    """Training forward / backward of the input block beyond ``embedding rows + scalars -> concat``: every feature is
    produced as its own fp32 [M, width] matrix (embedding rows: features/embedding.py:226-249; soft embeddings:
    :517-556; the continuous projection MLP: features/tabular.py:88-118), optionally LayerNorm'd
    (tabular/transformations.py:95-141), then aggregated in sorted-name order (tabular/aggregation.py:35-47 concat,
    :139-157 element-wise-sum, :160-193 element-wise-sum-item-multi = item * sum of the others).  The inference
    path does all of this in one kernel (t4r_input_block_fwd); here each piece keeps what its backward needs."""

    def __init__(self, inp, layout, C):
        from .features import ContinuousProjection
        self.inp, self.layout, self.C = inp, layout, C
        self.agg = inp.AGGREGATIONS[inp.aggregation or "concat"]
        cm, cont = inp.categorical_module, inp.continuous_module
        if any(kind == "cat" and cm.is_sharded(n) for n, kind, *_ in layout):
            raise NotImplementedError("FusedTrainingStep: a row-sharded table with the widened input block")
        self.cat_ln = getattr(cm, "post", None) if cm is not None else None
        self.cont_ln = getattr(cont, "post", None) if cont is not None else None
        self.mlp = _MLPGraph(cont.mlp) if isinstance(cont, ContinuousProjection) else None
        if self.agg == _lib.AGG_SUM_ITEM_MULTI and not any(n == cm.item_id for n, *_ in layout):
            raise ValueError("element-wise-sum-item-multi needs the item-id feature")

    def _ln(self, post, name):
        if post is None or name not in post.feature_layer_norm:
            return None
        return post.feature_layer_norm[name]

    def fwd(self, batch, B, L):
        inp = self.inp
        cm, cont = inp.categorical_module, inp.continuous_module
        M = B * L

        def seq(v):   # context features [B] / [B, 1] repeat over the positions (tabular/base.py:53-63)
            if v.dim() == 1 or (v.dim() == 2 and v.shape[1] == 1 and L != 1):
                v = v.reshape(B, 1).expand(B, L)
            return v.reshape(-1)
        self.nodes = []
        for name, kind, col, width in self.layout:
            node = {"name": name, "kind": kind, "col": col, "width": width, "ln": None}
            if kind == "cat":
                table = cm.embedding_tables[name].weight
                node["ids"], node["table"] = seq(batch[name]).contiguous(), table
                y = ops.gather_rows(table.detach().float(), node["ids"])
                node["ln"] = self._ln(self.cat_ln, name)
            elif kind == "cont":
                y = seq(batch[name]).float().reshape(M, 1).contiguous()
            elif kind == "soft":
                mod = cont.embedding_tables[name]
                node["x"], node["mod"] = seq(batch[name]).float().contiguous(), mod
                y, node["p"] = ops.soft_emb_fwd(node["x"], mod.projection_layer.weight.detach(),
                                                mod.projection_layer.bias.detach(), mod.embedding_table.weight.detach())
                node["ln"] = self._ln(self.cont_ln, name)
            else:             # "dense": the continuous projection
                vals = [(seq(batch[n]), i) for i, n in enumerate(cont.features)]
                x, _, _ = ops.embed_concat([], vals, M, len(vals), want_f32=True, want_planes=False)
                y = self.mlp.fwd(x)
            if node["ln"] is not None:
                node["pre_ln"] = y
                ln = node["ln"]
                y = ops.layer_norm_fwd(y, ln.weight.detach(), ln.bias.detach(), float(ln.eps))
            node["y"] = y
            self.nodes.append(node)
        if self.agg == _lib.AGG_CONCAT:
            out = torch.empty((M, self.C), dtype=torch.float32, device=self.nodes[0]["y"].device)
            for nd in self.nodes:      # column placement of whole feature matrices: plumbing
                out[:, nd["col"]:nd["col"] + nd["width"]] = nd["y"]
            return out
        item = cm.item_id if self.agg == _lib.AGG_SUM_ITEM_MULTI else None
        acc = None
        for nd in self.nodes:
            if nd["name"] == item:
                continue
            acc = nd["y"] if acc is None else ops.ew_add(acc, nd["y"])
        if item is None:
            return acc
        self.item_y = next(nd["y"] for nd in self.nodes if nd["name"] == item)
        self.others = acc
        return ops.ew_mul(self.item_y, acc)

    def bwd(self, dagg):
        inp = self.inp
        cm = inp.categorical_module
        item = cm.item_id if self.agg == _lib.AGG_SUM_ITEM_MULTI else None
        if item is not None:
            d_item, d_others = ops.ew_mul(dagg, self.others), ops.ew_mul(dagg, self.item_y)
        for nd in self.nodes:
            kind = nd["kind"]
            if kind == "cont":
                continue
            if self.agg == _lib.AGG_CONCAT:
                dy = dagg[:, nd["col"]:nd["col"] + nd["width"]].contiguous()
            elif item is None:
                dy = dagg
            else:
                dy = d_item if nd["name"] == item else d_others
            if nd["ln"] is not None:
                ln = nd["ln"]
                dy, dg, db = ops.layer_norm_bwd(nd["pre_ln"], ln.weight.detach(), float(ln.eps), dy)
                _acc(ln.weight, dg); _acc(ln.bias, db)
            if kind == "cat":
                param = nd["table"]
                g = torch.zeros_like(param.detach(), dtype=torch.float32) if param.grad is None else param.grad
                ops.index_add_rows(g, nd["ids"], dy, 0, nd["width"], skip_index=inp.masking.padding_idx)
                param.grad = g
            elif kind == "soft":
                mod = nd["mod"]
                table = mod.embedding_table.weight
                dl, dlx = ops.soft_emb_bwd(nd["x"], table.detach(), nd["p"], dy)
                _acc(table, gemm_nt(ops.transpose(nd["p"]), ops.transpose(dy)))          # p^T dOut  [n, dim]
                _acc(mod.projection_layer.weight, ops.col_sum(dlx))
                _acc(mod.projection_layer.bias, ops.col_sum(dl))
            else:
                self.mlp.bwd(dy)


# --------------------------------------------------------------------------------------------------------------
# the whole step
# --------------------------------------------------------------------------------------------------------------
class FusedTrainingStep:
    def __init__(self, model, head_chunk: int = 32768, want_rank: bool = False):
        # want_rank: also produce the label ranks among the training logits (replicated full softmax), so that the
        # streaming ranking metrics of ``Model.fit(compute_metric=True)`` (model/base.py:704-707) need no [T, V] logits
        self.want_rank = bool(want_rank)
        self.row_rank = None
        if len(model.heads) != 1 or len(model.heads[0].prediction_task_dict) != 1:
            raise NotImplementedError("FusedTrainingStep: one head with one NextItemPredictionTask")
        head = model.heads[0]
        self.inputs, self.tblock = head.body[0], head.body[1]
        self.task = next(iter(head.prediction_task_dict.values()))
        inp, task = self.inputs, self.task
        layout, self.C = inp._layout()
        cat_post = getattr(inp.categorical_module, "post", None)
        plain = ((inp.aggregation or "concat") == "concat" and all(kind in ("cat", "cont") for _, kind, *_ in layout)
                 and not (cat_post is not None and len(cat_post.feature_layer_norm) > 0))
        self.wide = None if plain else _WideInput(inp, layout, self.C)
        from .features import StochasticSwapNoise
        if (inp.pre is not None and not isinstance(inp.pre, StochasticSwapNoise)) or (inp.projection_module is not None and inp._projection_linear() is None) or (
                plain and inp._projection_linear() is None):
            raise NotImplementedError("FusedTrainingStep: the default Linear (+ReLU) projection; StochasticSwapNoise as "
                                      "the only pre-transform")
        if not isinstance(inp.masking, (MaskedLanguageModeling, CausalLanguageModeling, PermutationLanguageModeling)):
            raise NotImplementedError("FusedTrainingStep: MLM, CLM or PLM masking")
        if not task.weight_tying:
            raise NotImplementedError("FusedTrainingStep: tied weights")
        task.output_weight()             # refreshes task.item_embedding_table (the table may have been sharded after build)
        self.sharded = task._sharded()   # row-sharded item table (BASELINE configs 4-5): see _forward/_backward_sharded
        enc = self.tblock.transformer
        self.graph = _XLNetGraph(enc) if isinstance(enc, XLNetEncoder) else _GPT2Graph(enc)
        self.layout = layout
        self.head_chunk = int(head_chunk)
        # Dropout (config/transformer.py:217-260, :432-482 default to 0.3): applied at HF's sites whenever the encoder
        # module is in train() mode, as in the reference (Model.fit calls self.train(), model/base.py:692); the masks
        # are counter-based (seed, step, site), regenerated in the backward.  ``set_dropout_seed`` pins the stream.
        self.dropout_seed = int(torch.initial_seed()) & ((1 << 62) - 1)
        self.dropout_step = 0

    def set_dropout_seed(self, seed: int, step: int = 0):
        self.dropout_seed, self.dropout_step = int(seed) & ((1 << 62) - 1), int(step)
        return self

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        inp, task = self.inputs, self.task
        self.graph.seed = (self.dropout_seed + 0x9E3779B97F4A7C15 * self.dropout_step) & ((1 << 64) - 1)
        self.dropout_step += 1
        cm = inp.categorical_module
        if inp.pre is not None:    # StochasticSwapNoise (tabular/transformations.py:29-92): a permutation of the RAW
            batch = inp.pre(dict(batch))   # inputs, constant w.r.t. every parameter -- applied once, then forgotten
        ids = batch[cm.item_id]
        B, L = ids.shape
        M = B * L
        self.B, self.L, self.M = B, L, M
        cm.item_seq = ids
        self.row_rank = None
        inp.masking.compute_masked_targets(ids, training=True, testing=False)
        code = inp.masking.row_code.reshape(-1)
        self.code = code
        # K1 gather + concat (fp32 rows; the gather is recomputed in the backward instead of being kept)
        cats, conts = [], []
        for name, kind, col, width in (self.layout if self.wide is None else ()):
            v = batch[name]
            if v.dim() == 1 or (v.dim() == 2 and v.shape[1] == 1 and L != 1):
                # a context feature ([B] / [B, 1]) is repeated for every position, as in the forward-only path
                # (features.py seq(); tabular/base.py:53-63 in the reference)
                v = v.reshape(B, 1).expand(B, L)
            v = v.reshape(-1)
            if v.numel() != M:
                raise ValueError(f"feature {name!r} has {v.numel()} values, expected batch x length = {M}")
            if kind == "cat":
                cats.append((cm.embedding_tables[name].weight.detach(), v, col))
            else:
                conts.append((v, col))
        self.cats, self.conts = cats, conts
        self.item_plan = None
        if self.wide is not None:
            concat = self.wide.fwd(batch, B, L)
        elif self.sharded:
            # the item rows come through the table's exchange (all-gather of ids + one all-to-all); the other features
            # are gathered locally into the same [M, C] buffer
            table = cm.embedding_tables[cm.item_id]
            item = next(c for c in cats if c[0].data_ptr() == table.weight.data_ptr())
            local = [c for c in cats if c is not item]
            plans: list = []
            rows, _ = table.lookup(batch[cm.item_id], plan_out=plans)
            self.item_plan, self.item_col = plans[0], item[2]
            if local or conts:
                concat, _, _ = ops.embed_concat(local, conts, M, self.C, want_f32=True, want_planes=False)
            else:
                concat = torch.empty((M, self.C), dtype=torch.float32, device=rows.device)
            concat[:, item[2]:item[2] + rows.shape[1]] = rows
            self.cats = local
        else:
            concat, _, _ = ops.embed_concat(cats, conts, M, self.C, want_f32=True, want_planes=False)
        # K2 projection + activation, then the mask replace (apply_mask_to_inputs)
        lin = inp._projection_linear()
        if lin is None:          # element-wise aggregation straight into the encoder (widened block only)
            self.proj, y = None, concat
        else:
            self.proj = _Linear(lin.weight, lin.bias)
            self.proj_pre = self.proj.fwd(concat)
            self.proj_act = inp._projection_act()
            y = ops.act_fwd(self.proj_act, self.proj_pre) if self.proj_act != _lib.ACT_NONE else self.proj_pre
        x0 = ops.apply_row_codes(y, code, inp.masking.masked_item_embedding.detach().float())
        # encoder (PLM: XLNet's two-stream forward under the permutation mask)
        if isinstance(inp.masking, PermutationLanguageModeling):
            h = self.graph.fwd(x0, B, L, plm_mask=inp.masking.perm_mask)
        else:
            h = self.graph.fwd(x0, B, L)
        # head: label rows, optional task_block, fused loss (keeps the per-row log-sum-exp)
        self.tgt_rows, self.labels, count = ops.compact_targets(inp.masking.masked_targets, task.padding_idx)
        T = int(count.item())                                  # the reference's masked_select synchronises too
        self.T = T
        xt = ops.gather_rows(h, self.tgt_rows[:T])
        self.tb = []
        if task.task_block is not None:
            for blk in task.task_block:
                if blk.act_code() != _lib.ACT_NONE:
                    raise NotImplementedError("FusedTrainingStep: task_block without activation (the reference default)")
                lin_tb = _Linear(blk[0].weight, blk[0].bias)
                xt = lin_tb.fwd(xt)
                self.tb.append((blk[0], lin_tb))
        self.xt = xt
        W = task.output_weight().detach().float()
        inv_tau = task._inv_tau()
        y_lab = self.labels[:T]
        self.sampled = bool(task.sampled_softmax)
        self.smooth = float(task.label_smoothing or 0.0)
        if self.sampled and self.sharded:
            raise NotImplementedError("FusedTrainingStep: sampled softmax over a row-sharded table")
        if self.smooth and (self.sampled or self.sharded):
            raise NotImplementedError("FusedTrainingStep: label smoothing with the replicated full softmax only")
        if self.sampled:
            # model/prediction_task.py:673-696: the positive's logit + the S sampled negatives, logQ-corrected
            neg, _, _ = task.sampler.sample(y_lab[:1], raw_draws=task._neg_draws)
            self.neg = neg
            nlq = task.sampler.neg_log_q
            self.col_bias = nlq[neg].contiguous()
            self.neg_rows = ops.gather_rows(W, neg)
            self.pos = ops.label_logit(xt, W, y_lab, class_bias=nlq, inv_temperature=inv_tau)
            res = ops.head_softmax_ce(ops.split_planes(xt), xt, y_lab, ops.split_planes(self.neg_rows), None,
                                      inv_temperature=inv_tau, col_bias=self.col_bias, col_ids=neg, col_ids_sorted_unique=True,
                                      hit_value=float(torch.finfo(torch.float16).min / 100.0), pos_logit=self.pos)
            self.row_lse = res["row_lse"]
            self.loss = res["loss"].reshape(())
            return self.loss
        if self.sharded:
            # forward and backward of the sharded head in one go (its only upstream gradient is d loss = 1)
            from . import distributed as D
            table = task.item_embedding_table
            self.loss, self.dxt_sharded, self.dW_sharded = D.sharded_softmax_ce_train(
                xt, y_lab, W, table.num_embeddings, table.group, w_planes=None, inv_tau=inv_tau,
                head_chunk=self.head_chunk, head_rows=getattr(table, "train_head_rows", None),
                train_head=getattr(table, "train_head", None))
            return self.loss
        res = ops.head_softmax_ce(ops.split_planes(xt), xt, y_lab, ops.split_planes(W), W, inv_temperature=inv_tau,
                                  label_smoothing=self.smooth, want_rank=self.want_rank)
        self.row_rank = res["row_rank"][:T] if self.want_rank else None
        self.row_lse = res["row_lse"]
        self.loss = res["loss"].reshape(())
        return self.loss

    # ------------------------------------------------------------------------------------------------ backward
    def backward(self, grad_loss: float = 1.0):
        inp, task = self.inputs, self.task
        cm = inp.categorical_module
        T, M = self.T, self.M
        Wp = task.output_weight()
        W = Wp.detach().float()
        V, De = W.shape
        inv_tau = task._inv_tau()
        y_lab = self.labels[:T]
        if self.sharded:
            return self._backward_sharded(float(grad_loss))
        scale = float(grad_loss) / max(T, 1)
        if self.sampled:
            return self._backward_sampled(Wp, W, y_lab, inv_tau, scale)
        dxt = torch.zeros_like(self.xt)
        dW = torch.zeros_like(W)
        xt_t = ops.transpose(self.xt)                                   # [De, T]
        xt_planes = ops.split_planes(self.xt)
        for v0 in range(0, V, self.head_chunk):
            v1 = min(V, v0 + self.head_chunk)
            Wc = W[v0:v1].contiguous()
            z = ops.head_logits(xt_planes, ops.split_planes(Wc), De, inv_temperature=inv_tau)          # [T, Vc]
            P = ops.softmax_ce_bwd(z, self.row_lse, y_lab, v0, scale * inv_tau,   # (softmax - target) * dL/dz scale
                                   label_smoothing=self.smooth, V_total=V)
            dxt = gemm_nt(P, ops.transpose(Wc), residual=dxt)              # dX_t += P W_c
            dW[v0:v1] = gemm_nt(ops.transpose(P), xt_t)                      # dW_c = P^T X_t
        _acc(Wp, dW)                                                        # tied: the item table's grad starts here
        self._backward_body(dxt)
        return self.loss

    def _backward_sampled(self, Wp, W, y_lab, inv_tau, scale):
        """Sampled softmax: the logits [T, 1 + S] are small enough to be recomputed whole.  P_neg = softmax over the
        negatives (0 at accidental hits: constants in the forward), P_pos = softmax(positive) - 1; the logQ terms are
        constants.  dX_t = (P_neg W_neg + P_pos W[y]) / tau; dW gets P_neg^T X_t at the negatives' rows and P_pos X_t at
        the labels' rows (scatter-add: duplicates among the labels)."""
        T, De = self.xt.shape
        z = ops.head_logits(ops.split_planes(self.xt), ops.split_planes(self.neg_rows), De, inv_temperature=inv_tau)
        P = ops.sampled_ce_bwd(z, self.row_lse, y_lab, self.col_bias, self.neg, inv_tau, scale * inv_tau)
        p_pos = (torch.exp(self.pos[:T] - self.row_lse[:T]) - 1.0) * (scale * inv_tau)      # [T]: tiny, plain torch
        w_pos = ops.gather_rows(W, y_lab)
        dxt = gemm_nt(P, ops.transpose(self.neg_rows), residual=(w_pos * p_pos.unsqueeze(1)).contiguous())
        dW = torch.zeros_like(W)
        dneg = gemm_nt(ops.transpose(P), ops.transpose(self.xt))                              # [S, De]
        ops.index_add_rows(dW, self.neg, dneg, 0, De)
        ops.index_add_rows(dW, y_lab, (self.xt * p_pos.unsqueeze(1)).contiguous(), 0, De)
        _acc(Wp, dW)
        self._backward_body(dxt)
        return self.loss

    def _backward_sharded(self, grad_loss: float):
        """Row-sharded table: the head's gradients were produced with the forward; the item rows' gradients travel back
        through the transposed all-to-all; the replicated parameters' gradients are summed over the ranks (every rank
        back-propagated its own sessions of the global mean loss)."""
        import torch.distributed as dist

        from . import distributed as D
        task, inp = self.task, self.inputs
        table = task.item_embedding_table
        Wp = table.weight
        _acc(Wp, self.dW_sharded * grad_loss)
        dconcat = self._backward_body(self.dxt_sharded * grad_loss)
        De = Wp.shape[1]
        g = torch.zeros_like(Wp.detach(), dtype=torch.float32)
        pad_local = inp.masking.padding_idx - table.lo
        D.sharded_embedding_lookup_bwd(dconcat[:, self.item_col:self.item_col + De].contiguous(), self.item_plan, g,
                                       table.group, skip_local_index=pad_local if 0 <= pad_local < Wp.shape[0] else -1,
                                       scatter_rows=getattr(table, "train_scatter_rows", None),
                                       index_add=getattr(table, "train_index_add", None))
        _acc(Wp, g)
        sharded_ptr = Wp.data_ptr()
        for p in self._all_params():
            if p.grad is not None and p.data_ptr() != sharded_ptr:
                dist.all_reduce(p.grad, group=table.group)
        return self.loss

    def _all_params(self):
        seen = set()
        for mod in (self.inputs, self.tblock, self.task):
            for p in mod.parameters():
                if p.data_ptr() not in seen:
                    seen.add(p.data_ptr())
                    yield p

    def _backward_body(self, dxt):
        """Everything below the head: task_block, label-row scatter, encoder, mask replace, projection, embedding rows.
        Returns d concat [M, C]."""
        inp, task = self.inputs, self.task
        cm = inp.categorical_module
        M, T = self.M, self.T
        for lin_mod, lin_tb in reversed(self.tb):
            dxt, dwt, dbt = lin_tb.bwd(dxt)
            _acc(lin_mod.weight, dwt)
            if lin_mod.bias is not None:
                _acc(lin_mod.bias, dbt)
        dh = ops.scatter_rows(dxt, self.tgt_rows[:T], M)                    # zero outside the label rows
        dx0 = self.graph.bwd(dh)
        # mask replace: rows with code 1 took masked_item_embedding, code 2 are constant zero
        dmask, dy = ops.row_codes_bwd(dx0, self.code)
        _acc(inp.masking.masked_item_embedding, dmask)
        if self.proj is None:
            dconcat = dy
        else:
            dpre = ops.act_bwd(self.proj_act, self.proj_pre, dy) if self.proj_act != _lib.ACT_NONE else dy
            dconcat, dwp, dbp = self.proj.bwd(dpre)
            lin = inp._projection_linear()
            _acc(lin.weight, dwp)
            if lin.bias is not None:
                _acc(lin.bias, dbp)
        if self.wide is not None:
            self.wide.bwd(dconcat)
            return dconcat
        # embedding rows: scatter-add the column slice of every categorical feature (padding row gets no gradient)
        for table, ids, col in self.cats:
            param = next(p for p in cm.parameters() if p.data_ptr() == table.data_ptr())
            g = torch.zeros_like(table) if param.grad is None else param.grad
            ops.index_add_rows(g, ids.reshape(-1), dconcat, col, table.shape[1], skip_index=inp.masking.padding_idx)
            param.grad = g
        return dconcat


class _FusedLossFn(torch.autograd.Function):
    """The loss of a step that has ALREADY run forward and backward, as one autograd node over exactly the parameters
    that received a gradient: ``backward`` scales the parked gradients by the incoming one and hands them to autograd
    (AccumulateGrad, hooks -- DistributedDataParallel's included).  Parameters the path does not touch (XLNet's
    ``seg_embed`` / ``r_s_bias``, HF:xlnet:225-233, unused without token types) are not inputs of the node, so
    ``find_unused_parameters`` sees them exactly as it does in the reference's autograd graph."""

    @staticmethod
    def forward(ctx, loss, grads, *params):
        ctx.grads = grads
        return loss.clone()

    @staticmethod
    def backward(ctx, grad_out):
        return (None, None) + tuple(g * grad_out for g in ctx.grads)


def training_loss(model, batch, step: Optional[FusedTrainingStep] = None) -> torch.Tensor:
    """Differentiable training loss of the fused path: ``training_loss(model, batch).backward()`` fills ``.grad``.
    The step's forward and backward run here, eagerly (gradients parked aside, the parameters' own ``.grad`` untouched);
    the returned scalar carries them as one autograd node."""
    step = step or FusedTrainingStep(model)
    params = [p for p in model.parameters() if p.requires_grad]
    saved = [p.grad for p in params]
    for p in params:
        p.grad = None
    try:
        with torch.no_grad():
            loss = step.forward(batch)
            step.backward(1.0)
        grads = [p.grad for p in params]
    finally:
        for p, g in zip(params, saved):
            p.grad = g
    used = [(p, g) for p, g in zip(params, grads) if g is not None]
    return _FusedLossFn.apply(loss, [g for _, g in used], *[p for p, _ in used])


class FusedAdamW(torch.optim.Optimizer):
    """AdamW whose update runs in the t4r kernel (``t4r_train_adamw``: torch.optim.AdamW's rule, decoupled weight decay,
    one element per thread, in place on the parameter and its two moment buffers).  Same constructor arguments and
    state-dict layout (``step`` / ``exp_avg`` / ``exp_avg_sq``) as ``torch.optim.AdamW``."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
                st["step"] += 1
                g = p.grad.float().contiguous()
                data = p.data if p.data.is_contiguous() else p.data.contiguous()
                ops.adamw_step(data.view(-1), g.view(-1), st["exp_avg"].view(-1), st["exp_avg_sq"].view(-1), group["lr"],
                               b1, b2, group["eps"], group["weight_decay"], st["step"])
                if data.data_ptr() != p.data.data_ptr():
                    p.data.copy_(data)
                else:
                    # the kernel wrote through a raw pointer: tell torch (and with it every cache keyed on
                    # ``param._version``, e.g. ops.PlaneCache's split planes of this weight) that the values moved
                    torch.autograd.graph.increment_version(p)
        return loss
