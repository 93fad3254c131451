"""ctypes binding of ``libt4r_b200.so`` (the C ABI declared in ``include/t4r_b200.h``).

The shared library is built in-tree by ``transformers4rec_b200.build()`` (nvcc,
``-gencode arch=compute_100a,code=sm_100a``).  There is no fallback of any kind:
if the library is missing, or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# T4R_LIB_PATH: developer switch for A/B runs of an alternative build of the same sources (never set in production)
LIB_PATH = os.environ.get("T4R_LIB_PATH") or os.path.join(_HERE, "libt4r_b200.so")
CSRC = os.path.join(_HERE, "csrc")

T4R_MAX_FEATURES = 32

# modes (include/t4r_b200.h)
MLM_TRAIN, MLM_EVAL_LAST, MLM_EVAL_ALL, MLM_INFERENCE = 0, 1, 2, 3
CLM_ALL, CLM_LAST, CLM_INFERENCE = 0, 1, 2
PLM_TRAIN, PLM_EVAL_LAST, PLM_EVAL_ALL = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2

c_void_p, c_int, c_int64, c_float, c_size_t = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t


class T4RError(RuntimeError):
    pass


class FeatureList(C.Structure):
    _fields_ = [
        ("n_cat", c_int),
        ("n_cont", c_int),
        ("table", c_void_p * T4R_MAX_FEATURES),
        ("ids", c_void_p * T4R_MAX_FEATURES),
        ("table_rows", c_int64 * T4R_MAX_FEATURES),
        ("dim", c_int * T4R_MAX_FEATURES),
        ("cat_col", c_int * T4R_MAX_FEATURES),
        ("cont", c_void_p * T4R_MAX_FEATURES),
        ("cont_col", c_int * T4R_MAX_FEATURES),
    ]


class Feature(C.Structure):
    _fields_ = [
        ("kind", c_int), ("dim", c_int), ("col", c_int), ("card", c_int), ("per_session", c_int), ("reserved", c_int),
        ("input", c_void_p), ("table", c_void_p), ("soft_w", c_void_p), ("soft_b", c_void_p),
        ("ln_gamma", c_void_p), ("ln_beta", c_void_p),
    ]


FEAT_CAT, FEAT_CONT, FEAT_SOFT, FEAT_DENSE = 0, 1, 2, 3
AGG_CONCAT, AGG_SUM, AGG_SUM_ITEM_MULTI = 0, 1, 2
METRIC_RECALL, METRIC_PRECISION, METRIC_RR, METRIC_DCG = 0, 1, 2, 3


class LinearArgs(C.Structure):
    _fields_ = [
        ("M", c_int64),
        ("N", c_int),
        ("K", c_int),
        ("x_planes", c_void_p),
        ("w_planes", c_void_p),
        ("m_dev", c_void_p),
        ("bias", c_void_p),
        ("act", c_int),
        ("row_code", c_void_p),
        ("mask_vec", c_void_p),
        ("residual", c_void_p),
        ("ln_gamma", c_void_p),
        ("ln_beta", c_void_p),
        ("ln_eps", c_float),
        ("out_pre_ln", c_void_p),
        ("out_f32", c_void_p),
        ("out_planes", c_void_p),
        ("nprod", c_int),
    ]


class XLNetLayer(C.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "wqkv_planes", "wr", "r_w_bias", "r_r_bias", "wo_planes", "ln1_gamma", "ln1_beta",
        "w1_planes", "b1", "w2_planes", "b2", "ln2_gamma", "ln2_beta")]


class GPT2Layer(C.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "ln1_gamma", "ln1_beta", "wqkv_planes", "bqkv", "wo_planes", "bo", "ln2_gamma", "ln2_beta",
        "w1_planes", "b1", "w2_planes", "b2")]


class HeadArgs(C.Structure):
    _fields_ = [
        ("T_cap", c_int),
        ("t_dev", c_void_p),
        ("De", c_int),
        ("V", c_int64),
        ("xt_planes", c_void_p),
        ("xt_f32", c_void_p),
        ("labels", c_void_p),
        ("w_planes", c_void_p),
        ("w_f32", c_void_p),
        ("inv_temperature", c_float),
        ("col_bias", c_void_p),
        ("col_ids", c_void_p),
        ("hit_value", c_float),
        ("pos_logit", c_void_p),
        ("v_offset", c_int64),
        ("row_lse", c_void_p),
        ("row_tgt", c_void_p),
        ("row_loss", c_void_p),
        ("loss", c_void_p),
        ("row_rank", c_void_p),
        ("workspace", c_void_p),
        ("workspace_bytes", c_size_t),
        ("nprod", c_int),
        ("ev_gemm_start", c_void_p),
        ("ev_gemm_stop", c_void_p),
        ("label_smoothing", c_float),
        ("rank_tgt", c_void_p),
        ("xt_inv_scale", c_void_p),
        ("w_inv_scale", c_void_p),
        ("col_ids_sorted_unique", c_int),
    ]


T4R_MAX_PEERS, T4R_PEER_HANDLE_BYTES = 16, 64


class PeerPtrs(C.Structure):
    _fields_ = [("base", c_void_p * T4R_MAX_PEERS), ("world", c_int), ("rank", c_int)]


# name -> (restype, argtypes); every symbol include/t4r_b200.h declares
_P = c_void_p
SIGNATURES = {
    "t4r_last_error": (C.c_char_p, []),
    "t4r_version": (c_int, []),
    "t4r_launch_count": (C.c_longlong, []),
    "t4r_sizeof_struct": (c_size_t, [c_int]),
    "t4r_head_args_last_offset": (c_size_t, []),
    "t4r_embed_concat_fwd": (c_int, [C.POINTER(FeatureList), c_int64, c_int, _P, _P, _P, _P]),
    "t4r_pad_ragged": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, _P, _P]),
    "t4r_mask_mlm": (c_int, [_P, c_int, c_int, c_int64, c_int, c_float, _P, _P, _P, _P, _P]),
    "t4r_mask_clm": (c_int, [_P, c_int, c_int, c_int64, c_int, _P, _P, _P, _P]),
    "t4r_mask_plm": (c_int, [_P, c_int, c_int, c_int64, c_int, c_int, C.POINTER(C.c_int32), _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "t4r_debug_mask_plm_host": (c_int, [_P, c_int, c_int, c_int64, c_int, c_int, C.POINTER(C.c_int32), _P, _P, _P, _P, _P,
                                        _P, _P, _P]),
    "t4r_xlnet_encoder_plm_fwd": (c_int, [C.POINTER(XLNetLayer), c_int, c_int, c_int, c_int, c_int, c_float, _P, _P, _P,
                                          _P, c_size_t, _P]),
    "t4r_compact_targets": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P]),
    "t4r_split_planes": (c_int, [_P, c_int64, c_int, c_int, _P, _P, _P, _P, _P]),
    "t4r_split_planes_mixed": (c_int, [_P, c_int64, c_int, c_int, _P, _P, _P]),
    "t4r_split_planes_mixed_n": (c_int, [_P, c_int64, c_int, c_int, _P, _P, _P, _P]),
    "t4r_debug_split_planes_mixed_host": (c_int, [_P, c_int64, c_int, c_int, _P, _P]),
    "t4r_gather_rows_split": (c_int, [_P, c_int, c_int, _P, _P, c_int, _P, _P, _P]),
    "t4r_gather_rows_split_i64": (c_int, [_P, c_int, c_int, _P, c_int, _P, _P, _P]),
    "t4r_linear_fwd": (c_int, [C.POINTER(LinearArgs), _P]),
    "t4r_input_block_fwd": (c_int, [C.POINTER(Feature), c_int, c_int64, c_int, c_int, c_int, c_float, c_int, _P, _P, _P, _P]),
    "t4r_swap_noise": (c_int, [_P, c_int, c_int64, _P, c_int64, c_int, _P, c_float, _P, _P, _P, _P]),
    "t4r_metrics_from_ranks": (c_int, [_P, _P, c_int, c_int, C.POINTER(C.c_int32), c_int, _P, _P]),
    "t4r_ffn_fwd": (c_int, [_P, c_int64, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, c_float, _P, _P, _P, _P]),
    "t4r_debug_ts_mma": (c_int, [_P, _P, c_int, _P, _P]),
    "t4r_debug_sgemm_nt": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "t4r_xlnet_encoder_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "t4r_xlnet_encoder_fwd": (c_int, [C.POINTER(XLNetLayer), c_int, c_int, c_int, c_int, c_int, c_float, _P, _P, _P,
                                      _P, _P, c_size_t, _P]),
    "t4r_gpt2_encoder_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "t4r_gpt2_encoder_fwd": (c_int, [C.POINTER(GPT2Layer), c_int, c_int, c_int, c_int, c_int, c_float, _P, _P, _P, _P,
                                     _P, _P, _P, c_size_t, _P]),
    "t4r_head_workspace_bytes": (c_size_t, [c_int, c_int64, c_int]),
    "t4r_head_softmax_ce_fwd": (c_int, [C.POINTER(HeadArgs), _P]),
    "t4r_label_logit": (c_int, [_P, _P, _P, c_int, _P, c_int, c_int64, _P, c_float, c_int64, _P, _P]),
    "t4r_head_logits": (c_int, [_P, _P, c_int, _P, c_int64, c_int, c_float, _P, c_int64, c_int, _P]),
    "t4r_head_logits_mixed": (c_int, [_P, _P, c_int, _P, c_int64, c_int, c_float, _P, c_int64, _P, _P, _P]),
    "t4r_train_transpose": (c_int, [_P, c_int64, c_int64, _P, _P, c_int]),
    "t4r_train_act_fwd": (c_int, [c_int, _P, _P, c_int64, _P, c_int]),
    "t4r_train_act_bwd": (c_int, [c_int, _P, _P, _P, c_int64, _P, c_int]),
    "t4r_train_add_positions": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, c_int]),
    "t4r_train_sum_sessions": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_int]),
    "t4r_train_row_codes_fwd": (c_int, [_P, _P, _P, c_int64, c_int, _P, _P, c_int]),
    "t4r_train_row_codes_bwd": (c_int, [_P, _P, c_int64, c_int, _P, _P, _P, _P, c_int]),
    "t4r_train_gather_rows": (c_int, [_P, _P, c_int64, c_int, _P, _P, c_int]),
    "t4r_train_scatter_rows": (c_int, [_P, _P, c_int64, c_int, c_int64, _P, _P, c_int]),
    "t4r_train_softmax_ce_bwd": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, c_float, c_float, c_int64, _P, c_int]),
    "t4r_train_sampled_ce_bwd": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, c_float, c_float, _P, c_int]),
    "t4r_train_index_add_rows": (c_int, [_P, _P, _P, c_int64, c_int64, c_int, c_int, c_int64, _P, c_int]),
    "t4r_train_soft_emb_fwd": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, _P, _P, _P, c_int]),
    "t4r_train_soft_emb_bwd": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, _P, _P, _P, c_int]),
    "t4r_train_binary": (c_int, [c_int, _P, _P, _P, c_int64, _P, c_int]),
    "t4r_train_adamw": (c_int, [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_float, c_int, _P, c_int]),
    "t4r_train_col_sum": (c_int, [_P, c_int64, c_int64, _P, _P, c_int]),
    "t4r_train_layer_norm_fwd": (c_int, [_P, _P, _P, c_int64, c_int, c_float, _P, _P, c_int]),
    "t4r_train_layer_norm_bwd": (c_int, [_P, _P, c_int64, c_int, c_float, _P, _P, _P, _P, _P, _P, _P, c_int]),
    "t4r_train_attn_bwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, c_int]),
    "t4r_train_dropout": (c_int, [_P, _P, c_int64, c_float, C.c_uint64, C.c_uint32, _P, c_int]),
    "t4r_train_attn_drop_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, c_float, C.c_uint64, C.c_uint32, _P, _P,
                                        c_int]),
    "t4r_train_attn_drop_bwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, c_float,
                                        C.c_uint64, C.c_uint32, _P, c_int]),
    "t4r_train_xlnet_attn_plm_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "t4r_train_xlnet_attn_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "t4r_train_causal_attn_fwd": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P]),
    "t4r_train_rel_pos_proj": (c_int, [C.POINTER(c_void_p), c_int, c_int, c_int, _P, _P]),
    "t4r_recall_from_ranks": (c_int, [_P, _P, c_int, C.POINTER(C.c_int32), c_int, _P, _P]),
    "t4r_topk": (c_int, [_P, c_int64, c_int64, c_int64, c_int, _P, _P, _P]),
    "t4r_combine_shard_lse": (c_int, [_P, c_int, c_int, _P, _P, _P, _P]),
    "t4r_peer_export": (c_int, [_P, _P, C.POINTER(c_int64)]),
    "t4r_peer_open": (c_int, [_P, c_int64, C.POINTER(c_void_p)]),
    "t4r_peer_close": (c_int, [_P, c_int64]),
    "t4r_peer_gather_rows": (c_int, [C.POINTER(PeerPtrs), c_int64, c_int64, c_int, _P, _P, c_int, c_int64, _P, _P, _P, _P]),
    "t4r_peer_pull_rows": (c_int, [C.POINTER(PeerPtrs), C.POINTER(PeerPtrs), _P, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    "t4r_peer_combine_lse": (c_int, [C.POINTER(PeerPtrs), c_int64, _P, c_int, _P, _P, _P, _P]),
}

_lib = None
_lock = threading.Lock()


def build(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into ``libt4r_b200.so`` (in-tree)."""
    cmd = ["make", "-C", CSRC, "-j4"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise T4RError("building libt4r_b200.so failed (see output above)")
    return LIB_PATH


def load():
    """Load the shared library (once) and attach the C signatures."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise T4RError(
                f"{LIB_PATH} not found: run transformers4rec_b200.build() (or __graft_entry__.build()). "
                "There is no CPU/eager fallback for the t4r_b200 hot path."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI and the header drift apart
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().t4r_last_error()
        raise T4RError(f"{what}: {msg.decode() if msg else 'unknown error'} (code {rc})")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return t.data_ptr()
