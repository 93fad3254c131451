#!/usr/bin/env python
"""Per-kernel timings at BASELINE config-2 shapes (B=2048, L=20, d=256, H=8): each op is
run in isolation with CUDA events; also the target of the per-kernel ncu captures."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from transformers4rec_b200 import _lib, ops  # noqa: E402


def timeit(name, fn, iters=20, warm=3, flops=None, bytes_=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    extra = ""
    if flops:
        extra += f"  {flops / us / 1e6:8.1f} TFLOP/s(alg)"
    if bytes_:
        extra += f"  {bytes_ / us / 1e3:8.1f} GB/s"
    print(f"{name:34s} {us:9.1f} us{extra}", flush=True)
    if os.environ.get("T4R_GEMM_DEBUG", "0") not in ("0", "1"):
        import ctypes
        lib = _lib.load()
        arr = (ctypes.c_ulonglong * 8)()
        if "ffn" in name:
            a16 = (ctypes.c_ulonglong * 16)()
            lib.t4r_debug_ffn_cycles(a16, 1)
            n, t = max(1, a16[7]), max(1, a16[13])
            print("    epilogue warp per chunk: wait_S %d  tmem_ld %d  gelu+split %d  wait_G_free %d  tmem_st %d | per tile: wait_Y %d  final epilogue %d  (%d chunks)"
                  % (a16[0] // n, a16[1] // n, a16[2] // n, a16[3] // n, a16[4] // n, a16[5] // t, a16[6] // t, a16[7]), flush=True)
            print("    mma thread per chunk: wait_S_free %d  gemm1 %d  wait_G %d  wait_Y_free %d  gemm2 %d  (%d tiles)"
                  % (a16[8] // n, a16[9] // n, a16[10] // n, a16[11] // n, a16[12] // n, a16[13]), flush=True)
            lib.t4r_debug_gemm_cycles(arr, 1)
            print("    final LN epilogue per tile: tmem_ld %d  bias+residual %d  stats+exchange %d  normalise+stores %d"
                  % (arr[2] // t, arr[3] // t, arr[4] // t, arr[5] // t), flush=True)
            return
        lib.t4r_debug_gemm_cycles(arr, 1)
        n = max(1, arr[6])
        print("    cta0/warp2 per tile: wait_tfull %d  epilogue %d  | per chunk-sum: tmem_ld %d math %d f32store %d planes %d (cycles; %d tiles)"
              % (arr[0] // n, arr[1] // n, arr[2] // n, arr[3] // n, arr[4] // n, arr[5] // n, arr[6]), flush=True)


def main():
    which = sys.argv[1:] or ["all"]
    B, L, d, H = 2048, 20, 256, 8
    M = B * L
    dev = "cuda"
    torch.manual_seed(0)
    x = torch.randn(M, d, device=dev)
    xp = ops.split_planes(x)
    w_qkv = ops.split_planes(torch.randn(3 * d, d, device=dev) * 0.05)
    w_o = ops.split_planes(torch.randn(d, d, device=dev) * 0.05)
    w_1 = ops.split_planes(torch.randn(4 * d, d, device=dev) * 0.05)
    w_2 = ops.split_planes(torch.randn(d, 4 * d, device=dev) * 0.05)
    b1, b2 = torch.randn(4 * d, device=dev), torch.randn(d, device=dev)
    g, bt = torch.ones(d, device=dev), torch.zeros(d, device=dev)
    ffp = ops.split_planes(torch.randn(M, 4 * d, device=dev))
    if ("head2" in which or "headres" in which) and "head" not in which:
        which.append("head")
    on = lambda k: "all" in which or k in which
    if on("qkv"):
        timeit("qkv gemm N=768 K=256 (f32 out)", lambda: ops.linear(xp, w_qkv, d, want_planes=False),
               flops=2 * M * 3 * d * d)
    if on("qkvp"):
        timeit("qkv gemm N=768 K=256 (planes out)", lambda: ops.linear(xp, w_qkv, d, want_f32=False, want_planes=True),
               flops=2 * M * 3 * d * d)
    if on("ffn1"):
        timeit("ffn1 gemm N=1024 K=256 gelu", lambda: ops.linear(xp, w_1, d, bias=b1, act=_lib.ACT_GELU, want_f32=False),
               flops=2 * M * 4 * d * d)
    if on("oproj"):
        timeit("oproj gemm N=256 K=256 res+LN", lambda: ops.linear(xp, w_o, d, residual=x, ln=(g, bt), ln_eps=0.03),
               flops=2 * M * d * d)
    if on("ffn2"):
        timeit("ffn2 gemm N=256 K=1024 res+LN", lambda: ops.linear(ffp, w_2, 4 * d, bias=b2, residual=x, ln=(g, bt), ln_eps=0.03),
               flops=2 * M * 4 * d * d)
    if on("ffn"):
        timeit("fused ffn d=256 hidden=1024 res+LN", lambda: ops.ffn(xp, w_1, b1, w_2, b2, (g, bt), 0.03, want_f32=False, want_planes=True),
               flops=2 * 2 * M * 4 * d * d)
    if on("proj"):
        code = torch.zeros(M, dtype=torch.uint8, device=dev)
        timeit("proj gemm N=256 K=256 relu+mask", lambda: ops.linear(xp, w_o, d, bias=b2, act=_lib.ACT_RELU, row_code=code, mask_vec=b2),
               flops=2 * M * d * d)
    if on("attn"):
        import transformers4rec_b200.torch as tr
        cfg = tr.XLNetConfig.build(d_model=d, n_head=H, n_layer=1, total_seq_length=L)
        enc = tr.TransformerBlock(cfg).cuda()
        xin = x.view(B, L, d)
        timeit("xlnet layer (qkv+attn+o+ffn)", lambda: enc(xin))
    if "attn50" in which:
        # BASELINE configs[4] shape (L = 50): FFMA attention vs the two-warp tensor-path kernel (T4R_ATTN_MMA64=1)
        import transformers4rec_b200.torch as tr
        L50 = 50
        cfg = tr.XLNetConfig.build(d_model=d, n_head=H, n_layer=1, total_seq_length=L50)
        enc = tr.TransformerBlock(cfg).cuda()
        x50 = torch.randn(B, L50, d, device=dev)
        for flag in ("0", "1"):
            os.environ["T4R_ATTN_MMA64"] = flag
            timeit(f"xlnet layer L=50 (T4R_ATTN_MMA64={flag})", lambda: enc(x50))
        os.environ["T4R_ATTN_MMA64"] = "0"
    if on("embed"):
        # HBM-honest: 8 different id sets (8 x 42 MB of random 1 KB rows > L2) replayed from a CUDA graph
        # so neither L2 residency of the rows nor host launch overhead flatters / hides the kernel
        V = 1_000_001
        table = torch.randn(V, d, device=dev)
        id_sets = [torch.randint(1, V, (M,), device=dev) for _ in range(8)]
        for ids in id_sets:
            ops.embed_concat([(table, ids, 0)], [], M, d, False, True)
        torch.cuda.synchronize()
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph):
            outs = [ops.embed_concat([(table, ids, 0)], [], M, d, False, True) for ids in id_sets]
        timeit("embed_concat 1M x 256 (8 id sets, graph)", gph.replay, iters=10,
               bytes_=8 * M * (8 + 4 * d + 4 * d))
    if on("head4"):
        # config-4-like shard: 10 240 label rows (two ranks' worth) against 2.5 M local table rows
        V, T = 2_500_001, 10240
        W = torch.randn(V, d, device=dev) * 0.05
        wp = ops.split_planes(W)
        xt = torch.randn(T, d, device=dev)
        xtp = ops.split_planes(xt)
        y = torch.randint(1, V, (T,), device=dev)
        timeit("head 10240 x 2.5M x 256", lambda: ops.head_softmax_ce(xtp, xt, y, wp, W), iters=3, flops=2 * T * V * d)
        del W, wp
    if on("head64"):
        # config-3-like head: De = 64 puts a single 64-wide K block under each 128x256 tile, so the online-LSE
        # epilogue, not the MMA main loop, is what bounds it
        V, T, De = 1_000_001, 8192, 64
        W = torch.randn(V, De, device=dev) * 0.05
        wp = ops.split_planes(W)
        xt = torch.randn(T, De, device=dev)
        xtp = ops.split_planes(xt)
        y = torch.randint(1, V, (T,), device=dev)
        timeit("head 8192 x 1M x 64", lambda: ops.head_softmax_ce(xtp, xt, y, wp, W), iters=5, flops=2 * T * V * De)
        del W, wp
    if on("head"):
        V, T = 1_000_001, 5120
        W = torch.randn(V, d, device=dev) * 0.05
        wp = ops.split_planes(W)
        xt = torch.randn(T, d, device=dev)
        xtp = ops.split_planes(xt)
        y = torch.randint(1, V, (T,), device=dev)
        for nprod in (3, 1):  # 1 = plain bf16: same operand traffic, a third of the MMAs -> separates tensor- from L2-bound
            timeit(f"head 5120 x 1M x 256 nprod={nprod}", lambda: ops.head_softmax_ce(xtp, xt, y, wp, W, nprod=nprod),
                   iters=5, flops=2 * T * V * d)
        if "headres" in which:  # A-resident head kernel (T4R_HEAD_RESIDENT=1), all three arithmetics
            os.environ["T4R_HEAD_RESIDENT"] = "1"
            for nprod in (3, 1):
                timeit(f"head 5120 x 1M x 256 nprod={nprod} resident-A", lambda: ops.head_softmax_ce(xtp, xt, y, wp, W, nprod=nprod),
                       iters=5, flops=2 * T * V * d)
            os.environ["T4R_HEAD_RESIDENT"] = "0"
        if "head2" in which:  # 2-unit product (fp16 + two e4m3 cross terms): same operand bytes, 2/3 of the tensor passes
            ref = ops.head_softmax_ce(xtp, xt, y, wp, W, nprod=3)
            del wp
            wm, wi = ops.split_planes_mixed(W)
            xm, xi = ops.split_planes_mixed(xt)
            got = ops.head_softmax_ce(xm, xt, y, wm, W, nprod=2, xt_inv_scale=xi, w_inv_scale=wi)
            print("    nprod=2 vs nprod=3: loss %.7f vs %.7f, max |row_lse diff| %.3e" % (
                got["loss"].item(), ref["loss"].item(), (got["row_lse"] - ref["row_lse"]).abs().max().item()), flush=True)
            timeit("head 5120 x 1M x 256 nprod=2", lambda: ops.head_softmax_ce(xm, xt, y, wm, W, nprod=2, xt_inv_scale=xi,
                                                                                w_inv_scale=wi), iters=5, flops=2 * T * V * d)
            os.environ["T4R_HEAD_RESIDENT"] = "1"
            timeit("head 5120 x 1M x 256 nprod=2 resident-A", lambda: ops.head_softmax_ce(
                xm, xt, y, wm, W, nprod=2, xt_inv_scale=xi, w_inv_scale=wi), iters=5, flops=2 * T * V * d)
            os.environ["T4R_HEAD_RESIDENT"] = "0"


if __name__ == "__main__":
    main()
