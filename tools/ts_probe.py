import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transformers4rec_b200 import _lib, ops
lib = _lib.load()
torch.manual_seed(0)
for N in (64, 128, 256):
    A = torch.randn(128, 64, device="cuda")
    B = torch.randn(N, 64, device="cuda") * 0.1
    bp = ops.split_planes(B)
    D = torch.zeros(128, N, device="cuda")
    _lib.check(lib.t4r_debug_ts_mma(A.data_ptr(), bp.data_ptr(), N, D.data_ptr(), torch.cuda.current_stream().cuda_stream), "ts")
    torch.cuda.synchronize()
    ref = A.bfloat16().float() @ bp[0].float().t()
    err = (D - ref).abs().max().item()
    print(f"N={N}: max err vs bf16 reference {err:.3e}  (|ref| max {ref.abs().max().item():.3f})", flush=True)
    if err > 1e-3:
        # diagnose: which rows/cols match?
        ok_rows = ((D - ref).abs().max(1).values < 1e-3).nonzero().flatten().tolist()
        print("   rows matching:", ok_rows[:40], "count", len(ok_rows))
