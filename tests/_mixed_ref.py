"""Reference (plain torch, CPU) of the operand format of the 2-unit product (``nprod=2``,
``transformers4rec_b200/csrc/t4r_mixed_pack.cuh``) and of the product itself.  Test infrastructure."""
import torch

TARGET_EXP, HI8_SCALE, LO8_SCALE = 13, 1.0 / 64.0, 64.0


def row_scales(x: torch.Tensor):
    """scale[row] = 2^(13 - floor(log2 max|x_row|)) clamped to 2^+-126; 1 for an all-zero row."""
    m = x.abs().amax(dim=1)
    _, ex = torch.frexp(m)  # m = f * 2^ex, f in [0.5, 1)
    sh = (TARGET_EXP - (ex - 1)).clamp(-126, 126).to(torch.float32)
    scale = torch.where(m == 0, torch.ones_like(m), torch.exp2(sh))
    return scale, 1.0 / scale


def pack(x: torch.Tensor):
    """fp32 [rows, K] -> dict(h16 [rows, Kp] fp16, hi8 / lo8 [rows, Kp] e4m3, inv_scale [rows])."""
    rows, K = x.shape
    Kp = (K + 63) // 64 * 64
    scale, inv = row_scales(x)
    xs = torch.zeros((rows, Kp), dtype=torch.float32)
    xs[:, :K] = x * scale[:, None]
    h = xs.to(torch.float16)
    f = h.float()
    hi8 = (f * HI8_SCALE).to(torch.float8_e4m3fn)
    lo8 = ((xs - f) * LO8_SCALE).to(torch.float8_e4m3fn)
    return {"h16": h, "hi8": hi8, "lo8": lo8, "inv_scale": inv}


def unpack_planes(planes: torch.Tensor):
    """int16 words [2, rows, Kp] as written by t4r_split_planes_mixed -> (h16, hi8, lo8) tensors."""
    _, rows, Kp = planes.shape
    h16 = planes[0].view(torch.float16)
    b = planes[1].contiguous().view(torch.uint8).reshape(rows, Kp // 64, 2, 64)
    hi8 = b[:, :, 0, :].reshape(rows, Kp).contiguous().view(torch.float8_e4m3fn)
    lo8 = b[:, :, 1, :].reshape(rows, Kp).contiguous().view(torch.float8_e4m3fn)
    return h16, hi8, lo8


def product(pa, pb, terms=("main", "cross1", "cross2")):
    """What the tensor core accumulates (exact products, here summed in fp64), scaled back: [rows_a, rows_b].
    ``terms`` selects the partial products (the kernel's T4R_GEMM_DEBUG bring-up switches 64 / 128 / 256 drop
    main / cross1 / cross2)."""
    d = torch.float64
    acc = torch.zeros((pa["h16"].shape[0], pb["h16"].shape[0]), dtype=d)
    if "main" in terms:
        acc += pa["h16"].to(d) @ pb["h16"].to(d).t()
    if "cross1" in terms:
        acc += pa["lo8"].to(d) @ pb["hi8"].to(d).t()
    if "cross2" in terms:
        acc += pa["hi8"].to(d) @ pb["lo8"].to(d).t()
    return acc * pa["inv_scale"].to(d)[:, None] * pb["inv_scale"].to(d)[None, :]
