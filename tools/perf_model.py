#!/usr/bin/env python
"""First-principles time model of one forward step (fwd + loss) per BASELINE config, next to what was measured.

For every kernel of the step: algorithmic flops / bytes (SURVEY §8d formulas), tensor passes actually issued
(3 per MAC with the split-bf16 product, 2 with the fp16 + e4m3 product), HBM bytes, L2->SM operand bytes, and the time
each resource would need at the MEASURED peaks of this pool's B200s (MEASURED_PEAKS.json: sustained bf16 rate, HBM
copy bandwidth; L2->SM stream taken as 8 TB/s, the rate the head GEMM was observed to sustain).  The bound is the
largest of those; "measured" quotes profiles/ where a number exists (config 2).  It is a planning aid for round 2:
where the head stops being tensor-bound, what the resident-A kernel buys, how far the encoder GEMMs are from their
store floor.

    python tools/perf_model.py            # all configs, shipped arithmetic and the gated variants
"""
import json
import math
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
    PK = json.load(f)
TENSOR = PK["bf16_tflops_sustained"] * 1e12      # dense bf16 flop/s, sustained under the 1 kW cap
HBM = PK["hbm_gbs"] * 1e9
L2SM = 8.0e12                                     # observed L2->SM operand stream of the head GEMM

CONFIGS = {
    "config2": dict(B=2048, L=20, d=256, NL=4, V=1_000_001, De=256, C=256, T_per_session=2.5, feats=1, arch="xlnet"),
    "config3": dict(B=4096, L=20, d=256, NL=4, V=1_000_001, De=64, C=448, T_per_session=10.0, feats=7, arch="gpt2"),
    "config4/rank": dict(B=2048, L=20, d=256, NL=4, V=1_250_001, De=256, C=256, T_per_session=2.5 * 8, feats=1, arch="xlnet"),
    "config5/rank": dict(B=2048, L=50, d=256, NL=4, V=50_001, De=256, C=256, T_per_session=4.2, feats=1, arch="xlnet"),
}
MEASURED_US = {"config2": {"head": 5900.0, "ffn": 148.0, "qkv": 52.0, "oproj": 60.0, "attention": 84.0, "gather": 18.5}}


def gemm(M, N, K, passes, out_bytes_per_elem, a_resident=False, pair=True):
    """(tensor s, hbm s, l2 s) of C[M,N] = A[M,K] B[N,K]^T with 128 x 256 tiles (256 x 256 per CTA pair)."""
    flops_issued = 2.0 * M * N * K * passes
    t_tensor = flops_issued / TENSOR
    hbm = (M * K + N * K) * 4 + M * N * out_bytes_per_elem
    tiles = math.ceil(M / 256) * math.ceil(N / 256)
    per_tile = (256 * K * 4 * (0 if a_resident else 1) + 256 * K * 4)          # operand bytes per pair tile
    if a_resident:
        per_tile += 256 * K * 4 / 16                                            # A reloaded once per 16 column tiles
    return t_tensor, hbm / HBM, tiles * per_tile / L2SM


def row(name, parts, measured=None):
    t_tensor, t_hbm, t_l2 = parts
    bound = max(t_tensor, t_hbm, t_l2)
    which = ["tensor", "hbm", "l2->sm"][[t_tensor, t_hbm, t_l2].index(bound)]
    m = f"{measured:9.1f}" if measured else "        -"
    print(f"  {name:34s} tensor {t_tensor * 1e6:8.1f}  hbm {t_hbm * 1e6:8.1f}  l2 {t_l2 * 1e6:8.1f}  -> {bound * 1e6:8.1f} us ({which:6s})  measured {m}")
    return bound


def main():
    print(f"peaks: tensor {TENSOR / 1e12:.0f} TF/s (sustained bf16), HBM {HBM / 1e9:.0f} GB/s, L2->SM {L2SM / 1e12:.0f} TB/s (observed)\n")
    for name, c in CONFIGS.items():
        M = c["B"] * c["L"]
        T = int(c["B"] * c["T_per_session"])
        meas = MEASURED_US.get(name, {})
        print(f"== {name}: M = {M} rows, T = {T} label rows, V = {c['V']}, De = {c['De']}")
        total = {}
        for mode, passes, resident in (("shipped (3 passes)", 3, False), ("2 passes", 2, False), ("2 passes + resident A", 2, True),
                                       ("3 passes + resident A", 3, True)):
            h = gemm(T, c["V"], c["De"], passes, 0, a_resident=resident)
            total[mode] = row(f"head [{mode}]", h, meas.get("head") if mode.startswith("shipped") else None)
        enc = 0.0
        d = c["d"]
        enc += row("gather (K1)", (0.0, (M * (8 * c["feats"] + 4 * c["C"] + 4 * c["C"])) / HBM, 0.0), meas.get("gather"))
        enc += row("projection", gemm(M, d, c["C"], 3, 8), None)
        per_layer = 0.0
        per_layer += row("  qkv GEMM (planes out)", gemm(M, 3 * d, d, 3, 4), meas.get("qkv"))
        att_flops = c["B"] * 8 * (4 * c["L"] ** 2 * (d // 8) * (2 if c["arch"] == "xlnet" else 1)) * 3
        per_layer += row("  attention (mma.sync, latency)", (att_flops / (TENSOR / 8), M * 4 * d * 4 / HBM, 0.0), meas.get("attention"))
        per_layer += row("  o-proj + LN", gemm(M, d, d, 3, 8), meas.get("oproj"))
        f1 = gemm(M, 4 * d, d, 3, 0)
        f2 = gemm(M, d, 4 * d, 3, 8)
        per_layer += row("  fused FFN", (f1[0] + f2[0], (M * d * 4 * 2 + M * d * 8) / HBM, f1[2] + f2[2]), meas.get("ffn"))
        enc += per_layer * c["NL"]
        print(f"  encoder + input block (model): {enc * 1e6:8.1f} us")
        for mode, t in total.items():
            step = t + enc
            print(f"  step [{mode:24s}] {step * 1e3:7.2f} ms -> {c['B'] / step / 1e3:8.1f} k sessions/s (model: every kernel at its bound)")
        print()


if __name__ == "__main__":
    main()
