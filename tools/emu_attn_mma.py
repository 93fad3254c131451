#!/usr/bin/env python
"""Lane-level emulation of the index algebra of the tensor-path attention kernels (csrc/t4r_attn_mma.cu).

No GPU needed: 32 lanes, ldmatrix (x4 / x2 / x2.trans) and mma.sync.m16n8k16 are emulated with their PTX fragment
layouts, shared memory is a flat array, and the kernels' address / fragment / shuffle expressions are transcribed
LITERALLY from the CUDA source.  Values are kept in float64 and only the hi plane is populated (this checks indices,
not the split-bf16 arithmetic).  The one-warp transcription (the kernel that is proven on hardware) validates the
emulator; the two-warp transcription (attn_mma64_kernel, written without a GPU) is then checked the same way against
the plain formula

    s[i, j] = ((q_i + r_w_bias) . k_j + (q_i + r_r_bias) . R[j + L - i]) / sqrt(dh)      (XLNet, HF:xlnet:95-140)
    s[i, j] = q_i . k_j / sqrt(dh),  j <= i                                               (GPT-2)
    out = softmax_j(s) V
"""
import itertools
import math

import numpy as np


class Smem:
    def __init__(self, n):
        self.a = np.zeros(n, dtype=np.float64)


def ldsm(sm, addrs, n_mat, trans=False):
    """addrs[lane] = element offset of an 8-element row; lanes 8 m .. 8 m + 7 supply the rows of matrix m.
    Returns regs[lane][m] = (x0, x1): matrix m element pair of lane (row g, cols 2t, 2t+1), or transposed."""
    mats = []
    for m in range(n_mat):
        rows = [sm.a[addrs[8 * m + r]: addrs[8 * m + r] + 8].copy() for r in range(8)]
        mats.append(np.stack(rows))             # [8 rows, 8 cols]
    out = np.zeros((32, n_mat, 2))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for m in range(n_mat):
            if trans:
                out[lane, m] = (mats[m][2 * t, g], mats[m][2 * t + 1, g])
            else:
                out[lane, m] = (mats[m][g, 2 * t], mats[m][g, 2 * t + 1])
    return out


def mma(c, a, b):
    """c[lane][4] += A(16x16) B(16x8): a[lane][4][2], b[lane][2][2] in the PTX m16n8k16 fragment layouts."""
    A = np.zeros((16, 16)); Bm = np.zeros((16, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        A[g, 2 * t: 2 * t + 2] = a[lane, 0]; A[g + 8, 2 * t: 2 * t + 2] = a[lane, 1]
        A[g, 2 * t + 8: 2 * t + 10] = a[lane, 2]; A[g + 8, 2 * t + 8: 2 * t + 10] = a[lane, 3]
        Bm[2 * t: 2 * t + 2, g] = b[lane, 0]; Bm[2 * t + 8: 2 * t + 10, g] = b[lane, 1]
    C = A @ Bm
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        c[lane] += (C[g, 2 * t], C[g, 2 * t + 1], C[g + 8, 2 * t], C[g + 8, 2 * t + 1])


def reference(q, k, v, R, rw, rr, rel, mask=None, stream=0):
    """mask / stream: XLNet two-stream attention (HF `score - 1e30 * mask`; the content stream h ignores the diagonal)."""
    L, dh = q.shape
    s = np.full((L, L), -np.inf)
    for i in range(L):
        for j in range(L):
            if rel:
                s[i, j] = ((q[i] + rw) @ k[j] + (q[i] + rr) @ R[j + L - i]) / math.sqrt(dh)
                if mask is not None and mask[i, j] and not (stream == 0 and i == j):
                    s[i, j] = -1e30
            elif j <= i:
                s[i, j] = q[i] @ k[j] / math.sqrt(dh)
    p = np.exp(s - s.max(1, keepdims=True))
    p /= p.sum(1, keepdims=True)
    return p @ v


def run(L, DH, rel, q, k, v, R, rw, rr, two_warp, mask=None, stream=0):
    """Transcription of attn_mma_kernel (two_warp=False) / attn_mma64_kernel (two_warp=True), hi plane only."""
    LDS, KT, NTC = DH + 8, DH // 16, DH // 8
    NT = 8 if two_warp else 4
    nwarps = 2 if two_warp else 1
    QR = L + 2 if rel else L
    RR = 2 * L
    S2LD = ((2 * L + 7) // 8) * 8 + 1
    # shared memory: zero row | R | Q | K | V (hi plane only: plane 1 would sit QR / L / RR rows further) | S2 | WB
    z0 = 0
    Rs = z0 + LDS
    Qs = Rs + 2 * RR * LDS
    Ks = Qs + 2 * QR * LDS
    Vs = Ks + 2 * L * LDS
    sm = Smem(Vs + 2 * L * LDS)
    S2 = np.zeros((QR, S2LD)); WB = np.zeros(64)
    for m in range(RR):
        sm.a[Rs + m * LDS: Rs + m * LDS + DH] = R[m] if rel else 0
    for i in range(L):
        sm.a[Qs + i * LDS: Qs + i * LDS + DH] = q[i]
        sm.a[Ks + i * LDS: Ks + i * LDS + DH] = k[i]
        sm.a[Vs + i * LDS: Vs + i * LDS + DH] = v[i]
    if rel:
        sm.a[Qs + L * LDS: Qs + L * LDS + DH] = rw
        sm.a[Qs + (L + 1) * LDS: Qs + (L + 1) * LDS + DH] = rr
    scale = 1.0 / math.sqrt(DH)
    out = np.zeros((L, DH))
    lanes = range(32)
    state = []
    # ---------------- phase 1 per warp: fragments, S1, S2 (+ WB)
    for wp in range(nwarps):
        mt0 = 2 * wp
        aq = np.zeros((2, KT, 32, 4, 2))
        for ml, kt in itertools.product(range(2), range(KT)):
            addrs = []
            for lane in lanes:
                row = (mt0 + ml) * 16 + (lane & 15)
                col = kt * 16 + (lane >> 4) * 8
                addrs.append(Qs + row * LDS + col if row < QR else z0)
            aq[ml, kt] = ldsm(sm, addrs, 4)
        s1 = np.zeros((2, NT, 32, 4))
        for nt, kt in itertools.product(range(NT), range(KT)):
            addrs = []
            for lane in lanes:
                row = nt * 8 + (lane & 7)
                col = kt * 16 + ((lane >> 3) & 1) * 8
                addrs.append(Ks + row * LDS + col if row < L else z0)
            bh = ldsm(sm, addrs, 2)
            for ml in range(2):
                mma(s1[ml, nt], aq[ml, kt], bh)
        if rel:
            for nt2 in range((2 * L + 7) // 8):
                acc = np.zeros((2, 32, 4))
                for kt in range(KT):
                    addrs = []
                    for lane in lanes:
                        row = nt2 * 8 + (lane & 7)
                        col = kt * 16 + ((lane >> 3) & 1) * 8
                        addrs.append(Rs + row * LDS + col if row < RR else z0)
                    bh = ldsm(sm, addrs, 2)
                    for ml in range(2):
                        mma(acc[ml], aq[ml, kt], bh)
                for ml, lane in itertools.product(range(2), lanes):
                    g, t = lane >> 2, lane & 3
                    col = nt2 * 8 + 2 * t
                    r0 = (mt0 + ml) * 16 + g
                    r1 = r0 + 8
                    if r0 < QR:
                        S2[r0, col] = acc[ml, lane, 0]; S2[r0, col + 1] = acc[ml, lane, 1]
                    if r1 < QR:
                        S2[r1, col] = acc[ml, lane, 2]; S2[r1, col + 1] = acc[ml, lane, 3]
            mtL, rL = L >> 4, L & 15
            if two_warp:
                if (mtL >> 1) == wp:
                    mlL = mtL & 1
                    for lane in lanes:
                        g, t = lane >> 2, lane & 3
                        if g == (rL & 7):
                            for nt in range(NT):
                                v0 = s1[mlL, nt, lane, 0] if rL < 8 else s1[mlL, nt, lane, 2]
                                v1 = s1[mlL, nt, lane, 1] if rL < 8 else s1[mlL, nt, lane, 3]
                                WB[nt * 8 + 2 * t] = v0
                                WB[nt * 8 + 2 * t + 1] = v1
        state.append((mt0, s1))
    # ---------------- phase 2 per warp (after the pair barrier): scores, softmax, P V, store
    for wp in range(nwarps):
        mt0, s1 = state[wp]
        mtL, rL = L >> 4, L & 15
        srcL = ((rL & 7) << 2)
        rmax = np.full((2, 32, 2), -np.inf)
        for nt in range(NT):
            wb = np.zeros((32, 2))
            if rel and not two_warp:
                for lane in lanes:
                    t = lane & 3
                    src = srcL | t
                    wb[lane, 0] = s1[mtL, nt, src, 0] if rL < 8 else s1[mtL, nt, src, 2]
                    wb[lane, 1] = s1[mtL, nt, src, 1] if rL < 8 else s1[mtL, nt, src, 3]
            for ml, lane, e in itertools.product(range(2), lanes, range(4)):
                g, t = lane >> 2, lane & 3
                i = (mt0 + ml) * 16 + g + ((e >> 1) << 3)
                j = nt * 8 + 2 * t + (e & 1)
                val = s1[ml, nt, lane, e]
                if rel:
                    ii = i if i < L else 0
                    jj = j if j < L else 0
                    m = jj + L - ii
                    val += (WB[j] if two_warp else wb[lane, e & 1]) + S2[ii, m] + S2[L + 1, m]
                val *= scale
                if mask is not None:   # MASKED instantiation
                    if i < L and j < L and not (stream == 0 and i == j) and mask[i, j]:
                        val = -1e30
                if j >= L or (not rel and j > i):
                    val = -np.inf
                s1[ml, nt, lane, e] = val
                rmax[ml, lane, e >> 1] = max(rmax[ml, lane, e >> 1], val)
        # quad reductions (shfl_xor 1, 2)
        for ml, hf in itertools.product(range(2), range(2)):
            for lane0 in range(0, 32, 4):
                rmax[ml, lane0:lane0 + 4, hf] = rmax[ml, lane0:lane0 + 4, hf].max()
        rsum = np.zeros((2, 32, 2))
        for nt, ml, lane, e in itertools.product(range(NT), range(2), lanes, range(4)):
            val = s1[ml, nt, lane, e]
            p = 0.0 if val == -np.inf else math.exp(val - rmax[ml, lane, e >> 1])
            s1[ml, nt, lane, e] = p
            rsum[ml, lane, e >> 1] += p
        for ml, hf in itertools.product(range(2), range(2)):
            for lane0 in range(0, 32, 4):
                tot = rsum[ml, lane0:lane0 + 4, hf].sum()
                rsum[ml, lane0:lane0 + 4, hf] = 1.0 / tot if tot > 0 else 0.0
        o = np.zeros((2, NTC, 32, 4))
        for kt in range(NT // 2):
            ph = np.zeros((2, 32, 4, 2))
            for ml, lane in itertools.product(range(2), lanes):
                ph[ml, lane, 0] = (s1[ml, 2 * kt, lane, 0], s1[ml, 2 * kt, lane, 1])
                ph[ml, lane, 1] = (s1[ml, 2 * kt, lane, 2], s1[ml, 2 * kt, lane, 3])
                ph[ml, lane, 2] = (s1[ml, 2 * kt + 1, lane, 0], s1[ml, 2 * kt + 1, lane, 1])
                ph[ml, lane, 3] = (s1[ml, 2 * kt + 1, lane, 2], s1[ml, 2 * kt + 1, lane, 3])
            for nc in range(NTC):
                addrs = []
                for lane in lanes:
                    row = kt * 16 + (lane & 15)
                    addrs.append(Vs + row * LDS + nc * 8 if row < L else z0)
                bh = ldsm(sm, addrs, 2, trans=True)
                for ml in range(2):
                    mma(o[ml, nc], ph[ml], bh)
        for ml, hf, lane in itertools.product(range(2), range(2), lanes):
            g, t = lane >> 2, lane & 3
            i = (mt0 + ml) * 16 + g + hf * 8
            if i < L:
                inv = rsum[ml, lane, hf]
                for nc in range(NTC):
                    out[i, nc * 8 + 2 * t] = o[ml, nc, lane, 2 * hf] * inv
                    out[i, nc * 8 + 2 * t + 1] = o[ml, nc, lane, 2 * hf + 1] * inv
    return out


def check(L, DH, rel, two_warp, seed=0, masked=False):
    """masked: the two-stream (PLM) instantiation -- random permutation-style mask incl. a fully masked row; both the
    content stream (queries = the keys' own rows, diagonal exempt) and the query stream (separate query rows)."""
    rng = np.random.default_rng(seed)
    q, k, v = (rng.standard_normal((L, DH)) for _ in range(3))
    R = rng.standard_normal((2 * L, DH))
    rw, rr = rng.standard_normal(DH), rng.standard_normal(DH)
    if not masked:
        got = run(L, DH, rel, q, k, v, R, rw, rr, two_warp)
        ref = reference(q, k, v, R, rw, rr, rel)
        return float(np.abs(got - ref).max())
    mask = rng.random((L, L)) < 0.4
    mask[min(3, L - 1), :] = True                      # a query that sees nothing: uniform attention (HF's -1e30 rule)
    qg = rng.standard_normal((L, DH))
    err = 0.0
    for stream, qs in ((0, q), (1, qg)):
        got = run(L, DH, rel, qs, k, v, R, rw, rr, two_warp, mask=mask, stream=stream)
        ref = reference(qs, k, v, R, rw, rr, rel, mask=mask, stream=stream)
        err = max(err, float(np.abs(got - ref).max()))
    return err


if __name__ == "__main__":
    # the proven kernel validates the emulator ...
    for L, DH, rel in [(20, 32, True), (30, 16, True), (7, 64, True), (32, 32, False), (17, 16, False)]:
        e = check(L, DH, rel, two_warp=False)
        print(f"one-warp kernel  L={L:2d} dh={DH:2d} rel={int(rel)}: max err {e:.2e}")
        assert e < 1e-9
    # ... and the same emulator checks the two-warp generalisation
    for L, DH, rel in [(50, 32, True), (33, 16, True), (62, 32, True), (47, 64, True), (40, 16, True), (31, 32, True),
                       (64, 32, False), (33, 16, False), (50, 64, False)]:
        e = check(L, DH, rel, two_warp=True)
        print(f"two-warp kernel  L={L:2d} dh={DH:2d} rel={int(rel)}: max err {e:.2e}")
        assert e < 1e-9
    for L, DH, two in [(20, 32, False), (30, 16, False), (50, 32, True), (62, 16, True)]:
        e = check(L, DH, True, two_warp=two, masked=True)
        print(f"two-stream (PLM) {'two' if two else 'one'}-warp  L={L:2d} dh={DH:2d}: max err {e:.2e}")
        assert e < 1e-9
    print("attention index algebra: OK")
