// t4r_train.cu -- element / row kernels of the training step (transformers4rec_b200/training.py).
//
// Every matmul of the backward is the tcgen05 GEMM of t4r_gemm.cu on transposed operands; what is left are the small
// pieces below.  Each is written as a __host__ __device__ function of ONE work item (element, row, column slab,
// or attention query row) that a trivial kernel calls once per thread -- and that the same entry point calls in a
// plain loop when `on_host` is set (HOST pointers, no CUDA call), so that the CPU suite can pin the arithmetic
// against torch without a GPU.  Work items own what they write; the few cross-item sums (column sums, the row
// scatter-add into embedding tables) use one atomicAdd per item / per duplicate row on the device (plain += on the
// host), never one per element.  These are first, correct versions, not tuned (one thread per row or per (session,
// head) leaves bandwidth on the table); none of this is on the inference path.
#include <math.h>

#include "t4r_common.cuh"
#include "t4r_internal.h"

namespace t4r {

#if defined(__CUDA_ARCH__)
#define T4R_ATOMIC_ADD(ptr, val) atomicAdd((ptr), (val))
#else
#define T4R_ATOMIC_ADD(ptr, val) (*(ptr) += (val))
#endif
#define T4R_HD __host__ __device__ inline

// ---------------------------------------------------------------------------------------------- element-wise
T4R_HD float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
T4R_HD float gelu_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}
// ---------------------------------------------------------------------------------------------- dropout
// Counter-based, so the backward regenerates the forward's mask instead of storing it: element `idx` of dropout site
// `site` under `seed` keeps its value (scaled by 1 / (1 - p)) iff word (idx & 3) of Philox4x32-10(key = seed,
// counter = (idx >> 2, site)) >= p * 2^32.  Same bits on the host and on the device.
T4R_HD void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4]) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c0;
    const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = static_cast<uint32_t>(p1);
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = static_cast<uint32_t>(p0);
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
T4R_HD float keep_scale(uint64_t seed, uint32_t site, uint64_t idx, float p) {
  if (p <= 0.f) return 1.f;
  uint32_t w[4];
  const uint64_t ctr = idx >> 2;
  philox4x32_10(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), static_cast<uint32_t>(ctr),
                static_cast<uint32_t>(ctr >> 32), site, 0x7434720u, w);
  const uint32_t bits = w[idx & 3];
  const float u = static_cast<float>(bits >> 8) * (1.0f / 16777216.0f);   // 24 uniform bits in [0, 1)
  return u >= p ? 1.0f / (1.0f - p) : 0.f;
}
T4R_HD void dropout_item(const float* x, float* y, float p, uint64_t seed, uint32_t site, int64_t i) {
  y[i] = x[i] * keep_scale(seed, site, static_cast<uint64_t>(i), p);
}

T4R_HD void transpose_item(const float* x, int64_t R, int64_t C, float* out, int64_t i) {
  const int64_t r = i / C, c = i % C;
  out[c * R + r] = x[i];
}
T4R_HD void act_fwd_item(int kind, const float* x, float* y, int64_t i) {
  const float v = x[i];
  y[i] = kind == T4R_ACT_GELU ? gelu_exact(v) : (kind == T4R_ACT_RELU ? fmaxf(v, 0.f) : v);
}
T4R_HD void act_bwd_item(int kind, const float* pre, const float* dy, float* dx, int64_t i) {
  const float v = pre[i];
  dx[i] = dy[i] * (kind == T4R_ACT_GELU ? gelu_grad(v) : (kind == T4R_ACT_RELU ? (v > 0.f ? 1.f : 0.f) : 1.f));
}
// y[b, l, :] = x[b, l, :] + wpe[l, :]
T4R_HD void add_pos_item(const float* x, const float* wpe, int L, int d, float* y, int64_t i) {
  const int64_t row = i / d;
  y[i] = x[i] + wpe[(row % L) * d + (i % d)];
}
// out[l, c] = sum_b x[b, l, c]
T4R_HD void sum_sessions_item(const float* x, int B, int L, int d, float* out, int64_t i) {
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += x[static_cast<int64_t>(b) * L * d + i];
  out[i] = s;
}
// apply_mask_to_inputs as row codes: 0 keep, 1 -> mask_vec, 2 -> 0
T4R_HD void row_codes_fwd_item(const float* y, const uint8_t* code, const float* mask_vec, int d, float* out, int64_t i) {
  const int c = code[i / d];
  out[i] = c == 0 ? y[i] : (c == 1 ? mask_vec[i % d] : 0.f);
}
// dy = dx on kept rows; tmp = dx on rows that took the mask embedding (its gradient is their column sum)
T4R_HD void row_codes_bwd_item(const float* dx, const uint8_t* code, int d, float* dy, float* tmp, int64_t i) {
  const int c = code[i / d];
  dy[i] = c == 0 ? dx[i] : 0.f;
  tmp[i] = c == 1 ? dx[i] : 0.f;
}
T4R_HD void gather_rows_item(const float* x, const int32_t* idx, int d, float* out, int64_t i) {
  out[i] = x[static_cast<int64_t>(idx[i / d]) * d + (i % d)];
}
T4R_HD void scatter_rows_item(const float* src, const int32_t* idx, int d, float* out, int64_t i) {
  out[static_cast<int64_t>(idx[i / d]) * d + (i % d)] = src[i];  // label rows are unique
}
// P = exp(z - lse) * scale, minus scale at the label's column (if it falls in this column chunk)
// label smoothing e (transformers4rec/torch/losses.py:4-20): the target distribution is (1 - e) onehot + e / V
T4R_HD void softmax_ce_bwd_item(float* z, const float* lse, const int64_t* labels, int64_t Vc, int64_t v0, float scale,
                                float smooth, float inv_V, int64_t i) {
  const int64_t t = i / Vc, j = i % Vc;
  float p = (expf(z[i] - lse[t]) - smooth * inv_V) * scale;
  if (labels[t] - v0 == j) p -= (1.0f - smooth) * scale;
  z[i] = p;
}
// sampled softmax (model/prediction_task.py:673-696): z holds x.w_s / tau for the S sampled negatives; the logit the
// forward used is z + bias_s / tau, except accidental hits (col_ids[s] == labels[t]) which were a constant -> no gradient.
// In place: z[t, s] <- exp(logit - lse[t]) * scale, 0 at hits.
T4R_HD void sampled_ce_bwd_item(float* z, const float* lse, const int64_t* labels, const float* col_bias,
                                const int64_t* col_ids, int64_t S, float inv_tau, float scale, int64_t i) {
  const int64_t t = i / S, s_ = i % S;
  if (col_ids[s_] == labels[t]) { z[i] = 0.f; return; }
  z[i] = expf(z[i] + col_bias[s_] * inv_tau - lse[t]) * scale;
}
// dst[idx[r], 0:width] += src[r, col:col+width]   (skip rows whose index is skip_index)
T4R_HD void index_add_item(float* dst, const int64_t* idx, const float* src, int64_t ld_src, int col, int width,
                           int64_t skip_index, int64_t i) {
  const int64_t r = i / width;
  const int c = static_cast<int>(i % width);
  const int64_t row = idx[r];
  if (row == skip_index) return;
  T4R_ATOMIC_ADD(dst + row * width + c, src[r * ld_src + col + c]);
}
// soft embedding (features/embedding.py:517-556): p = softmax(x w + b) over the n rows of the table, out = p^T table;
// p is kept for the backward.  One scalar per item.
T4R_HD void soft_emb_fwd_row(const float* x, const float* w, const float* b, const float* table, int n, int dim, float* p,
                             float* out, int64_t r) {
  const float xv = x[r];
  float* pr = p + r * n;
  float mx = -INFINITY;
  for (int j = 0; j < n; ++j) mx = fmaxf(mx, xv * w[j] + b[j]);
  float s = 0.f;
  for (int j = 0; j < n; ++j) {
    const float e = expf(xv * w[j] + b[j] - mx);
    pr[j] = e;
    s += e;
  }
  const float inv = 1.0f / s;
  for (int j = 0; j < n; ++j) pr[j] *= inv;
  for (int c = 0; c < dim; ++c) {
    float acc = 0.f;
    for (int j = 0; j < n; ++j) acc += pr[j] * table[static_cast<int64_t>(j) * dim + c];
    out[r * dim + c] = acc;
  }
}
// its backward per scalar: dp_j = <dout_r, table_j>, dlogit_j = p_j (dp_j - sum_k p_k dp_k); dlogit x for the slope's
// gradient.  (d table = p^T dout, d w = column sums of dlogit x, d b = column sums of dlogit: composed by the caller.)
T4R_HD void soft_emb_bwd_row(const float* x, const float* table, const float* p, const float* dout, int n, int dim,
                             float* dlogit, float* dlogit_x, int64_t r) {
  const float* pr = p + r * n;
  const float* dr = dout + r * dim;
  float s = 0.f;
  for (int j = 0; j < n; ++j) {
    float dp = 0.f;
    for (int c = 0; c < dim; ++c) dp += dr[c] * table[static_cast<int64_t>(j) * dim + c];
    dlogit[r * n + j] = dp;
    s += pr[j] * dp;
  }
  for (int j = 0; j < n; ++j) {
    const float g = pr[j] * (dlogit[r * n + j] - s);
    dlogit[r * n + j] = g;
    dlogit_x[r * n + j] = g * x[r];
  }
}
// element-wise a + b (op 0) or a * b (op 1): the element-wise aggregations and their product rule
T4R_HD void binary_item(int op, const float* a, const float* b, float* out, int64_t i) {
  out[i] = op == 0 ? a[i] + b[i] : a[i] * b[i];
}
// AdamW (decoupled weight decay, torch.optim.AdamW's update rule), one element per item:
//   p *= 1 - lr wd;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= lr / bc1 * m / (sqrt(v) / sqrt(bc2) + eps)
T4R_HD void adamw_item(float* p, const float* g, float* m, float* v, float lr, float b1, float b2, float eps, float wd,
                       float bc1, float bc2_sqrt, int64_t i) {
  float pi = p[i] * (1.0f - lr * wd);
  const float gi = g[i];
  const float mi = b1 * m[i] + (1.0f - b1) * gi;
  const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = pi - (lr / bc1) * (mi / denom);
}
// column sums over a slab of rows: one item = (column, slab)
T4R_HD void col_sum_item(const float* x, int64_t M, int64_t N, int rows_per_slab, float* out, int64_t i) {
  const int64_t c = i % N, slab = i / N;
  const int64_t r0 = slab * rows_per_slab, r1 = (r0 + rows_per_slab < M) ? r0 + rows_per_slab : M;
  float s = 0.f;
  for (int64_t r = r0; r < r1; ++r) s += x[r * N + c];
  T4R_ATOMIC_ADD(out + c, s);
}

// ---------------------------------------------------------------------------------------------- LayerNorm (rows)
T4R_HD void ln_fwd_row(const float* x, const float* g, const float* b, int d, float eps, float* y, int64_t row) {
  const float* xr = x + row * d;
  float mean = 0.f;
  for (int c = 0; c < d; ++c) mean += xr[c];
  mean /= d;
  float var = 0.f;
  for (int c = 0; c < d; ++c) { const float t = xr[c] - mean; var += t * t; }
  const float rstd = 1.0f / sqrtf(var / d + eps);
  for (int c = 0; c < d; ++c) y[row * d + c] = (xr[c] - mean) * rstd * g[c] + b[c];
}
// dx = rstd (g dy - mean(g dy) - xhat mean(g dy xhat)) (+ add);  dy_xhat = dy xhat
T4R_HD void ln_bwd_row(const float* x, const float* g, int d, float eps, const float* dy, const float* add, float* dx,
                       float* dy_xhat, int64_t row) {
  const float* xr = x + row * d;
  const float* dyr = dy + row * d;
  float mean = 0.f;
  for (int c = 0; c < d; ++c) mean += xr[c];
  mean /= d;
  float var = 0.f;
  for (int c = 0; c < d; ++c) { const float t = xr[c] - mean; var += t * t; }
  const float rstd = 1.0f / sqrtf(var / d + eps);
  float m1 = 0.f, m2 = 0.f;
  for (int c = 0; c < d; ++c) {
    const float xh = (xr[c] - mean) * rstd, gd = g[c] * dyr[c];
    m1 += gd;
    m2 += gd * xh;
  }
  m1 /= d;
  m2 /= d;
  for (int c = 0; c < d; ++c) {
    const float xh = (xr[c] - mean) * rstd, gd = g[c] * dyr[c];
    float v = rstd * (gd - m1 - xh * m2);
    if (add) v += add[row * d + c];
    dx[row * d + c] = v;
    dy_xhat[row * d + c] = dyr[c] * xh;   // dgamma = its column sum (dbeta = the column sum of dy): no atomics per element
  }
}

// ---------------------------------------------------------------------------------------------- attention backward
// One item = (session b, head h): it owns the head's slices of dq / dk / dv for all rows of its session, its own
// [2L, dh] slice of dR_part[b] and [dh] slices of drw_part[b] / drr_part[b], so nothing is shared between items and
// there is no atomic; dR / drw / drr are the sums of the partials over the sessions (separate reductions).
// qkv [M, 3d] fp32 (q | k | v), R [2L, d], rw / rr [d] (XLNet; null for the causal GPT-2 form), dout [M, d].
// Scores are recomputed per query row (L <= 64), then
//   dp_j = do_i . v_j ;  ds_j = p_j (dp_j - sum_k p_k dp_k) / sqrt(dh)
//   dq_i = sum_j ds_j (k_j + R_m)        dk_j += ds_j (q_i + rw)        dv_j += p_j do_i
//   dR_m += ds_j (q_i + rr)              drw  += ds_j k_j               drr  += ds_j R_m          m = j + L - i
constexpr int kAttnMaxL = 64;
// flat index of probability (stream st, session b, head h, query i, key j): the [n_streams, B, H, L, L] order of HF's
// attn_prob ("bnij") -- what the dropout mask of the probabilities is keyed on
T4R_HD uint64_t attn_prob_index(int st, int B, int64_t b, int H, int h, int L, int i, int j) {
  return (((static_cast<uint64_t>(st) * B + b) * H + h) * L + i) * L + j;
}
// plm_mask != nullptr: XLNet's two-stream form (permutation language modeling).  qkv / dout / dqkv then hold 2 B L rows
// (content stream h, then query stream g); the item runs both streams: queries from the stream's rows, keys / values
// from the h rows, score (i, j) = -1e30 where plm_mask[b, i, j] (h: except i == j).  The backward uses the same
// ds_j = p_j (dp_j - sum) for every j, as autograd of HF's `score - 1e30 * mask` does (d/d score = 1): a masked entry
// next to a visible one has p_j = 0 exactly and passes nothing; in a FULLY masked row (the first target of the
// permutation when every item of a session is a target) p is uniform and the gradient flows as in the reference.
T4R_HD void attn_bwd_item(const float* qkv, const float* R, const float* rw, const float* rr, const float* dout, int B,
                          int L, int d, int H, float* dqkv, float* dR_part, float* drw_part, float* drr_part,
                          const uint8_t* plm_mask, int64_t item, float p_drop = 0.f, uint64_t seed = 0,
                          uint32_t site = 0) {
  const int dh = d / H;
  const int h = static_cast<int>(item % H);
  const int64_t b = item / H;
  const bool rel = R != nullptr;
  const int n_streams = plm_mask ? 2 : 1;
  const int64_t M = static_cast<int64_t>(B) * L;
  const float scale = 1.0f / sqrtf(static_cast<float>(dh));
  // zero what this item owns (for both streams' rows)
  for (int st = 0; st < n_streams; ++st)
    for (int r = 0; r < L; ++r)
      for (int part = 0; part < 3; ++part)
        for (int c = 0; c < dh; ++c) dqkv[(st * M + b * L + r) * 3 * d + part * d + h * dh + c] = 0.f;
  if (rel) {
    for (int m = 0; m < 2 * L; ++m)
      for (int c = 0; c < dh; ++c) dR_part[(b * 2 * L + m) * d + h * dh + c] = 0.f;
    for (int c = 0; c < dh; ++c) { drw_part[b * d + h * dh + c] = 0.f; drr_part[b * d + h * dh + c] = 0.f; }
  }
  float s[kAttnMaxL], dp[kAttnMaxL];
  for (int st = 0; st < n_streams; ++st)
  for (int i = 0; i < L; ++i) {
    const float* q = qkv + (st * M + b * L + i) * 3 * d + h * dh;
    const float* dor = dout + (st * M + b * L + i) * d + h * dh;
    float mx = -INFINITY;
    for (int j = 0; j < L; ++j) {
      const bool masked = plm_mask && plm_mask[(b * L + i) * L + j] && !(st == 0 && i == j);
      if (masked) { s[j] = -1e30f; mx = fmaxf(mx, s[j]); continue; }
      if (!rel && j > i) { s[j] = -INFINITY; continue; }
      const float* k = qkv + (b * L + j) * 3 * d + d + h * dh;
      float acc = 0.f;
      if (rel) {
        const float* Rm = R + static_cast<int64_t>(j + L - i) * d + h * dh;
        for (int c = 0; c < dh; ++c) acc += (q[c] + rw[h * dh + c]) * k[c] + (q[c] + rr[h * dh + c]) * Rm[c];
      } else {
        for (int c = 0; c < dh; ++c) acc += q[c] * k[c];
      }
      s[j] = acc * scale;
      mx = fmaxf(mx, s[j]);
    }
    float sum = 0.f;
    for (int j = 0; j < L; ++j) { s[j] = (s[j] == -INFINITY) ? 0.f : expf(s[j] - mx); sum += s[j]; }
    const float inv = 1.0f / sum;
    float dot = 0.f;
    float mk[kAttnMaxL];   // dropout of the probabilities (HF:xlnet:129 / HF:gpt2:66): a_i = sum_j (p_j mk_j) v_j
    for (int j = 0; j < L; ++j) {
      s[j] *= inv;                                            // p_j
      mk[j] = keep_scale(seed, site, attn_prob_index(st, B, b, H, h, L, i, j), p_drop);
      const float* v = qkv + (b * L + j) * 3 * d + 2 * d + h * dh;
      float acc = 0.f;
      for (int c = 0; c < dh; ++c) acc += dor[c] * v[c];
      dp[j] = acc * mk[j];                                    // d loss / d p_j
      dot += s[j] * dp[j];
    }
    float* dq = dqkv + (st * M + b * L + i) * 3 * d + h * dh;
    for (int j = 0; j < L; ++j) {
      if (!rel && j > i) continue;
      const float ds = s[j] * (dp[j] - dot) * scale;
      const float p = s[j] * mk[j];                            // the (dropped) weight v_j entered the output with
      const float* k = qkv + (b * L + j) * 3 * d + d + h * dh;
      float* dk = dqkv + (b * L + j) * 3 * d + d + h * dh;
      float* dv = dqkv + (b * L + j) * 3 * d + 2 * d + h * dh;
      if (rel) {
        const int64_t m = j + L - i;
        const float* Rm = R + m * d + h * dh;
        float* dRm = dR_part + (b * 2 * L + m) * d + h * dh;
        for (int c = 0; c < dh; ++c) {
          dq[c] += ds * (k[c] + Rm[c]);
          dk[c] += ds * (q[c] + rw[h * dh + c]);
          dv[c] += p * dor[c];
          dRm[c] += ds * (q[c] + rr[h * dh + c]);
          drw_part[b * d + h * dh + c] += ds * k[c];
          drr_part[b * d + h * dh + c] += ds * Rm[c];
        }
      } else {
        for (int c = 0; c < dh; ++c) {
          dq[c] += ds * k[c];
          dk[c] += ds * q[c];
          dv[c] += p * dor[c];
        }
      }
    }
  }
}

// Attention FORWARD of the training graph in the same one-item-per-(session, head) form, with dropout of the
// probabilities (training mode with a dropout rate; without one the training forward runs the inference kernels).
// out [n_streams * M, d] fp32.
T4R_HD void attn_fwd_item(const float* qkv, const float* R, const float* rw, const float* rr, int B, int L, int d, int H,
                          float* out, const uint8_t* plm_mask, float p_drop, uint64_t seed, uint32_t site, int64_t item) {
  const int dh = d / H;
  const int h = static_cast<int>(item % H);
  const int64_t b = item / H;
  const bool rel = R != nullptr;
  const int n_streams = plm_mask ? 2 : 1;
  const int64_t M = static_cast<int64_t>(B) * L;
  const float scale = 1.0f / sqrtf(static_cast<float>(dh));
  float s[kAttnMaxL];
  for (int st = 0; st < n_streams; ++st)
  for (int i = 0; i < L; ++i) {
    const float* q = qkv + (st * M + b * L + i) * 3 * d + h * dh;
    float mx = -INFINITY;
    for (int j = 0; j < L; ++j) {
      const bool masked = plm_mask && plm_mask[(b * L + i) * L + j] && !(st == 0 && i == j);
      if (masked) { s[j] = -1e30f; mx = fmaxf(mx, s[j]); continue; }
      if (!rel && j > i) { s[j] = -INFINITY; continue; }
      const float* k = qkv + (b * L + j) * 3 * d + d + h * dh;
      float acc = 0.f;
      if (rel) {
        const float* Rm = R + static_cast<int64_t>(j + L - i) * d + h * dh;
        for (int c = 0; c < dh; ++c) acc += (q[c] + rw[h * dh + c]) * k[c] + (q[c] + rr[h * dh + c]) * Rm[c];
      } else {
        for (int c = 0; c < dh; ++c) acc += q[c] * k[c];
      }
      s[j] = acc * scale;
      mx = fmaxf(mx, s[j]);
    }
    float sum = 0.f;
    for (int j = 0; j < L; ++j) { s[j] = (s[j] == -INFINITY) ? 0.f : expf(s[j] - mx); sum += s[j]; }
    const float inv = 1.0f / sum;
    float* o = out + (st * M + b * L + i) * d + h * dh;
    for (int c = 0; c < dh; ++c) o[c] = 0.f;
    for (int j = 0; j < L; ++j) {
      const float p = s[j] * inv * keep_scale(seed, site, attn_prob_index(st, B, b, H, h, L, i, j), p_drop);
      if (p == 0.f) continue;
      const float* v = qkv + (b * L + j) * 3 * d + 2 * d + h * dh;
      for (int c = 0; c < dh; ++c) o[c] += p * v[c];
    }
  }
}

// ---------------------------------------------------------------------------------------------- attention backward, device
// The same arithmetic as attn_bwd_item (which stays the host twin and the reference the device is tested against), but
// one WARP per (session, head) instead of one thread: the per-thread form was 49 % of a training step (24 ms per call
// at config-2 shapes: 16 384 threads each walking L^2 dh serially).  Here q / k / v / dout of the head and its R slice
// sit in shared memory; lanes run over the keys j for the scores, the softmax and dp (warp reductions for max / sum /
// dot), then over the head dimension c for the accumulations, where lane c owns column c of dq_i, dk, dv, dR and of
// drw / drr -- no atomics, same ownership as the item form.  dh <= 64, L <= 64.
template <int DH>
__global__ void __launch_bounds__(32)
attn_bwd_warp_kernel(const float* __restrict__ qkv, const float* __restrict__ R, const float* __restrict__ rw,
                     const float* __restrict__ rr, const float* __restrict__ dout, int B, int L, int d, int H,
                     float* __restrict__ dqkv, float* __restrict__ dR_part, float* __restrict__ drw_part,
                     float* __restrict__ drr_part, const uint8_t* __restrict__ plm_mask, float p_drop, uint64_t seed,
                     uint32_t site) {
  constexpr int LD = DH + 1;                     // padded rows: lanes over j read column c without bank conflicts
  constexpr int CPL = (DH + 31) / 32;            // columns per lane in the accumulation phase
  extern __shared__ float sm[];
  const int item = blockIdx.x;
  const int h = item % H;
  const int64_t b = item / H;
  const int lane = threadIdx.x;
  const bool rel = R != nullptr;
  const int n_streams = plm_mask ? 2 : 1;
  const int64_t M = static_cast<int64_t>(B) * L;
  const float scale = 1.0f / sqrtf(static_cast<float>(DH));
  float* ks = sm;                      // [L][LD]   keys of the content stream
  float* vs = ks + L * LD;             // [L][LD]
  float* qs = vs + L * LD;             // [L][LD]   queries of the current stream
  float* dos = qs + L * LD;            // [L][LD]   dout of the current stream
  float* Rs = dos + L * LD;            // [2L][LD]  (relative form)
  float* dks = Rs + (rel ? 2 * L * LD : 0);   // [L][DH] accumulators
  float* dvs = dks + L * DH;
  float* dRs = dvs + L * DH;           // [2L][DH]  (relative form)
  float* ps = dRs + (rel ? 2 * L * DH : 0);   // [L] probabilities of the current query row
  float* dss = ps + L;                 // [L] ds_j
  float* mks = dss + L;                // [L] dropout keep scales
  for (int idx = lane; idx < L * DH; idx += 32) {
    const int r = idx / DH, c = idx % DH;
    const float* row = qkv + (b * L + r) * 3 * d + h * DH + c;
    ks[r * LD + c] = row[d];
    vs[r * LD + c] = row[2 * d];
    dks[idx] = 0.f;
    dvs[idx] = 0.f;
  }
  if (rel)
    for (int idx = lane; idx < 2 * L * DH; idx += 32) {
      const int r = idx / DH, c = idx % DH;
      Rs[r * LD + c] = R[static_cast<int64_t>(r) * d + h * DH + c];
      dRs[idx] = 0.f;
    }
  float rwv[CPL], rrv[CPL], drw_acc[CPL], drr_acc[CPL];
#pragma unroll
  for (int u = 0; u < CPL; ++u) {
    const int c = lane + 32 * u;
    rwv[u] = (rel && c < DH) ? rw[h * DH + c] : 0.f;
    rrv[u] = (rel && c < DH) ? rr[h * DH + c] : 0.f;
    drw_acc[u] = 0.f; drr_acc[u] = 0.f;
  }
  for (int st = 0; st < n_streams; ++st) {
    __syncwarp();
    for (int idx = lane; idx < L * DH; idx += 32) {
      const int r = idx / DH, c = idx % DH;
      qs[r * LD + c] = qkv[(st * M + b * L + r) * 3 * d + h * DH + c];
      dos[r * LD + c] = dout[(st * M + b * L + r) * d + h * DH + c];
    }
    __syncwarp();
    for (int i = 0; i < L; ++i) {
      // ---- scores, softmax, dp: lanes over the keys
      float sj[2], dpj[2], mkj[2];
      float mx = -INFINITY;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = lane + 32 * u;
        float s = -INFINITY, dp = 0.f, mk = 1.f;
        if (j < L) {
          const bool masked = plm_mask && plm_mask[(b * L + i) * L + j] && !(st == 0 && i == j);
          if (masked) s = -1e30f;
          else if (rel || j <= i) {
            float acc = 0.f;
            if (rel) {
              const float* Rm = Rs + (j + L - i) * LD;
              for (int c = 0; c < DH; ++c) {
                const float q = qs[i * LD + c];
                acc += (q + rw[h * DH + c]) * ks[j * LD + c] + (q + rr[h * DH + c]) * Rm[c];
              }
            } else {
              for (int c = 0; c < DH; ++c) acc += qs[i * LD + c] * ks[j * LD + c];
            }
            s = acc * scale;
          }
          mk = keep_scale(seed, site, attn_prob_index(st, B, b, H, h, L, i, j), p_drop);
          for (int c = 0; c < DH; ++c) dp += dos[i * LD + c] * vs[j * LD + c];
          dp *= mk;
        }
        sj[u] = s; dpj[u] = dp; mkj[u] = mk;
        mx = fmaxf(mx, s);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
#pragma unroll
      for (int u = 0; u < 2; ++u) { sj[u] = (sj[u] == -INFINITY) ? 0.f : expf(sj[u] - mx); sum += sj[u]; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float inv = 1.0f / sum;
      float dot = 0.f;
#pragma unroll
      for (int u = 0; u < 2; ++u) { sj[u] *= inv; dot += sj[u] * dpj[u]; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = lane + 32 * u;
        if (j < L) {
          const bool live = rel || j <= i;
          ps[j] = live ? sj[u] * mkj[u] : 0.f;                            // the (dropped) weight of v_j
          dss[j] = live ? sj[u] * (dpj[u] - dot) * scale : 0.f;
        }
      }
      __syncwarp();
      // ---- accumulations: lane c owns column c
#pragma unroll
      for (int u = 0; u < CPL; ++u) {
        const int c = lane + 32 * u;
        if (c < DH) {
          const float q = qs[i * LD + c], dor = dos[i * LD + c];
          float dq = 0.f;
          const int jend = rel ? L : i + 1;
          for (int j = 0; j < jend; ++j) {
            const float ds = dss[j], p = ps[j];
            const float k = ks[j * LD + c];
            if (rel) {
              const int m = j + L - i;
              const float Rm = Rs[m * LD + c];
              dq += ds * (k + Rm);
              dks[j * DH + c] += ds * (q + rwv[u]);
              dRs[m * DH + c] += ds * (q + rrv[u]);
              drw_acc[u] += ds * k;
              drr_acc[u] += ds * Rm;
            } else {
              dq += ds * k;
              dks[j * DH + c] += ds * q;
            }
            dvs[j * DH + c] += p * dor;
          }
          dqkv[(st * M + b * L + i) * 3 * d + h * DH + c] = dq;
        }
      }
      __syncwarp();
    }
  }
  // ---- what the item owns: dk / dv of the content stream's rows (the query stream's stay zero), dR / drw / drr partials
  for (int idx = lane; idx < L * DH; idx += 32) {
    const int r = idx / DH, c = idx % DH;
    float* row = dqkv + (b * L + r) * 3 * d + h * DH + c;
    row[d] = dks[idx];
    row[2 * d] = dvs[idx];
    if (n_streams == 2) {
      float* row2 = dqkv + (M + b * L + r) * 3 * d + h * DH + c;
      row2[d] = 0.f;
      row2[2 * d] = 0.f;
    }
  }
  if (rel) {
    for (int idx = lane; idx < 2 * L * DH; idx += 32) {
      const int r = idx / DH, c = idx % DH;
      dR_part[(b * 2 * L + r) * d + h * DH + c] = dRs[idx];
    }
#pragma unroll
    for (int u = 0; u < CPL; ++u) {
      const int c = lane + 32 * u;
      if (c < DH) { drw_part[b * d + h * DH + c] = drw_acc[u]; drr_part[b * d + h * DH + c] = drr_acc[u]; }
    }
  }
}
static size_t attn_bwd_warp_smem(int L, int DH, bool rel) {
  const size_t LD = DH + 1;
  return sizeof(float) * (4 * L * LD + (rel ? 2 * L * LD : 0) + 2 * static_cast<size_t>(L) * DH +
                          (rel ? 2 * static_cast<size_t>(L) * DH : 0) + 3 * L);
}
template <int DH>
static int launch_attn_bwd_warp(const float* qkv, const float* R, const float* rw, const float* rr, const float* dout, int B,
                                int L, int d, int H, float* dqkv, float* dR_part, float* drw_part, float* drr_part,
                                const uint8_t* plm_mask, float p_drop, uint64_t seed, uint32_t site, cudaStream_t s) {
  const size_t smem = attn_bwd_warp_smem(L, DH, R != nullptr);
  auto kern = attn_bwd_warp_kernel<DH>;
  T4R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<static_cast<unsigned>(B) * H, 32, smem, s>>>(qkv, R, rw, rr, dout, B, L, d, H, dqkv, dR_part, drw_part, drr_part,
                                                     plm_mask, p_drop, seed, site);
  T4R_LAUNCH_CHECK("attn_bwd_warp_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------- launch plumbing
template <typename F>
__global__ void __launch_bounds__(256) items_kernel(int64_t n, F f) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) f(i);
}
template <typename F>
static int run_items(int64_t n, F f, const char* name, void* stream, int on_host) {
  if (n <= 0) return 0;
  if (on_host) {
    for (int64_t i = 0; i < n; ++i) f(i);
    return 0;
  }
  T4R_REQUIRE(n < (1ll << 31) * 256, "%s: too many work items", name);
  items_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(n, f);
  T4R_LAUNCH_CHECK(name);
  return 0;
}
static int zero(void* p, size_t bytes, void* stream, int on_host) {
  if (on_host) { memset(p, 0, bytes); return 0; }
  T4R_CUDA(cudaMemsetAsync(p, 0, bytes, static_cast<cudaStream_t>(stream)));
  return 0;
}

}  // namespace t4r

using namespace t4r;
#define T4R_ITEMS(n, name, ...) return run_items((n), [=] __host__ __device__(int64_t i) { __VA_ARGS__; }, name, stream, on_host)

extern "C" int t4r_train_transpose(const float* x, int64_t R, int64_t C, float* out, void* stream, int on_host) {
  T4R_REQUIRE(x && out && R > 0 && C > 0, "train_transpose: bad arguments");
  T4R_ITEMS(R * C, "train_transpose", transpose_item(x, R, C, out, i));
}
extern "C" int t4r_train_act_fwd(int kind, const float* x, float* y, int64_t n, void* stream, int on_host) {
  T4R_REQUIRE(x && y && n > 0, "train_act_fwd: bad arguments");
  T4R_ITEMS(n, "train_act_fwd", act_fwd_item(kind, x, y, i));
}
extern "C" int t4r_train_act_bwd(int kind, const float* pre, const float* dy, float* dx, int64_t n, void* stream,
                                 int on_host) {
  T4R_REQUIRE(pre && dy && dx && n > 0, "train_act_bwd: bad arguments");
  T4R_ITEMS(n, "train_act_bwd", act_bwd_item(kind, pre, dy, dx, i));
}
extern "C" int t4r_train_add_positions(const float* x, const float* wpe, int B, int L, int d, float* y, void* stream,
                                       int on_host) {
  T4R_REQUIRE(x && wpe && y && B > 0 && L > 0 && d > 0, "train_add_positions: bad arguments");
  T4R_ITEMS(static_cast<int64_t>(B) * L * d, "train_add_positions", add_pos_item(x, wpe, L, d, y, i));
}
extern "C" int t4r_train_sum_sessions(const float* x, int B, int L, int d, float* out, void* stream, int on_host) {
  T4R_REQUIRE(x && out && B > 0 && L > 0 && d > 0, "train_sum_sessions: bad arguments");
  T4R_ITEMS(static_cast<int64_t>(L) * d, "train_sum_sessions", sum_sessions_item(x, B, L, d, out, i));
}
extern "C" int t4r_train_row_codes_fwd(const float* y, const uint8_t* code, const float* mask_vec, int64_t M, int d,
                                       float* out, void* stream, int on_host) {
  T4R_REQUIRE(y && code && mask_vec && out && M > 0 && d > 0, "train_row_codes_fwd: bad arguments");
  T4R_ITEMS(M * d, "train_row_codes_fwd", row_codes_fwd_item(y, code, mask_vec, d, out, i));
}
static int col_sum_impl(const float* x, int64_t M, int64_t N, float* out, void* stream, int on_host);
extern "C" int t4r_train_row_codes_bwd(const float* dx, const uint8_t* code, int64_t M, int d, float* dy, float* dmask,
                                       float* tmp /*[M, d] scratch*/, void* stream, int on_host) {
  T4R_REQUIRE(dx && code && dy && dmask && tmp && M > 0 && d > 0, "train_row_codes_bwd: bad arguments");
  T4R_TRY(run_items(M * d, [=] __host__ __device__(int64_t i) { row_codes_bwd_item(dx, code, d, dy, tmp, i); },
                    "train_row_codes_bwd", stream, on_host));
  return col_sum_impl(tmp, M, d, dmask, stream, on_host);
}
extern "C" int t4r_train_gather_rows(const float* x, const int32_t* idx, int64_t n, int d, float* out, void* stream,
                                     int on_host) {
  T4R_REQUIRE(x && idx && out && n >= 0 && d > 0, "train_gather_rows: bad arguments");
  T4R_ITEMS(n * d, "train_gather_rows", gather_rows_item(x, idx, d, out, i));
}
extern "C" int t4r_train_scatter_rows(const float* src, const int32_t* idx, int64_t n, int d, int64_t out_rows, float* out,
                                      void* stream, int on_host) {
  T4R_REQUIRE(src && idx && out && n >= 0 && d > 0 && out_rows > 0, "train_scatter_rows: bad arguments");
  T4R_TRY(zero(out, sizeof(float) * out_rows * d, stream, on_host));
  T4R_ITEMS(n * d, "train_scatter_rows", scatter_rows_item(src, idx, d, out, i));
}
extern "C" int t4r_train_softmax_ce_bwd(float* z, const float* lse, const int64_t* labels, int64_t T, int64_t Vc,
                                        int64_t v0, float scale, float label_smoothing, int64_t V_total, void* stream,
                                        int on_host) {
  T4R_REQUIRE(z && lse && labels && T > 0 && Vc > 0 && V_total > 0, "train_softmax_ce_bwd: bad arguments");
  const float inv_V = 1.0f / static_cast<float>(V_total);
  T4R_ITEMS(T * Vc, "train_softmax_ce_bwd", softmax_ce_bwd_item(z, lse, labels, Vc, v0, scale, label_smoothing, inv_V, i));
}
extern "C" int t4r_train_sampled_ce_bwd(float* z, const float* lse, const int64_t* labels, const float* col_bias,
                                        const int64_t* col_ids, int64_t T, int64_t S, float inv_tau, float scale,
                                        void* stream, int on_host) {
  T4R_REQUIRE(z && lse && labels && col_bias && col_ids && T > 0 && S > 0, "train_sampled_ce_bwd: bad arguments");
  T4R_ITEMS(T * S, "train_sampled_ce_bwd", sampled_ce_bwd_item(z, lse, labels, col_bias, col_ids, S, inv_tau, scale, i));
}
extern "C" int t4r_train_soft_emb_fwd(const float* x, const float* w, const float* b, const float* table, int64_t M, int n,
                                      int dim, float* p, float* out, void* stream, int on_host) {
  T4R_REQUIRE(x && w && b && table && p && out && M > 0 && n > 0 && dim > 0, "train_soft_emb_fwd: bad arguments");
  T4R_ITEMS(M, "train_soft_emb_fwd", soft_emb_fwd_row(x, w, b, table, n, dim, p, out, i));
}
extern "C" int t4r_train_soft_emb_bwd(const float* x, const float* table, const float* p, const float* dout, int64_t M, int n,
                                      int dim, float* dlogit, float* dlogit_x, void* stream, int on_host) {
  T4R_REQUIRE(x && table && p && dout && dlogit && dlogit_x && M > 0 && n > 0 && dim > 0, "train_soft_emb_bwd: bad arguments");
  T4R_ITEMS(M, "train_soft_emb_bwd", soft_emb_bwd_row(x, table, p, dout, n, dim, dlogit, dlogit_x, i));
}
extern "C" int t4r_train_binary(int op, const float* a, const float* b, float* out, int64_t n, void* stream, int on_host) {
  T4R_REQUIRE(a && b && out && n > 0 && (op == 0 || op == 1), "train_binary: bad arguments");
  T4R_ITEMS(n, "train_binary", binary_item(op, a, b, out, i));
}
// One AdamW step on a flat fp32 tensor (step = 1, 2, ...: the bias corrections 1 - beta^step are formed here in double)
extern "C" int t4r_train_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                               float eps, float weight_decay, int step, void* stream, int on_host) {
  T4R_REQUIRE(p && g && m && v && n > 0 && step >= 1, "train_adamw: bad arguments");
  const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), step));
  const float bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), step)));
  T4R_ITEMS(n, "train_adamw", adamw_item(p, g, m, v, lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt, i));
}
extern "C" int t4r_train_index_add_rows(float* dst, const int64_t* idx, const float* src, int64_t n, int64_t ld_src,
                                        int col, int width, int64_t skip_index, void* stream, int on_host) {
  T4R_REQUIRE(dst && idx && src && n > 0 && width > 0 && col >= 0 && col + width <= ld_src, "train_index_add_rows: bad arguments");
  T4R_ITEMS(n * width, "train_index_add_rows", index_add_item(dst, idx, src, ld_src, col, width, skip_index, i));
}
// column sums: one item per (column, slab of 256 rows), one atomicAdd per item -> (M / 256) N atomics in total
static int col_sum_impl(const float* x, int64_t M, int64_t N, float* out, void* stream, int on_host) {
  T4R_TRY(zero(out, sizeof(float) * N, stream, on_host));
  const int rows_per_slab = 256;
  const int64_t slabs = (M + rows_per_slab - 1) / rows_per_slab;
  T4R_ITEMS(slabs * N, "train_col_sum", col_sum_item(x, M, N, rows_per_slab, out, i));
}
extern "C" int t4r_train_col_sum(const float* x, int64_t M, int64_t N, float* out, void* stream, int on_host) {
  T4R_REQUIRE(x && out && M > 0 && N > 0, "train_col_sum: bad arguments");
  return col_sum_impl(x, M, N, out, stream, on_host);
}
extern "C" int t4r_train_layer_norm_fwd(const float* x, const float* gamma, const float* beta, int64_t M, int d, float eps,
                                        float* y, void* stream, int on_host) {
  T4R_REQUIRE(x && gamma && beta && y && M > 0 && d > 0, "train_layer_norm_fwd: bad arguments");
  T4R_ITEMS(M, "train_layer_norm_fwd", ln_fwd_row(x, gamma, beta, d, eps, y, i));
}
extern "C" int t4r_train_layer_norm_bwd(const float* x, const float* gamma, int64_t M, int d, float eps, const float* dy,
                                        const float* add, float* dx, float* dgamma, float* dbeta,
                                        float* tmp /*[M, d] scratch*/, void* stream, int on_host) {
  T4R_REQUIRE(x && gamma && dy && dx && dgamma && dbeta && tmp && M > 0 && d > 0, "train_layer_norm_bwd: bad arguments");
  T4R_TRY(run_items(M, [=] __host__ __device__(int64_t i) { ln_bwd_row(x, gamma, d, eps, dy, add, dx, tmp, i); },
                    "train_layer_norm_bwd", stream, on_host));
  T4R_TRY(col_sum_impl(tmp, M, d, dgamma, stream, on_host));   // dgamma = sum_rows dy * xhat
  return col_sum_impl(dy, M, d, dbeta, stream, on_host);         // dbeta  = sum_rows dy
}
// R / rw / rr / dR / drw / drr all NULL selects the causal (GPT-2) form.  part: scratch of B (2L + 2) d floats (relative
// form): per-session partials of dR / drw / drr, reduced over the sessions here.
extern "C" int t4r_train_dropout(const float* x, float* y, int64_t n, float p, uint64_t seed, uint32_t site, void* stream,
                                 int on_host) {
  T4R_REQUIRE(x && y && n > 0 && p >= 0.f && p < 1.f, "train_dropout: bad arguments (0 <= p < 1)");
  T4R_ITEMS(n, "train_dropout", dropout_item(x, y, p, seed, site, i));
}

extern "C" int t4r_train_attn_drop_fwd(const float* qkv, const float* R, const float* rw, const float* rr, int B, int L, int d,
                                       int H, const uint8_t* plm_mask, float p_drop, uint64_t seed, uint32_t site, float* out,
                                       void* stream, int on_host) {
  T4R_REQUIRE(qkv && out && B > 0 && L > 0 && L <= kAttnMaxL && H > 0 && d % H == 0, "train_attn_drop_fwd: bad arguments (L <= 64)");
  T4R_REQUIRE((R == nullptr) == (rw == nullptr) && (R == nullptr) == (rr == nullptr), "train_attn_drop_fwd: R, rw, rr go together");
  T4R_REQUIRE(plm_mask == nullptr || R != nullptr, "train_attn_drop_fwd: the two-stream (PLM) form is XLNet's relative attention");
  T4R_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "train_attn_drop_fwd: 0 <= p < 1");
  T4R_ITEMS(static_cast<int64_t>(B) * H, "train_attn_drop_fwd",
            attn_fwd_item(qkv, R, rw, rr, B, L, d, H, out, plm_mask, p_drop, seed, site, i));
}

static int attn_bwd_impl(const float* qkv, const float* R, const float* rw, const float* rr, const float* dout, int B, int L,
                         int d, int H, float* dqkv, float* dR, float* drw, float* drr, float* part, const uint8_t* plm_mask,
                         float p_drop, uint64_t seed, uint32_t site, void* stream, int on_host);
extern "C" int t4r_train_attn_drop_bwd(const float* qkv, const float* R, const float* rw, const float* rr, const float* dout,
                                       int B, int L, int d, int H, float* dqkv, float* dR, float* drw, float* drr, float* part,
                                       const uint8_t* plm_mask, float p_drop, uint64_t seed, uint32_t site, void* stream,
                                       int on_host) {
  T4R_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "train_attn_drop_bwd: 0 <= p < 1");
  return attn_bwd_impl(qkv, R, rw, rr, dout, B, L, d, H, dqkv, dR, drw, drr, part, plm_mask, p_drop, seed, site, stream, on_host);
}
extern "C" int t4r_train_attn_bwd(const float* qkv, const float* R, const float* rw, const float* rr, const float* dout,
                                  int B, int L, int d, int H, float* dqkv, float* dR, float* drw, float* drr, float* part,
                                  const uint8_t* plm_mask, void* stream, int on_host) {
  return attn_bwd_impl(qkv, R, rw, rr, dout, B, L, d, H, dqkv, dR, drw, drr, part, plm_mask, 0.f, 0, 0, stream, on_host);
}
static int attn_bwd_impl(const float* qkv, const float* R, const float* rw, const float* rr, const float* dout, int B, int L,
                         int d, int H, float* dqkv, float* dR, float* drw, float* drr, float* part, const uint8_t* plm_mask,
                         float p_drop, uint64_t seed, uint32_t site, void* stream, int on_host) {
  T4R_REQUIRE(qkv && dout && dqkv && B > 0 && L > 0 && L <= kAttnMaxL && H > 0 && d % H == 0, "train_attn_bwd: bad arguments (L <= 64)");
  const bool rel = R != nullptr;
  T4R_REQUIRE(!rel || (rw && rr && dR && drw && drr && part), "train_attn_bwd: the relative form needs R, both biases, their gradients and the scratch");
  T4R_REQUIRE(plm_mask == nullptr || rel, "train_attn_bwd: the two-stream (PLM) form is XLNet's relative attention");
  float* dR_part = part;
  float* drw_part = rel ? part + static_cast<int64_t>(B) * 2 * L * d : nullptr;
  float* drr_part = rel ? drw_part + static_cast<int64_t>(B) * d : nullptr;
  const int dh = d / H;
  const bool warp_form = !on_host && (dh == 16 || dh == 32 || dh == 64) &&
                         attn_bwd_warp_smem(L, dh, rel) <= 200 * 1024 && getenv("T4R_TRAIN_ATTN_ITEMS") == nullptr;
  if (warp_form) {   // one warp per (session, head); same arithmetic, tested against the item form (the host twin)
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    if (dh == 16) T4R_TRY(launch_attn_bwd_warp<16>(qkv, R, rw, rr, dout, B, L, d, H, dqkv, dR_part, drw_part, drr_part, plm_mask, p_drop, seed, site, cs));
    else if (dh == 32) T4R_TRY(launch_attn_bwd_warp<32>(qkv, R, rw, rr, dout, B, L, d, H, dqkv, dR_part, drw_part, drr_part, plm_mask, p_drop, seed, site, cs));
    else T4R_TRY(launch_attn_bwd_warp<64>(qkv, R, rw, rr, dout, B, L, d, H, dqkv, dR_part, drw_part, drr_part, plm_mask, p_drop, seed, site, cs));
  } else {
    T4R_TRY(run_items(static_cast<int64_t>(B) * H, [=] __host__ __device__(int64_t i) {
              attn_bwd_item(qkv, R, rw, rr, dout, B, L, d, H, dqkv, dR_part, drw_part, drr_part, plm_mask, i, p_drop, seed, site); },
            "train_attn_bwd", stream, on_host));
  }
  if (!rel) return 0;
  T4R_TRY(run_items(static_cast<int64_t>(2) * L * d, [=] __host__ __device__(int64_t i) { sum_sessions_item(dR_part, B, 2 * L, d, dR, i); },
                    "train_attn_bwd_dR", stream, on_host));
  T4R_TRY(col_sum_impl(drw_part, B, d, drw, stream, on_host));
  return col_sum_impl(drr_part, B, d, drr, stream, on_host);
}

// ---------------------------------------------------------------------------------------------- forward pieces reused
// The training forward keeps q|k|v in fp32, so it runs the fp32-input attention kernels of the inference path
// (attn_kernel, t4r_kernels.cu: any L <= 64); outputs are split planes [2, M, d] like everywhere else.
extern "C" int t4r_train_xlnet_attn_fwd(const float* qkv, const float* R, const float* rw, const float* rr, int B, int L,
                                        int d, int H, void* out_planes, void* stream) {
  T4R_REQUIRE(qkv && R && rw && rr && out_planes && B > 0 && L > 0, "train_xlnet_attn_fwd: bad arguments");
  return launch_xlnet_attn(qkv, R, rw, rr, B, L, d, H, static_cast<__nv_bfloat16*>(out_planes),
                           static_cast<int64_t>(B) * L * d, static_cast<cudaStream_t>(stream));
}
extern "C" int t4r_train_causal_attn_fwd(const float* qkv, int B, int L, int d, int H, void* out_planes, void* stream) {
  T4R_REQUIRE(qkv && out_planes && B > 0 && L > 0, "train_causal_attn_fwd: bad arguments");
  return launch_causal_attn(qkv, B, L, d, H, static_cast<__nv_bfloat16*>(out_planes), static_cast<int64_t>(B) * L * d,
                            static_cast<cudaStream_t>(stream));
}
// R_l = pos(L, d) @ Wr_l for every layer: r_out [n_layer, 2L, d] fp32 (wr: HOST array of n_layer device pointers [d, d])
extern "C" int t4r_train_rel_pos_proj(const float* const* wr, int n_layer, int L, int d, float* r_out, void* stream) {
  T4R_REQUIRE(wr && r_out && n_layer >= 1 && n_layer <= T4R_MAX_FEATURES && L > 0 && d > 0, "train_rel_pos_proj: bad arguments");
  return launch_rel_pos_proj(wr, n_layer, L, d, r_out, nullptr, static_cast<cudaStream_t>(stream));
}

// two-stream (PLM) attention forward on split planes [2, 2 B L, 3d]: the MASKED tensor-path kernels of t4r_attn_mma.cu
extern "C" int t4r_train_xlnet_attn_plm_fwd(const void* qkv_planes, const void* r_planes, const float* rw, const float* rr,
                                            int B, int L, int d, int H, const uint8_t* plm_mask, void* out_planes,
                                            void* stream) {
  T4R_REQUIRE(qkv_planes && r_planes && rw && rr && plm_mask && out_planes && B > 0 && L > 0, "train_xlnet_attn_plm_fwd: bad arguments");
  const int64_t M2 = static_cast<int64_t>(2) * B * L;
  return launch_attn_mma_plm(static_cast<const __nv_bfloat16*>(qkv_planes), M2 * 3 * d,
                             static_cast<const __nv_bfloat16*>(r_planes), static_cast<int64_t>(2) * L * d, rw, rr, B, L, d, H,
                             static_cast<__nv_bfloat16*>(out_planes), M2 * d, plm_mask, static_cast<cudaStream_t>(stream));
}
