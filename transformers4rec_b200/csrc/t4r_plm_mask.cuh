// t4r_plm_mask.cuh -- labels and the permutation attention mask of XLNet Permutation Language Modeling for ONE session
// (transformers4rec/torch/masking.py:548-727, PermutationLanguageModeling._compute_masked_targets_extended), shared by
// the device kernel and its host twin (t4r_plm.cu).  Integer work, bit-exact against the upstream code for identical
// draws; the draw protocol (per session: u_span / u_start per loop iteration, u_force, u_unmask, the factorisation
// order perm) is the one of oracle/t4r_oracle.py::plm_compute_masked_targets.  target_mapping is the identity in every
// mode the reference produces it in (training spans, both evaluation modes), so it is not materialised.
#pragma once
#include <stdint.h>

namespace t4r {

constexpr int kPlmMaxL = 64;
constexpr int kPlmMaxSpan = 15;

struct PlmParams {
  int L;
  int mode;                       // T4R_PLM_TRAIN / EVAL_LAST / EVAL_ALL
  int max_span;                   // <= kPlmMaxSpan
  int ctx_len[kPlmMaxSpan + 1];   // int(span / plm_probability), computed by the host exactly like the reference
  int64_t padding_idx;
};

// min(floor(u * n), n - 1) in double: torch.randint / "k-th set position" stand-in (oracle: randint_from_uniform)
__host__ __device__ inline int plm_rfu(float u, int n) {
  int k = static_cast<int>(static_cast<double>(u) * static_cast<double>(n));  // u >= 0: truncation == floor
  return k < n - 1 ? k : n - 1;
}
__host__ __device__ inline int plm_pick_kth(const uint8_t* set, int L, float u) {
  int n = 0;
  for (int j = 0; j < L; ++j) n += set[j] ? 1 : 0;
  if (n == 0) return 0;
  int k = plm_rfu(u, n);
  for (int j = 0; j < L; ++j)
    if (set[j] && k-- == 0) return j;
  return 0;
}

__host__ __device__ inline void plm_mask_session(const PlmParams& p, const int64_t* ids, const float* u_span,
                                                 const float* u_start, float u_force, float u_unmask, const int32_t* perm,
                                                 uint8_t* mask_out, int64_t* labels_out, uint8_t* perm_mask_out) {
  const int L = p.L;
  uint8_t m[kPlmMaxL], nonpad[kPlmMaxL];
  int n_items = 0;
  for (int j = 0; j < L; ++j) {
    nonpad[j] = ids[j] != p.padding_idx;
    n_items += nonpad[j];
    m[j] = 0;
  }
  if (p.mode == 0) {
    // masking.py:598-627: spans of consecutive items, each inside its own context window
    int cur = 0, n = 0;
    while (cur < n_items && n < L) {
      const int span = 1 + plm_rfu(u_span[n], p.max_span);
      const int ctx = p.ctx_len[span];
      const int start = cur + plm_rfu(u_start[n], ctx - span + 1);
      if (start < n_items)
        for (int k = start; k < start + span && k < L; ++k) m[k] = 1;
      cur += ctx;
      ++n;
    }
    int cnt = 0;
    for (int j = 0; j < L; ++j) cnt += m[j];
    if (cnt == 0) {  // :629-638 at least one item to predict
      const int k = plm_pick_kth(nonpad, L, u_force);
      m[k] = ids[k] != 0;  // the reference assigns the item id into a bool tensor
    }
    for (int j = 0; j < L; ++j) labels_out[j] = m[j] ? ids[j] : p.padding_idx;
    // :646-657 a session with only labels gets one of them removed (the count includes masked PADDED positions)
    cnt = 0;
    for (int j = 0; j < L; ++j) cnt += m[j];
    const int sampled = plm_pick_kth(m, L, u_unmask);
    if (cnt == n_items) labels_out[sampled] = p.padding_idx;
    for (int j = 0; j < L; ++j) m[j] = labels_out[j] != p.padding_idx;
    // :659-683 factorisation order: non-labels get index -1 (seen by everyone, see no label)
    for (int i = 0; i < L; ++i) {
      const int pi = m[i] ? perm[i] : -1;
      for (int j = 0; j < L; ++j) {
        const int pj = m[j] ? perm[j] : -1;
        perm_mask_out[i * L + j] = (pi <= pj) && m[j];
      }
    }
  } else if (p.mode == 1) {
    // :686-703 evaluation on the last item (python's negative index for an empty session: column L - 1)
    for (int j = 0; j < L; ++j) labels_out[j] = p.padding_idx;
    const int last = n_items > 0 ? n_items - 1 : L - 1;
    labels_out[last] = ids[last];
    for (int j = 0; j < L; ++j) m[j] = labels_out[j] != p.padding_idx;
    for (int i = 0; i < L; ++i)
      for (int j = 0; j < L; ++j) perm_mask_out[i * L + j] = (j > i) || (j == last);
  } else {
    // :705-725 predict all next items, causal mask
    for (int j = 0; j < L; ++j) {
      labels_out[j] = (j + 1 < L) ? ids[j + 1] : 0;
      m[j] = labels_out[j] != p.padding_idx;
    }
    for (int i = 0; i < L; ++i)
      for (int j = 0; j < L; ++j) perm_mask_out[i * L + j] = (j > i);
  }
  for (int j = 0; j < L; ++j) mask_out[j] = m[j];
}

}  // namespace t4r
