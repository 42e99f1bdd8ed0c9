// Row-level arithmetic of the GFLv2 / LDv2 additions, written as host+device inline functions so
// that the SAME code is (a) called one-thread-per-row by the kernels in dgqp.cu and (b) compiled for the host by
// tests/host_shim/v2_math_host.cpp and checked against torch autograd of the reference formulas on CPU
// (tests/test_v2_math_host.py).
//
//   * Distribution-Guided Quality Predictor, `GFocalHead.forward_single` (mmdet/models/dense_heads/gfocal_head.py:
//     198-215): per side softmax over the bins, top-`reg_topk` probabilities (+ their mean, `add_mean`), the
//     4*(topk+1) statistics -> 1x1 conv -> ReLU -> 1x1 conv -> sigmoid = quality; joint score = sigmoid(cls_feat) * q.
//   * Quality focal loss on probabilities, `quality_focal_loss(use_sigmoid=False)` (losses/gfocal_loss.py:29-49):
//     F.binary_cross_entropy (log clamped at -100, as ATen does) times |target - p|^beta.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define LDB_HD __host__ __device__ __forceinline__
#else
#define LDB_HD inline
#endif

namespace ldb {

constexpr int DGQP_MAX_BINS = 32;
constexpr int DGQP_MAX_TOPK = 8;
constexpr int DGQP_MAX_STAT = 4 * (DGQP_MAX_TOPK + 1);
constexpr int DGQP_MAX_HID = 64;

struct DgqpShape {
  int bins;      // reg_max + 1
  int topk;      // reg_topk
  int add_mean;  // 0 / 1
  int hid;       // reg_channels
  LDB_HD int td() const { return topk + (add_mean ? 1 : 0); }
  LDB_HD int nstat() const { return 4 * td(); }
};

// softmax of one side + its `topk` largest probabilities in descending order (indices kept for the backward pass)
LDB_HD void dgqp_side(const float* z, int bins, int topk, float* p, float* top_val, int* top_idx) {
  float m = z[0];
  for (int i = 1; i < bins; ++i) m = fmaxf(m, z[i]);
  float S = 0.f;
  for (int i = 0; i < bins; ++i) {
    p[i] = expf(z[i] - m);
    S += p[i];
  }
  const float inv = 1.f / S;
  for (int i = 0; i < bins; ++i) p[i] *= inv;
  unsigned taken = 0;
  for (int j = 0; j < topk; ++j) {
    int bi = -1;
    float bv = -1.f;
    for (int i = 0; i < bins; ++i)
      if (!((taken >> i) & 1u) && p[i] > bv) { bv = p[i]; bi = i; }
    taken |= 1u << bi;
    top_val[j] = bv;
    top_idx[j] = bi;
  }
}

// statistics vector of a row: channel = side * td + j (gfocal_head.py:203-213: stat.reshape(N, -1, H, W))
LDB_HD void dgqp_stats(const float* reg, const DgqpShape& s, float* stat, int* idx /*[4*topk]*/, float* prob /*[4*bins] or null*/) {
  const int td = s.td();
  for (int side = 0; side < 4; ++side) {
    float p[DGQP_MAX_BINS], tv[DGQP_MAX_TOPK];
    int ti[DGQP_MAX_TOPK];
    dgqp_side(reg + side * s.bins, s.bins, s.topk, p, tv, ti);
    float mean = 0.f;
    for (int j = 0; j < s.topk; ++j) {
      stat[side * td + j] = tv[j];
      idx[side * s.topk + j] = ti[j];
      mean += tv[j];
    }
    if (s.add_mean) stat[side * td + s.topk] = mean / (float)s.topk;
    if (prob)
      for (int i = 0; i < s.bins; ++i) prob[side * s.bins + i] = p[i];
  }
}

// quality = sigmoid(W2 . relu(W1 stat + b1) + b2).  W1 [hid][nstat], W2 [hid].  hpre (pre-ReLU) is returned for backward.
LDB_HD float dgqp_mlp(const float* stat, const DgqpShape& s, const float* W1, const float* b1, const float* W2, float b2,
                      float* hpre /*[hid] or null*/) {
  const int ns = s.nstat();
  float o = b2;
  for (int j = 0; j < s.hid; ++j) {
    float a = b1[j];
    for (int i = 0; i < ns; ++i) a += W1[j * ns + i] * stat[i];
    if (hpre) hpre[j] = a;
    o += W2[j] * fmaxf(a, 0.f);
  }
  return 1.f / (1.f + expf(-o));
}

LDB_HD float dgqp_row_forward(const float* reg, const DgqpShape& s, const float* W1, const float* b1, const float* W2,
                              float b2) {
  float stat[DGQP_MAX_STAT];
  int idx[4 * DGQP_MAX_TOPK];
  dgqp_stats(reg, s, stat, idx, nullptr);
  return dgqp_mlp(stat, s, W1, b1, W2, b2, nullptr);
}

// Backward of one row.  In: dq = d(loss)/d(quality).  Out: d_reg[4*bins] (ADDED to), dhpre[hid] (d loss / d pre-ReLU
// hidden; the caller reduces dW1 += dhpre x stat, db1 += dhpre over rows), stat[nstat], h[hid] (post-ReLU, for
// dW2 += d_o * h) and the return value d_o = d(loss)/d(pre-sigmoid output) (db2 += d_o).
LDB_HD float dgqp_row_backward(const float* reg, const DgqpShape& s, const float* W1, const float* b1, const float* W2,
                               float b2, float dq, float* d_reg, float* dhpre, float* stat, float* h) {
  const int ns = s.nstat(), td = s.td();
  int idx[4 * DGQP_MAX_TOPK];
  float prob[4 * DGQP_MAX_BINS];
  dgqp_stats(reg, s, stat, idx, prob);
  float hpre[DGQP_MAX_HID];
  const float q = dgqp_mlp(stat, s, W1, b1, W2, b2, hpre);
  const float d_o = dq * q * (1.f - q);
  float dstat[DGQP_MAX_STAT];
  for (int i = 0; i < ns; ++i) dstat[i] = 0.f;
  for (int j = 0; j < s.hid; ++j) {
    h[j] = fmaxf(hpre[j], 0.f);
    const float g = hpre[j] > 0.f ? d_o * W2[j] : 0.f;
    dhpre[j] = g;
    if (g != 0.f)
      for (int i = 0; i < ns; ++i) dstat[i] += W1[j * ns + i] * g;
  }
  for (int side = 0; side < 4; ++side) {
    float dp[DGQP_MAX_BINS];
    for (int i = 0; i < s.bins; ++i) dp[i] = 0.f;
    const float dmean = s.add_mean ? dstat[side * td + s.topk] / (float)s.topk : 0.f;
    for (int j = 0; j < s.topk; ++j) dp[idx[side * s.topk + j]] += dstat[side * td + j] + dmean;
    const float* p = prob + side * s.bins;
    float dot = 0.f;
    for (int i = 0; i < s.bins; ++i) dot += p[i] * dp[i];
    for (int i = 0; i < s.bins; ++i) d_reg[side * s.bins + i] += p[i] * (dp[i] - dot);
  }
  return d_o;
}

// ---- quality focal loss on probabilities -------------------------------------------------------------------
LDB_HD float qflp_pow(float a, float beta) { return beta == 2.f ? a * a : powf(a, beta); }
LDB_HD float qflp_pow1(float a, float beta) { return beta == 2.f ? a : powf(a, beta - 1.f); }

// element supervised by 0 (gfocal_loss.py:36-39): BCE(p, 0) * p^beta and d/dp
LDB_HD void qflp_neg(float p, float beta, float& loss, float& grad) {
  const float bce = -fmaxf(logf(1.f - p), -100.f);
  const float dbce = p / fmaxf((1.f - p) * p, 1e-12f);     // ATen binary_cross_entropy_backward, target 0
  loss = bce * qflp_pow(p, beta);
  grad = dbce * qflp_pow(p, beta) + bce * beta * qflp_pow1(p, beta);
}
// positive class with quality target y (gfocal_loss.py:44-47): BCE(p, y) * |y - p|^beta and d/dp
LDB_HD void qflp_pos(float p, float y, float beta, float& loss, float& grad) {
  const float bce = -(y * fmaxf(logf(p), -100.f) + (1.f - y) * fmaxf(logf(1.f - p), -100.f));
  const float dbce = (p - y) / fmaxf((1.f - p) * p, 1e-12f);
  const float d = y - p, ad = fabsf(d);
  const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  loss = bce * qflp_pow(ad, beta);
  grad = dbce * qflp_pow(ad, beta) - bce * beta * qflp_pow1(ad, beta) * sgn;
}

}  // namespace ldb
