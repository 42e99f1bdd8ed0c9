// Fused LDHead loss (forward + gradient in one launch) and the dense LD-KL micro-benchmark kernel.
//
// Replaces (reference, paths under /root/reference/mmdet):
//   models/dense_heads/ld_head.py:116-282   LDHead.loss_single (x5 levels, ~150 launches + 4 host
//                                           syncs per level in eager PyTorch)
//   models/losses/gfocal_loss.py:8-74       quality_focal_loss / distribution_focal_loss
//   models/losses/iou_loss.py:85-102        giou_loss  (+ core/bbox/iou_calculators/iou2d_calculator.py)
//   models/losses/kd_loss.py:10-36,91-96    knowledge_distillation_kl_div_loss / im_loss
//   models/losses/utils.py:27-100           weight_reduce_loss
//   core/bbox/transforms.py:119-180, models/dense_heads/gfl_head.py:32-44  distance2bbox,
//                                           bbox2distance, Integral
// and their autograd backward.  HBM-bound: the only dense traffic is one read of the class logits
// and one write of their gradient; everything else runs on the compacted positive / VLR / fine-
// grained row lists produced by targets.cu (one warp per row, warp-shuffle softmaxes).
#include <stdlib.h>

#include "common.cuh"
#include "v2_math.cuh"

namespace ldb {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_WARPS = LOSS_THREADS / 32;
constexpr int MAX_CPL = 4;  // class logits per lane: C <= 128

struct LossParams {
  Pyr pyr;
  int C, bins, F;
  float qfl_w, beta, giou_w, eps, dfl_w, ld_w, ld_T, ldv_w, ldv_T, kd_w, kd_T, im_w;
  const float *cls, *reg, *t_cls, *t_reg;
  const void *x, *t_x;
  int feat_bf16;
  const int* labels;
  const float *label_w, *bbox_t, *vlr;
  const int *sparse_rows, *fg_rows, *counts;
  const float* norm;
  float tg[LDB_NUM_TERMS - 1][LDB_MAX_LEVELS];
  int want_grad;
  int phases;   // bit0 dense QFL, bit1 sparse rows, bit2 imitation rows (all set normally)
  float* losses;
  void *d_cls, *d_reg, *d_im;
  int dcls_ld, dreg_ld;   // row strides (elements) of d_cls / d_reg
  int grad_bf16;
  float* score;  // prepass only (may be null)
  float* wsum;   // prepass only
  // LDv2Head (ld_gflv2.py; V2 kernels only -- appended so that the layout seen by the V1 kernels is unchanged):
  // `cls` then holds the joint scores sigmoid(cls_feat) * quality, `t_cls` the teacher's cls_feat logits
  const float* cls_feat;   // [R, C] student cls_feat logits (KD term, ld_gflv2.py:242-246)
  void* d_kd;              // [R, dcls_ld] gradient of the KD term w.r.t. cls_feat (zero-filled by the call)
  int num_fg;              // foreground classes (bg_class_ind, ld_head.py:163); C = num_fg + 1 in the LDv2 configs
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float softplusf_(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

// Dense-phase variant of qfl_neg: one ex2 + one lg2 + one fast reciprocal per logit (relative error
// ~1e-6, far inside the 1e-4 parity tolerance; the sparse positive rows keep the precise path).
__device__ __forceinline__ void qfl_neg_fast(float x, float beta, float& loss, float& grad) {
  const float e = __expf(-fabsf(x));
  const float r = __fdividef(1.f, 1.f + e);
  const float s = x >= 0.f ? r : e * r;
  const float sp = fmaxf(x, 0.f) + __logf(1.f + e);
  const float sb = (beta == 2.f) ? s * s : __powf(s, beta);
  loss = sp * sb;
  grad = sb * (s + sp * beta * (1.f - s));
}
// QFL element, negative class (gfocal_loss.py:34-39): sp(x) * sigma^beta and d/dx.
__device__ __forceinline__ void qfl_neg(float x, float beta, float& loss, float& grad) {
  float s = sigmoidf_(x), sp = softplusf_(x);
  float sb = (beta == 2.f) ? s * s : powf(s, beta);
  loss = sp * sb;
  grad = s * sb + sp * beta * sb * (1.f - s);
}
// QFL element, positive class with quality target y (gfocal_loss.py:41-47).
__device__ __forceinline__ void qfl_pos(float x, float y, float beta, float& loss, float& grad) {
  float s = sigmoidf_(x), bce = softplusf_(x) - x * y;
  float d = y - s, ad = fabsf(d);
  float pb = (beta == 2.f) ? d * d : powf(ad, beta);
  float pb1 = (beta == 2.f) ? ad : powf(ad, beta - 1.f);
  loss = bce * pb;
  float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  grad = (s - y) * pb + bce * beta * pb1 * sgn * (-s * (1.f - s));
}

template <typename T> __device__ __forceinline__ void store_grad(T* p, float v);
template <> __device__ __forceinline__ void store_grad<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void store_grad<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

struct RowGeom {
  int l, n, ix, iy;
};
__device__ __forceinline__ RowGeom row_geom(const Pyr& p, int r) {
  RowGeom g;
  g.l = level_of_row(p, r);
  int q = r - p.off[g.l] * p.N;
  int Al = p.h[g.l] * p.w[g.l];
  g.n = q / Al;
  int a = q - g.n * Al;
  g.iy = a / p.w[g.l];
  g.ix = a - g.iy * p.w[g.l];
  return g;
}

// softmax over the first `bins` lanes of a warp; returns p and log p for this lane.
__device__ __forceinline__ void warp_softmax(float z, bool active, float& p, float& logp) {
  float zz = active ? z : -INFINITY;
  float m = warp_max(zz);
  float e = active ? expf(zz - m) : 0.f;
  float S = warp_sum(e);
  p = e / S;
  logp = zz - m - logf(S);
}

// 1 - GIoU (iou_loss.py:85-102 via iou2d_calculator.py:111-188, aligned) with analytic gradient
// w.r.t. the predicted box (checked against autograd through the reference).
__device__ __forceinline__ void giou_fwd_bwd(const float p[4], const float t[4], float eps, float& iou,
                                             float& loss, float g[4]) {
  float pw = p[2] - p[0], ph = p[3] - p[1];
  float ap = pw * ph, at = (t[2] - t[0]) * (t[3] - t[1]);
  float iwr = fminf(p[2], t[2]) - fmaxf(p[0], t[0]);
  float ihr = fminf(p[3], t[3]) - fmaxf(p[1], t[1]);
  float iw = fmaxf(iwr, 0.f), ih = fmaxf(ihr, 0.f);
  float I = iw * ih;
  float Ur = ap + at - I;
  float U = fmaxf(Ur, eps);
  iou = I / U;
  float ewr = fmaxf(p[2], t[2]) - fminf(p[0], t[0]);
  float ehr = fmaxf(p[3], t[3]) - fminf(p[1], t[1]);
  float ew = fmaxf(ewr, 0.f), eh = fmaxf(ehr, 0.f);
  float Er = ew * eh;
  float E = fmaxf(Er, eps);
  loss = 1.f - (iou - (E - U) / E);
  float gUr = (Ur > eps) ? (I / (U * U) - 1.f / E) : 0.f;
  float gI = -1.f / U - gUr;
  float gEr = (Er > eps) ? U / (E * E) : 0.f;
  float gew = (ewr >= 0.f) ? gEr * eh : 0.f, geh = (ehr >= 0.f) ? gEr * ew : 0.f;
  float giw = (iwr >= 0.f) ? gI * ih : 0.f, gih = (ihr >= 0.f) ? gI * iw : 0.f;
  g[0] = -gUr * ph - (p[0] > t[0] ? giw : 0.f) - (p[0] < t[0] ? gew : 0.f);
  g[2] = gUr * ph + (p[2] < t[2] ? giw : 0.f) + (p[2] > t[2] ? gew : 0.f);
  g[1] = -gUr * pw - (p[1] > t[1] ? gih : 0.f) - (p[1] < t[1] ? geh : 0.f);
  g[3] = gUr * pw + (p[3] < t[3] ? gih : 0.f) + (p[3] > t[3] ? geh : 0.f);
}

// decode the student's box for a positive row: Integral (gfl_head.py:42-43) + distance2bbox
// (transforms.py:119-156) in grid units.  All lanes return the same values.
struct BoxBranch {
  float p[4], lse[4], d[4];  // p = T=1 softmax prob of this lane's bin per side, lse per side
  float pred[4], tgt[4];
};
__device__ __forceinline__ void decode_row(const float z[4], bool active, int lane, const RowGeom& g,
                                           const float* __restrict__ bbox_t, int r, float stride,
                                           BoxBranch& b) {
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    float logp;
    warp_softmax(z[s], active, b.p[s], logp);
    b.lse[s] = z[s] - logp;  // valid on active lanes; broadcast below
    b.d[s] = warp_sum(active ? (float)lane * b.p[s] : 0.f);
    b.lse[s] = __shfl_sync(FULL, b.lse[s], 0);
  }
  const float cx = (float)g.ix, cy = (float)g.iy;  // anchor centre / stride (ld_head.py:196)
  b.pred[0] = cx - b.d[0]; b.pred[1] = cy - b.d[1];
  b.pred[2] = cx + b.d[2]; b.pred[3] = cy + b.d[3];
  const float4 bt = reinterpret_cast<const float4*>(bbox_t)[r];
  b.tgt[0] = bt.x / stride; b.tgt[1] = bt.y / stride; b.tgt[2] = bt.z / stride; b.tgt[3] = bt.w / stride;
}

__device__ __forceinline__ float row_weight_target(const float* __restrict__ cls_row, int C, int lane) {
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, cls_row[c]);
  return sigmoidf_(warp_max(m));  // max_c sigmoid(x_c) (ld_head.py:198-199), sigmoid is monotone
}
// LDv2: weight_targets = max_c joint score, no sigmoid (ld_gflv2.py:200-201)
__device__ __forceinline__ float row_weight_target_prob(const float* __restrict__ score_row, int C, int lane) {
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, score_row[c]);
  return warp_max(m);
}

// ---------------------------------------------------------------------------------------------
// Pre-pass: sum of weight_targets (and optional IoU score map) over the positive rows.
template <bool V2>
__global__ void __launch_bounds__(LOSS_THREADS) ld_prepass_kernel(LossParams P) {
  __shared__ float s_w[LOSS_WARPS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n_sparse = P.counts[0];
  float acc = 0.f;
  for (int i = blockIdx.x * LOSS_WARPS + wid; i < n_sparse; i += gridDim.x * LOSS_WARPS) {
    const int r = P.sparse_rows[i];
    const int lab = P.labels[r];
    if (lab < 0 || lab >= (V2 ? P.num_fg : P.C)) continue;
    const RowGeom g = row_geom(P.pyr, r);
    const float w = V2 ? row_weight_target_prob(P.cls + (size_t)r * P.C, P.C, lane)
                       : row_weight_target(P.cls + (size_t)r * P.C, P.C, lane);
    acc += w;
    if (P.score) {
      const bool active = lane < P.bins;
      float z[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) z[s] = active ? P.reg[(size_t)r * 4 * P.bins + s * P.bins + lane] : 0.f;
      BoxBranch b;
      decode_row(z, active, lane, g, P.bbox_t, r, (float)P.pyr.stride[g.l], b);
      float iou, loss, gg[4];
      giou_fwd_bwd(b.pred, b.tgt, P.eps, iou, loss, gg);
      if (lane == 0) P.score[r] = iou;
    }
  }
  if (lane == 0) s_w[wid] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < LOSS_WARPS; ++i) t += s_w[i];
    if (t != 0.f) atomicAdd(P.wsum, t);
  }
}

// ---------------------------------------------------------------------------------------------
// The fused loss kernel.
template <typename GT, typename FT, bool V2>
__global__ void __launch_bounds__(LOSS_THREADS, 3) ld_loss_fused_kernel(LossParams P) {
  __shared__ float s_loss[LDB_NUM_TERMS][LDB_MAX_LEVELS];
  __shared__ float s_red[LOSS_WARPS];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < LDB_NUM_TERMS * LDB_MAX_LEVELS; i += LOSS_THREADS) (&s_loss[0][0])[i] = 0.f;
  __syncthreads();
  const int C = P.C, bins = P.bins, N = P.pyr.N;
  const int nfg = V2 ? P.num_fg : C;   // labels in [0, nfg) are positives (ld_head.py:163-165)
  const float nts = P.norm[0], avg = P.norm[1];
  const bool wg = P.want_grad != 0;
  GT* __restrict__ d_cls = (GT*)P.d_cls;
  GT* __restrict__ d_reg = (GT*)P.d_reg;

  // ---------------- phase 1: dense QFL over every non-positive row (gfocal_loss.py:34-49) -----
  const int vpr = C >> 2;  // float4 per row (C % 4 == 0 fast path), else scalar path below
  // block roles: two of every three blocks stream the dense QFL phase, the third walks the compacted rows, so the
  // latency-bound sparse phases overlap the bandwidth-bound dense one instead of following it
  const int role_sparse = (blockIdx.x % 3) == 2;
  const int dense_blocks = gridDim.x - gridDim.x / 3, sparse_blocks = gridDim.x / 3;
  const int dense_id = blockIdx.x - blockIdx.x / 3, sparse_id = blockIdx.x / 3;
  for (int l = 0; l < (((P.phases & 1) && !role_sparse) ? P.pyr.L : 0); ++l) {
    const long long r0 = (long long)P.pyr.off[l] * N, r1 = (long long)P.pyr.off[l + 1] * N;
    const float gsc = P.tg[0][l];
    float acc = 0.f;
    if ((C & 3) == 0) {
      const long long v0 = r0 * vpr, v1 = r1 * vpr;
      for (long long v = v0 + (long long)dense_id * LOSS_THREADS + tid; v < v1;
           v += (long long)dense_blocks * LOSS_THREADS) {
        const int r = (int)(v / vpr);
        const int lab = __ldg(P.labels + r);
        if (lab >= 0 && lab < nfg) continue;  // positive rows are owned by phase 2
        const float sc = P.qfl_w * __ldg(P.label_w + r) / nts;
        const float4 xv = __ldg(reinterpret_cast<const float4*>(P.cls) + v);
        float lo[4], gr[4];
        if (V2) {   // joint scores: BCE on probabilities (gfocal_loss.py:29-39, use_sigmoid=False)
          qflp_neg(xv.x, P.beta, lo[0], gr[0]);
          qflp_neg(xv.y, P.beta, lo[1], gr[1]);
          qflp_neg(xv.z, P.beta, lo[2], gr[2]);
          qflp_neg(xv.w, P.beta, lo[3], gr[3]);
        } else {
          if (sizeof(GT) == 4) {   // fp32 gradients (parity mode / loss API): precise exp / log1p, no approximations
            qfl_neg(xv.x, P.beta, lo[0], gr[0]);
            qfl_neg(xv.y, P.beta, lo[1], gr[1]);
            qfl_neg(xv.z, P.beta, lo[2], gr[2]);
            qfl_neg(xv.w, P.beta, lo[3], gr[3]);
          } else {
            qfl_neg_fast(xv.x, P.beta, lo[0], gr[0]);
            qfl_neg_fast(xv.y, P.beta, lo[1], gr[1]);
            qfl_neg_fast(xv.z, P.beta, lo[2], gr[2]);
            qfl_neg_fast(xv.w, P.beta, lo[3], gr[3]);
          }
        }
        acc += sc * ((lo[0] + lo[1]) + (lo[2] + lo[3]));
        if (wg) {
          const float k = sc * gsc;
          const size_t go = (size_t)r * P.dcls_ld + (size_t)(v - (long long)r * vpr) * 4;
          if (sizeof(GT) == 4) {
            *reinterpret_cast<float4*>((float*)d_cls + go) = make_float4(k * gr[0], k * gr[1], k * gr[2], k * gr[3]);
          } else {
            __nv_bfloat162 a = __floats2bfloat162_rn(k * gr[0], k * gr[1]);
            __nv_bfloat162 b = __floats2bfloat162_rn(k * gr[2], k * gr[3]);
            uint2 pk;
            pk.x = *reinterpret_cast<unsigned*>(&a);
            pk.y = *reinterpret_cast<unsigned*>(&b);
            *reinterpret_cast<uint2*>((__nv_bfloat16*)d_cls + go) = pk;
          }
        }
      }
    } else {
      const long long e0 = r0 * C, e1 = r1 * C;
      for (long long e = e0 + (long long)dense_id * LOSS_THREADS + tid; e < e1;
           e += (long long)dense_blocks * LOSS_THREADS) {
        const int r = (int)(e / C);
        const int lab = __ldg(P.labels + r);
        if (lab >= 0 && lab < nfg) continue;
        const float sc = P.qfl_w * __ldg(P.label_w + r) / nts;
        float lo, gr;
        if (V2) qflp_neg(__ldg(P.cls + e), P.beta, lo, gr);
        else qfl_neg(__ldg(P.cls + e), P.beta, lo, gr);
        acc += sc * lo;
        if (wg) store_grad<GT>(d_cls + (size_t)r * P.dcls_ld + (e - (long long)r * C), sc * gsc * gr);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0 && acc != 0.f) atomicAdd(&s_loss[0][l], acc);
  }

  // ---------------- phase 2: compacted positive / VLR rows, one warp per row ------------------
  const int n_sparse = ((P.phases & 2) && role_sparse) ? P.counts[0] : 0;
  const bool active = lane < bins;
  for (int i = sparse_id * LOSS_WARPS + wid; i < n_sparse; i += sparse_blocks * LOSS_WARPS) {
    const int r = P.sparse_rows[i];
    const RowGeom g = row_geom(P.pyr, r);
    const int l = g.l;
    const int lab = P.labels[r];
    const bool pos = lab >= 0 && lab < nfg;
    const float v = P.vlr[r];
    const float* reg_row = P.reg + (size_t)r * 4 * bins;
    const float* treg_row = P.t_reg + (size_t)r * 4 * bins;
    float z[4], tz[4], gz[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      z[s] = active ? reg_row[s * bins + lane] : 0.f;
      tz[s] = active ? treg_row[s * bins + lane] : 0.f;
    }
    float l_bbox = 0.f, l_dfl = 0.f, l_ld = 0.f, l_ldv = 0.f, l_kd = 0.f, l_cls = 0.f, w = 0.f;
    if (pos) {
      const float* cls_row = P.cls + (size_t)r * C;
      const float* tcls_row = P.t_cls + (size_t)r * C;
      w = V2 ? row_weight_target_prob(cls_row, C, lane) : row_weight_target(cls_row, C, lane);
      BoxBranch b;
      decode_row(z, active, lane, g, P.bbox_t, r, (float)P.pyr.stride[l], b);
      float iou, lg, gb[4];
      giou_fwd_bwd(b.pred, b.tgt, P.eps, iou, lg, gb);
      // GIoU: loss_weight * w * (1-giou) / 1.0, later / avg_factor (ld_head.py:222-226,364)
      const float k_bbox = P.giou_w * w / avg;
      l_bbox = k_bbox * lg;
      const float gd[4] = {-gb[0], -gb[1], gb[2], gb[3]};
      // DFL (gfocal_loss.py:68-74) on target corners (transforms.py:159-180, eps = 0.1)
      const float cx = (float)g.ix, cy = (float)g.iy;
      const float corner[4] = {cx - b.tgt[0], cy - b.tgt[1], b.tgt[2] - cx, b.tgt[3] - cy};
      const float k_dfl = P.dfl_w * w / 4.0f / avg;
      const float k_ld = P.ld_w * w / 4.0f;
      const float invT = 1.f / P.ld_T;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const float y = fminf(fmaxf(corner[s], 0.f), (float)(bins - 1) - 0.1f);
        const int il = (int)y;
        const float wl = (float)(il + 1) - y, wr = y - (float)il;
        const float zl = __shfl_sync(FULL, z[s], il), zr = __shfl_sync(FULL, z[s], il + 1);
        l_dfl += k_dfl * ((b.lse[s] - zl) * wl + (b.lse[s] - zr) * wr);
        float gdfl = b.p[s] * (wl + wr) - (lane == il ? wl : 0.f) - (lane == il + 1 ? wr : 0.f);
        // LD: KL(softmax(t/T) || softmax(s/T)) * T^2, mean over bins (kd_loss.py:28-34)
        float ps, lps, pt, lpt;
        warp_softmax(z[s] * invT, active, ps, lps);
        warp_softmax(tz[s] * invT, active, pt, lpt);
        float kl = warp_sum((active && pt > 0.f) ? pt * (lpt - lps) : 0.f);
        l_ld += k_ld * kl * (P.ld_T * P.ld_T) / (float)bins;
        if (active) {
          gz[s] += P.tg[1][l] * k_bbox * gd[s] * b.p[s] * ((float)lane - b.d[s]) +
                   P.tg[2][l] * k_dfl * gdfl +
                   P.tg[3][l] * k_ld * (P.ld_T / (float)bins) * (ps - pt);
        }
      }
      // KD on the class logits (ld_head.py:240-244): weight = label_weights, avg = P_level
      const int P_l = P.counts[4 + 3 * l + 0];
      const float k_kd = P.kd_w * P.label_w[r] / (float)P_l;
      const float invTk = 1.f / P.kd_T;
      // V2: the KD logits are the student's cls_feat (ld_gflv2.py:243), QFL keeps reading the joint scores
      const float* kd_row = V2 ? P.cls_feat + (size_t)r * C : cls_row;
      float xs[MAX_CPL], xt[MAX_CPL];
      float ms = -INFINITY, mt = -INFINITY;
#pragma unroll
      for (int q = 0; q < MAX_CPL; ++q) {
        const int c = lane + 32 * q;
        xs[q] = c < C ? kd_row[c] : -INFINITY;
        xt[q] = c < C ? tcls_row[c] : -INFINITY;
        ms = fmaxf(ms, xs[q] * invTk);
        mt = fmaxf(mt, xt[q] * invTk);
      }
      ms = warp_max(ms);
      mt = warp_max(mt);
      float Ss = 0.f, St = 0.f;
#pragma unroll
      for (int q = 0; q < MAX_CPL; ++q) {
        const int c = lane + 32 * q;
        if (c < C) {
          Ss += expf(xs[q] * invTk - ms);
          St += expf(xt[q] * invTk - mt);
        }
      }
      Ss = warp_sum(Ss);
      St = warp_sum(St);
      const float lSs = logf(Ss), lSt = logf(St);
      const float y = iou;  // quality target = IoU(pred.detach(), target) (ld_head.py:204-207)
      const float k_cls = P.qfl_w * P.label_w[r] / nts;
      float kl = 0.f, qsum = 0.f;
#pragma unroll
      for (int q = 0; q < MAX_CPL; ++q) {
        const int c = lane + 32 * q;
        if (c < C) {
          const float lps = xs[q] * invTk - ms - lSs, lpt = xt[q] * invTk - mt - lSt;
          const float ps = expf(lps), pt = expf(lpt);
          if (pt > 0.f) kl += pt * (lpt - lps);
          float lo, gr;
          if (V2) {
            const float pj = cls_row[c];
            if (c == lab) qflp_pos(pj, y, P.beta, lo, gr);
            else qflp_neg(pj, P.beta, lo, gr);
          } else {
            if (c == lab) qfl_pos(xs[q], y, P.beta, lo, gr);
            else qfl_neg(xs[q], P.beta, lo, gr);
          }
          qsum += lo;
          if (wg) {
            if (V2) {
              store_grad<GT>(d_cls + (size_t)r * P.dcls_ld + c, P.tg[0][l] * k_cls * gr);
              store_grad<GT>((GT*)P.d_kd + (size_t)r * P.dcls_ld + c,
                             P.tg[5][l] * k_kd * (P.kd_T / (float)C) * (ps - pt));
            } else {
              store_grad<GT>(d_cls + (size_t)r * P.dcls_ld + c,
                             P.tg[0][l] * k_cls * gr + P.tg[5][l] * k_kd * (P.kd_T / (float)C) * (ps - pt));
            }
          }
        }
      }
      l_kd = k_kd * warp_sum(kl) * (P.kd_T * P.kd_T) / (float)C;
      l_cls = k_cls * warp_sum(qsum);
    }
    if (v > 0.f) {
      // LD on the Valuable Localization Region (ld_head.py:254-266): weight = vlr IoU, avg 16
      const float k_ldv = P.ldv_w * v / 16.0f;
      const float invT = 1.f / P.ldv_T;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        float ps, lps, pt, lpt;
        warp_softmax(z[s] * invT, active, ps, lps);
        warp_softmax(tz[s] * invT, active, pt, lpt);
        float kl = warp_sum((active && pt > 0.f) ? pt * (lpt - lps) : 0.f);
        l_ldv += k_ldv * kl * (P.ldv_T * P.ldv_T) / (float)bins;
        if (active) gz[s] += P.tg[4][l] * k_ldv * (P.ldv_T / (float)bins) * (ps - pt);
      }
    }
    if (wg && active) {
#pragma unroll
      for (int s = 0; s < 4; ++s) store_grad<GT>(d_reg + (size_t)r * P.dreg_ld + s * bins + lane, gz[s]);
    }
    if (lane == 0) {
      if (pos) {
        atomicAdd(&s_loss[0][l], l_cls);
        atomicAdd(&s_loss[1][l], l_bbox);
        atomicAdd(&s_loss[2][l], l_dfl);
        atomicAdd(&s_loss[3][l], l_ld);
        atomicAdd(&s_loss[5][l], l_kd);
        atomicAdd(&s_loss[8][l], w);
      }
      if (v > 0.f) atomicAdd(&s_loss[4][l], l_ldv);
    }
  }

  // ---------------- phase 3: feature imitation on the fine-grained rows (kd_loss.py:91-96) ----
  const int n_fg = ((P.phases & 4) && role_sparse) ? P.counts[1] : 0;
  const int F = P.F;
  const FT* __restrict__ xs_ = (const FT*)P.x;
  const FT* __restrict__ xt_ = (const FT*)P.t_x;
  GT* __restrict__ d_im = (GT*)P.d_im;
  for (int i = sparse_id * LOSS_WARPS + wid; i < n_fg; i += sparse_blocks * LOSS_WARPS) {
    const int r = P.fg_rows[i];
    const int l = level_of_row(P.pyr, r);
    // ld_head.py:246-252: a level without positives contributes no loss_im (and no gradient)
    const bool live = P.counts[4 + 3 * l + 0] > 0;
    const float k = live ? P.im_w / ((float)P.counts[4 + 3 * l + 2] * (float)F) : 0.f;
    float acc = 0.f;
    for (int c = lane; c < F; c += 32) {
      float a, b;
      if (sizeof(FT) == 4) {
        a = ((const float*)xs_)[(size_t)r * F + c];
        b = ((const float*)xt_)[(size_t)r * F + c];
      } else {
        a = __bfloat162float(((const __nv_bfloat16*)xs_)[(size_t)r * F + c]);
        b = __bfloat162float(((const __nv_bfloat16*)xt_)[(size_t)r * F + c]);
      }
      const float d = a - b;
      acc += d * d;
      if (wg) store_grad<GT>(d_im + (size_t)i * F + c, P.tg[7][l] * k * 2.f * d);
    }
    acc = warp_sum(acc);
    if (lane == 0 && live) atomicAdd(&s_loss[7][l], k * acc);
  }

  __syncthreads();
  for (int i = tid; i < LDB_NUM_TERMS * LDB_MAX_LEVELS; i += LOSS_THREADS) {
    const float vv = (&s_loss[0][0])[i];
    if (vv != 0.f) atomicAdd(P.losses + i, vv);
  }
  (void)s_red;
}

template <typename T>
__global__ void scatter_add_rows_kernel(T* __restrict__ dst, const T* __restrict__ src,
                                        const int* __restrict__ rows, const int* __restrict__ n_rows,
                                        int max_rows, int ch) {
  const int n = min(*n_rows, max_rows);
  const long long total = (long long)n * ch;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / ch), c = (int)(e - (long long)i * ch);
    const size_t o = (size_t)rows[i] * ch + c;
    if (sizeof(T) == 4) ((float*)dst)[o] += ((const float*)src)[e];
    else ((__nv_bfloat16*)dst)[o] = __float2bfloat16_rn(__bfloat162float(((const __nv_bfloat16*)dst)[o]) +
                                                        __bfloat162float(((const __nv_bfloat16*)src)[e]));
  }
}

// =============================================================================================
// Dense LD-KL forward+backward (BASELINE.json configs[4]).  HBM-bound: 208 B per 17-bin row
// (read student 68 + teacher 68 + weight 4, write gradient 68).  Persistent CTAs stream 256-row
// chunks through shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier, double
// buffered), one thread per row computes both softmaxes from conflict-free smem (row stride 17
// words), the gradient chunk leaves through a TMA bulk store.
constexpr int KL_ROWS = 256;  // rows per chunk == threads per CTA
constexpr int KL_STAGES = 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}

template <int BINS>
__global__ void __launch_bounds__(KL_ROWS, 2)
kl_dense_kernel(const float* __restrict__ S, const float* __restrict__ Tt, const float* __restrict__ Wt,
                long long rows, float T, float kscale /* loss_weight/avg_factor */, float grad_out,
                float* __restrict__ loss, float* __restrict__ dS) {
  extern __shared__ __align__(128) unsigned char kl_smem[];
  constexpr int CH = KL_ROWS * BINS;  // floats per chunk
  float* sS = (float*)kl_smem;                       // [STAGES][CH]
  float* sT = sS + KL_STAGES * CH;                   // [STAGES][CH]
  float* sW = sT + KL_STAGES * CH;                   // [STAGES][KL_ROWS]
  float* sO = sW + KL_STAGES * KL_ROWS;              // [CH]
  __shared__ uint64_t bar[KL_STAGES];
  __shared__ float s_red[KL_ROWS / 32];
  const int tid = threadIdx.x;
  const long long full_chunks = rows / KL_ROWS;  // tail handled separately
  if (tid == 0) {
    for (int s = 0; s < KL_STAGES; ++s) mbar_init(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const float invT = 1.f / T;
  const float kl_scale = T * T / (float)BINS, g_scale = T / (float)BINS * kscale * grad_out;
  float acc = 0.f;

  auto issue = [&](long long chunk, int stage) {
    mbar_expect_tx(&bar[stage], (uint32_t)(2 * CH * 4 + KL_ROWS * 4));
    bulk_load(sS + stage * CH, S + chunk * CH, CH * 4, &bar[stage]);
    bulk_load(sT + stage * CH, Tt + chunk * CH, CH * 4, &bar[stage]);
    bulk_load(sW + stage * KL_ROWS, Wt + chunk * KL_ROWS, KL_ROWS * 4, &bar[stage]);
  };

  long long c = blockIdx.x;
  if (tid == 0) {
    long long cc = c;
    for (int s = 0; s < KL_STAGES && cc < full_chunks; ++s, cc += gridDim.x) issue(cc, s);
  }
  int it = 0;
  for (; c < full_chunks; c += gridDim.x, ++it) {
    const int stage = it % KL_STAGES;
    const uint32_t parity = (it / KL_STAGES) & 1;
    mbar_wait(&bar[stage], parity);
    const float* s = sS + stage * CH + tid * BINS;
    const float* t = sT + stage * CH + tid * BINS;
    const float w = sW[stage * KL_ROWS + tid];
    float zs[BINS], zt[BINS];
    float ms = -INFINITY, mt = -INFINITY;
#pragma unroll
    for (int k = 0; k < BINS; ++k) {
      zs[k] = s[k] * invT;
      zt[k] = t[k] * invT;
      ms = fmaxf(ms, zs[k]);
      mt = fmaxf(mt, zt[k]);
    }
    float Ss = 0.f, St = 0.f;
#pragma unroll
    for (int k = 0; k < BINS; ++k) {
      zs[k] = __expf(zs[k] - ms);   // inputs are <= 0 after the max subtraction: ex2.approx is
      zt[k] = __expf(zt[k] - mt);   // accurate to ~2 ulp there (tolerance 1e-4, tests/test_kl_dense)
      Ss += zs[k];
      St += zt[k];
    }
    const float rS = 1.f / Ss, rT = 1.f / St;
    // sum_k pt*(log pt - log ps) = sum_k pt*((t_k - mt) - (s_k - ms)) - log St + log Ss
    float kl = 0.f;
    // wait until the previous chunk's bulk store has finished reading sO
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();
    float* o = sO + tid * BINS;
    const float gw = g_scale * w;
#pragma unroll
    for (int k = 0; k < BINS; ++k) {
      const float ps = zs[k] * rS, pt = zt[k] * rT;
      kl += pt * ((t[k] * invT - mt) - (s[k] * invT - ms));
      o[k] = gw * (ps - pt);
    }
    kl += __logf(Ss) - __logf(St);
    acc += w * kl;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();   // all reads of this stage + writes of sO done
    if (tid == 0) {
      bulk_store(dS + c * CH, sO, CH * 4);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      const long long nxt = c + (long long)KL_STAGES * gridDim.x;
      if (nxt < full_chunks) issue(nxt, stage);
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  // tail rows (< KL_ROWS), plain loads, first CTA only
  if (blockIdx.x == 0) {
    const long long r = full_chunks * KL_ROWS + tid;
    if (r < rows) {
      const float* s = S + r * BINS;
      const float* t = Tt + r * BINS;
      const float w = Wt[r];
      float ms = -INFINITY, mt = -INFINITY;
      for (int k = 0; k < BINS; ++k) {
        ms = fmaxf(ms, s[k] * invT);
        mt = fmaxf(mt, t[k] * invT);
      }
      float Ss = 0.f, St = 0.f;
      for (int k = 0; k < BINS; ++k) {
        Ss += __expf(s[k] * invT - ms);
        St += __expf(t[k] * invT - mt);
      }
      float kl = 0.f;
      for (int k = 0; k < BINS; ++k) {
        const float ps = __expf(s[k] * invT - ms) / Ss, pt = __expf(t[k] * invT - mt) / St;
        kl += pt * ((t[k] * invT - mt) - (s[k] * invT - ms));
        dS[r * BINS + k] = g_scale * w * (ps - pt);
      }
      kl += __logf(Ss) - __logf(St);
      acc += w * kl;
    }
  }
  acc = warp_sum(acc);
  if ((tid & 31) == 0) s_red[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    float tsum = 0.f;
    for (int i = 0; i < KL_ROWS / 32; ++i) tsum += s_red[i];
    atomicAdd(loss, tsum * kl_scale * kscale);
  }
}

// generic-bins fallback: one warp per row
__global__ void __launch_bounds__(256)
kl_dense_generic_kernel(const float* __restrict__ S, const float* __restrict__ Tt, const float* __restrict__ Wt,
                        long long rows, int bins, float T, float kscale, float grad_out,
                        float* __restrict__ loss, float* __restrict__ dS) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float invT = 1.f / T;
  float acc = 0.f;
  for (long long r = warp; r < rows; r += nwarps) {
    float ms = -INFINITY, mt = -INFINITY;
    for (int k = lane; k < bins; k += 32) {
      ms = fmaxf(ms, S[r * bins + k] * invT);
      mt = fmaxf(mt, Tt[r * bins + k] * invT);
    }
    ms = warp_max(ms);
    mt = warp_max(mt);
    float Ss = 0.f, St = 0.f;
    for (int k = lane; k < bins; k += 32) {
      Ss += expf(S[r * bins + k] * invT - ms);
      St += expf(Tt[r * bins + k] * invT - mt);
    }
    Ss = warp_sum(Ss);
    St = warp_sum(St);
    const float w = Wt[r];
    float kl = 0.f;
    for (int k = lane; k < bins; k += 32) {
      const float ls = S[r * bins + k] * invT - ms - logf(Ss), lt = Tt[r * bins + k] * invT - mt - logf(St);
      const float ps = expf(ls), pt = expf(lt);
      if (pt > 0.f) kl += pt * (lt - ls);
      dS[r * bins + k] = grad_out * kscale * w * (T / (float)bins) * (ps - pt);
    }
    acc += w * warp_sum(kl);
  }
  if (lane == 0 && acc != 0.f) atomicAdd(loss, acc * (T * T / (float)bins) * kscale);
}

static int fill_params(const ldb_pyramid* pyr, const ldb_loss_cfg* cfg, LossParams* P) {
  memset(P, 0, sizeof(*P));
  LDB_REQUIRE(make_pyr(pyr, &P->pyr) == 0, LDB_EINVAL, "ld_loss: bad pyramid");
  LDB_REQUIRE(cfg, LDB_EINVAL, "ld_loss: null cfg");
  LDB_REQUIRE(cfg->num_classes >= 1 && cfg->num_classes <= 32 * MAX_CPL, LDB_ENOSUP,
              "ld_loss: num_classes %d not in [1,%d]", cfg->num_classes, 32 * MAX_CPL);
  LDB_REQUIRE(cfg->reg_max >= 1 && cfg->reg_max + 1 <= 32, LDB_ENOSUP, "ld_loss: reg_max %d > 31", cfg->reg_max);
  LDB_REQUIRE((long long)P->pyr.N * P->pyr.off[P->pyr.L] < (1ll << 31) / 128, LDB_ENOSUP, "ld_loss: too many rows");
  P->C = cfg->num_classes;
  P->bins = cfg->reg_max + 1;
  P->F = cfg->feat_channels;
  P->qfl_w = cfg->qfl_weight; P->beta = cfg->qfl_beta;
  P->giou_w = cfg->giou_weight; P->eps = cfg->giou_eps;
  P->dfl_w = cfg->dfl_weight;
  P->ld_w = cfg->ld_weight; P->ld_T = cfg->ld_T;
  P->ldv_w = cfg->ld_vlr_weight; P->ldv_T = cfg->ld_vlr_T;
  P->kd_w = cfg->kd_weight; P->kd_T = cfg->kd_T;
  P->im_w = cfg->im_weight;
  return 0;
}

static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

}  // namespace ldb

using namespace ldb;

extern "C" int ldb200_ld_loss_prepass(const ldb_pyramid* pyr, const ldb_loss_cfg* cfg, const float* cls,
                                      const float* reg, const int32_t* labels, const float* bbox_targets,
                                      const int32_t* sparse_rows, const int32_t* counts, float* score,
                                      float* wsum, void* stream_) {
  LossParams P;
  int e = fill_params(pyr, cfg, &P);
  if (e) return e;
  LDB_REQUIRE(cls && reg && labels && bbox_targets && sparse_rows && counts && wsum, LDB_EINVAL,
              "ld_loss_prepass: null pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  P.cls = cls; P.reg = reg; P.labels = labels; P.bbox_t = bbox_targets;
  P.sparse_rows = sparse_rows; P.counts = counts; P.score = score; P.wsum = wsum;
  const size_t R = (size_t)P.pyr.N * P.pyr.off[P.pyr.L];
  if (score) LDB_CUDA(cudaMemsetAsync(score, 0, R * sizeof(float), stream));
  LDB_CUDA(cudaMemsetAsync(wsum, 0, sizeof(float), stream));
  ld_prepass_kernel<false><<<num_sms(), LOSS_THREADS, 0, stream>>>(P);
  LDB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ldb200_ld_loss_fused(const ldb_pyramid* pyr, const ldb_loss_cfg* cfg, const float* cls,
                                    const float* reg, const float* t_cls, const float* t_reg, const void* x,
                                    const void* t_x, int feat_dtype, const int32_t* labels,
                                    const float* label_weights, const float* bbox_targets, const float* vlr,
                                    const float* score, const int32_t* sparse_rows, const int32_t* fg_rows,
                                    const int32_t* counts, const float* norm, const float* term_grad_host,
                                    float* losses, void* d_cls, void* d_reg, void* d_im, int d_cls_ld,
                                    int d_reg_ld, int grad_dtype, void* stream_) {
  (void)score;
  LossParams P;
  int e = fill_params(pyr, cfg, &P);
  if (e) return e;
  LDB_REQUIRE(cls && reg && t_cls && t_reg && x && t_x && labels && label_weights && bbox_targets && vlr &&
              sparse_rows && fg_rows && counts && norm && losses, LDB_EINVAL, "ld_loss_fused: null pointer");
  LDB_REQUIRE(feat_dtype == LDB_F32 || feat_dtype == LDB_BF16, LDB_EINVAL, "ld_loss_fused: feat_dtype");
  LDB_REQUIRE(grad_dtype == LDB_F32 || grad_dtype == LDB_BF16, LDB_EINVAL, "ld_loss_fused: grad_dtype");
  const bool wg = term_grad_host != nullptr;
  if (wg) LDB_REQUIRE(d_cls && d_reg && d_im, LDB_EINVAL, "ld_loss_fused: gradients requested but null outputs");
  cudaStream_t stream = (cudaStream_t)stream_;
  P.cls = cls; P.reg = reg; P.t_cls = t_cls; P.t_reg = t_reg; P.x = x; P.t_x = t_x;
  P.feat_bf16 = feat_dtype == LDB_BF16;
  P.labels = labels; P.label_w = label_weights; P.bbox_t = bbox_targets; P.vlr = vlr;
  P.sparse_rows = sparse_rows; P.fg_rows = fg_rows; P.counts = counts; P.norm = norm;
  P.want_grad = wg ? 1 : 0;
  {
    const char* ph = getenv("LDB200_LOSS_PHASES");   // profiling aid only
    P.phases = ph ? atoi(ph) : 7;
  }
  if (wg) memcpy(P.tg, term_grad_host, sizeof(P.tg));
  P.losses = losses; P.d_cls = d_cls; P.d_reg = d_reg; P.d_im = d_im;
  P.grad_bf16 = grad_dtype == LDB_BF16;
  P.dcls_ld = d_cls_ld > 0 ? d_cls_ld : P.C;
  P.dreg_ld = d_reg_ld > 0 ? d_reg_ld : 4 * P.bins;
  LDB_REQUIRE(P.dcls_ld >= P.C && P.dreg_ld >= 4 * P.bins && P.dcls_ld % 4 == 0, LDB_EINVAL,
              "ld_loss_fused: gradient leading dimensions");
  const size_t R = (size_t)P.pyr.N * P.pyr.off[P.pyr.L];
  const size_t gsz = P.grad_bf16 ? 2 : 4;
  LDB_CUDA(cudaMemsetAsync(losses, 0, sizeof(float) * LDB_NUM_TERMS * LDB_MAX_LEVELS, stream));
  if (wg) LDB_CUDA(cudaMemsetAsync(d_reg, 0, R * P.dreg_ld * gsz, stream));
  const int grid = num_sms() * 3;
  if (!P.grad_bf16 && !P.feat_bf16) ld_loss_fused_kernel<float, float, false><<<grid, LOSS_THREADS, 0, stream>>>(P);
  else if (!P.grad_bf16 && P.feat_bf16) ld_loss_fused_kernel<float, __nv_bfloat16, false><<<grid, LOSS_THREADS, 0, stream>>>(P);
  else if (P.grad_bf16 && !P.feat_bf16) ld_loss_fused_kernel<__nv_bfloat16, float, false><<<grid, LOSS_THREADS, 0, stream>>>(P);
  else ld_loss_fused_kernel<__nv_bfloat16, __nv_bfloat16, false><<<grid, LOSS_THREADS, 0, stream>>>(P);
  LDB_CUDA(cudaGetLastError());
  return 0;
}

// ---- LDv2Head (ld_gflv2.py:116-375): same launches, V2 instantiations.  `score` = joint scores sigmoid(cls_feat) *
// quality [R,C] (from ldb200_dgqp_forward), `cls_feat` / `t_cls_feat` = student / teacher class logits; C =
// cfg->num_classes = cls_out_channels (num_fg + 1 when loss_cls.use_sigmoid=False), labels >= num_fg are background.
extern "C" int ldb200_ld_loss_prepass_v2(const ldb_pyramid* pyr, const ldb_loss_cfg* cfg, const float* score,
                                         const float* reg, const int32_t* labels, const float* bbox_targets,
                                         const int32_t* sparse_rows, const int32_t* counts, int num_fg,
                                         float* iou_score, float* wsum, void* stream_) {
  LossParams P;
  int e = fill_params(pyr, cfg, &P);
  if (e) return e;
  LDB_REQUIRE(score && reg && labels && bbox_targets && sparse_rows && counts && wsum, LDB_EINVAL,
              "ld_loss_prepass_v2: null pointer");
  LDB_REQUIRE(num_fg >= 1 && num_fg <= P.C, LDB_EINVAL, "ld_loss_prepass_v2: num_fg");
  cudaStream_t stream = (cudaStream_t)stream_;
  P.cls = score; P.reg = reg; P.labels = labels; P.bbox_t = bbox_targets;
  P.sparse_rows = sparse_rows; P.counts = counts; P.score = iou_score; P.wsum = wsum;
  P.cls_feat = nullptr; P.d_kd = nullptr; P.num_fg = num_fg;
  const size_t R = (size_t)P.pyr.N * P.pyr.off[P.pyr.L];
  if (iou_score) LDB_CUDA(cudaMemsetAsync(iou_score, 0, R * sizeof(float), stream));
  LDB_CUDA(cudaMemsetAsync(wsum, 0, sizeof(float), stream));
  ld_prepass_kernel<true><<<num_sms(), LOSS_THREADS, 0, stream>>>(P);
  LDB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ldb200_ld_loss_fused_v2(const ldb_pyramid* pyr, const ldb_loss_cfg* cfg, const float* score,
                                       const float* cls_feat, const float* reg, const float* t_cls_feat,
                                       const float* t_reg, const void* x, const void* t_x, int feat_dtype,
                                       const int32_t* labels, const float* label_weights, const float* bbox_targets,
                                       const float* vlr, const int32_t* sparse_rows, const int32_t* fg_rows,
                                       const int32_t* counts, const float* norm, const float* term_grad_host,
                                       int num_fg, float* losses, void* d_score, void* d_kd, void* d_reg, void* d_im,
                                       int d_cls_ld, int d_reg_ld, int grad_dtype, void* stream_) {
  LossParams P;
  int e = fill_params(pyr, cfg, &P);
  if (e) return e;
  LDB_REQUIRE(score && cls_feat && reg && t_cls_feat && t_reg && x && t_x && labels && label_weights && bbox_targets &&
              vlr && sparse_rows && fg_rows && counts && norm && losses, LDB_EINVAL, "ld_loss_fused_v2: null pointer");
  LDB_REQUIRE(num_fg >= 1 && num_fg <= P.C, LDB_EINVAL, "ld_loss_fused_v2: num_fg");
  LDB_REQUIRE(feat_dtype == LDB_F32 || feat_dtype == LDB_BF16, LDB_EINVAL, "ld_loss_fused_v2: feat_dtype");
  LDB_REQUIRE(grad_dtype == LDB_F32 || grad_dtype == LDB_BF16, LDB_EINVAL, "ld_loss_fused_v2: grad_dtype");
  const bool wg = term_grad_host != nullptr;
  if (wg) LDB_REQUIRE(d_score && d_kd && d_reg && d_im, LDB_EINVAL, "ld_loss_fused_v2: gradients requested but null outputs");
  cudaStream_t stream = (cudaStream_t)stream_;
  P.cls = score; P.reg = reg; P.t_cls = t_cls_feat; P.t_reg = t_reg; P.x = x; P.t_x = t_x;
  P.cls_feat = cls_feat; P.d_kd = d_kd; P.num_fg = num_fg;
  P.feat_bf16 = feat_dtype == LDB_BF16;
  P.labels = labels; P.label_w = label_weights; P.bbox_t = bbox_targets; P.vlr = vlr;
  P.sparse_rows = sparse_rows; P.fg_rows = fg_rows; P.counts = counts; P.norm = norm;
  P.want_grad = wg ? 1 : 0;
  P.phases = 7;
  if (wg) memcpy(P.tg, term_grad_host, sizeof(P.tg));
  P.losses = losses; P.d_cls = d_score; P.d_reg = d_reg; P.d_im = d_im;
  P.grad_bf16 = grad_dtype == LDB_BF16;
  P.dcls_ld = d_cls_ld > 0 ? d_cls_ld : P.C;
  P.dreg_ld = d_reg_ld > 0 ? d_reg_ld : 4 * P.bins;
  LDB_REQUIRE(P.dcls_ld >= P.C && P.dreg_ld >= 4 * P.bins && P.dcls_ld % 4 == 0, LDB_EINVAL,
              "ld_loss_fused_v2: gradient leading dimensions");
  const size_t R = (size_t)P.pyr.N * P.pyr.off[P.pyr.L];
  const size_t gsz = P.grad_bf16 ? 2 : 4;
  LDB_CUDA(cudaMemsetAsync(losses, 0, sizeof(float) * LDB_NUM_TERMS * LDB_MAX_LEVELS, stream));
  if (wg) {
    LDB_CUDA(cudaMemsetAsync(d_reg, 0, R * P.dreg_ld * gsz, stream));
    LDB_CUDA(cudaMemsetAsync(d_kd, 0, R * P.dcls_ld * gsz, stream));
  }
  const int grid = num_sms() * 3;
  if (!P.grad_bf16 && !P.feat_bf16) ld_loss_fused_kernel<float, float, true><<<grid, LOSS_THREADS, 0, stream>>>(P);
  else if (!P.grad_bf16 && P.feat_bf16) ld_loss_fused_kernel<float, __nv_bfloat16, true><<<grid, LOSS_THREADS, 0, stream>>>(P);
  else if (P.grad_bf16 && !P.feat_bf16) ld_loss_fused_kernel<__nv_bfloat16, float, true><<<grid, LOSS_THREADS, 0, stream>>>(P);
  else ld_loss_fused_kernel<__nv_bfloat16, __nv_bfloat16, true><<<grid, LOSS_THREADS, 0, stream>>>(P);
  LDB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ldb200_scatter_add_rows(void* dst, const void* src, const int32_t* rows, const int32_t* n_rows_dev,
                                       int max_rows, int channels, int dtype, void* stream_) {
  LDB_REQUIRE(dst && src && rows && n_rows_dev && channels > 0, LDB_EINVAL, "scatter_add_rows: bad argument");
  if (max_rows <= 0) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  const long long total = (long long)max_rows * channels;
  const int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  if (dtype == LDB_F32)
    scatter_add_rows_kernel<float><<<grid, 256, 0, stream>>>((float*)dst, (const float*)src, rows, n_rows_dev, max_rows, channels);
  else if (dtype == LDB_BF16)
    scatter_add_rows_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((__nv_bfloat16*)dst, (const __nv_bfloat16*)src, rows, n_rows_dev, max_rows, channels);
  else
    LDB_REQUIRE(false, LDB_EINVAL, "scatter_add_rows: dtype");
  LDB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ldb200_kl_dense_fwd_bwd(const float* student, const float* teacher, const float* weight,
                                       int64_t rows, int bins, float T, float loss_weight, float avg_factor,
                                       float grad_out, float* loss, float* d_student, void* stream_) {
  LDB_REQUIRE(rows >= 0 && bins >= 2 && T >= 1.f && avg_factor > 0.f, LDB_EINVAL, "kl_dense: bad argument");
  if (rows == 0) return 0;   // empty input: loss unchanged, nothing to write (pointers may be null)
  LDB_REQUIRE(student && teacher && weight && loss && d_student, LDB_EINVAL, "kl_dense: null pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  const float kscale = loss_weight / avg_factor;
  const bool aligned = (((uintptr_t)student | (uintptr_t)teacher | (uintptr_t)weight | (uintptr_t)d_student) & 15) == 0;
  if (bins == 17 && aligned) {
    constexpr int CH = KL_ROWS * 17;
    const size_t smem = (size_t)(2 * KL_STAGES * CH + KL_STAGES * KL_ROWS + CH) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
      LDB_CUDA(cudaFuncSetAttribute(kl_dense_kernel<17>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    const long long chunks = rows / KL_ROWS;
    long long grid = (long long)num_sms() * 2;
    if (grid > chunks) grid = chunks > 0 ? chunks : 1;
    kl_dense_kernel<17><<<(unsigned)grid, KL_ROWS, smem, stream>>>(student, teacher, weight, rows, T, kscale, grad_out,
                                                                  loss, d_student);
  } else {
    kl_dense_generic_kernel<<<num_sms() * 8, 256, 0, stream>>>(student, teacher, weight, rows, bins, T, kscale,
                                                              grad_out, loss, d_student);
  }
  LDB_CUDA(cudaGetLastError());
  return 0;
}


// =============================================================================================
// Stand-alone loss modules of the registry surface ("Loss modules: forward(pred, target, weight,
// avg_factor, reduction_override)"): QualityFocalLoss (gfocal_loss.py:8-50, 77-135), DistributionFocalLoss
// (gfocal_loss.py:53-74, 138-179) and GIoULoss (iou_loss.py:85-102, 325-360) outside LDHead, forward + gradient in
// one launch, on the same row arithmetic as the fused kernel.  `scale` = loss_weight / avg_factor (or 1/rows, or 1):
//   row_loss[r] (optional, reduction='none') = loss_weight * w_r * l_r ;  loss[0] += scale * sum_r w_r * l_r
//   d_pred = scale * w_r * d l_r / d pred      (row_loss mode: scale = loss_weight, upstream applied by the caller)
namespace ldb {

__device__ __forceinline__ void block_accumulate(float v, float* out) {
  __shared__ float sh[32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    if (t != 0.f) atomicAdd(out, t);
  }
}

// one warp per row: lanes stride over the C class columns
__global__ void __launch_bounds__(256)
qfl_rows_kernel(const float* __restrict__ pred, const int* __restrict__ label, const float* __restrict__ score,
                const float* __restrict__ weight, int R, int C, float beta, int use_sigmoid, float scale,
                float* __restrict__ row_loss, float* __restrict__ loss, float* __restrict__ d_pred) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float acc = 0.f;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < R; r += gridDim.x * wpb) {
    const int lab = label[r];
    const float w = weight ? weight[r] : 1.f, y = score[r];
    float rl = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float x = pred[(size_t)r * C + c];
      float lo, gr;
      if (use_sigmoid) {
        if (c == lab) qfl_pos(x, y, beta, lo, gr); else qfl_neg(x, beta, lo, gr);
      } else {
        if (c == lab) qflp_pos(x, y, beta, lo, gr); else qflp_neg(x, beta, lo, gr);
      }
      rl += lo;
      if (d_pred) d_pred[(size_t)r * C + c] = scale * w * gr;
    }
    rl = warp_sum(rl);
    if (lane == 0) {
      if (row_loss) row_loss[r] = scale * w * rl;
      acc += w * rl;
    }
  }
  if (loss) block_accumulate(acc * scale, loss);
}

// one warp per row of `bins` logits (bins <= 32)
__global__ void __launch_bounds__(256)
dfl_rows_kernel(const float* __restrict__ pred, const float* __restrict__ label, const float* __restrict__ weight, int R,
                int bins, float scale, float* __restrict__ row_loss, float* __restrict__ loss,
                float* __restrict__ d_pred) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float acc = 0.f;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < R; r += gridDim.x * wpb) {
    const bool active = lane < bins;
    const float z = active ? pred[(size_t)r * bins + lane] : 0.f;
    float p, logp;
    warp_softmax(z, active, p, logp);
    const float y = label[r], w = weight ? weight[r] : 1.f;
    const int left = (int)y;                       // label.long(): truncation, labels are >= 0
    const float wl = (float)(left + 1) - y, wr = y - (float)left;
    // cross_entropy(pred, left) * wl + cross_entropy(pred, right) * wr (gfocal_loss.py:69-73)
    float contrib = 0.f;
    if (active && lane == left) contrib -= logp * wl;
    if (active && lane == left + 1) contrib -= logp * wr;
    const float rl = warp_sum(contrib);
    if (d_pred && active) {
      float g = p * (wl + wr);
      if (lane == left) g -= wl;
      if (lane == left + 1) g -= wr;
      d_pred[(size_t)r * bins + lane] = scale * w * g;
    }
    if (lane == 0) {
      if (row_loss) row_loss[r] = scale * w * rl;
      acc += w * rl;
    }
  }
  if (loss) block_accumulate(acc * scale, loss);
}

// one thread per box pair
__global__ void __launch_bounds__(256)
giou_rows_kernel(const float* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ weight,
                 int R, float eps, float scale, float* __restrict__ row_loss, float* __restrict__ loss,
                 float* __restrict__ d_pred) {
  float acc = 0.f;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
    const float4 pv = reinterpret_cast<const float4*>(pred)[r], tv = reinterpret_cast<const float4*>(target)[r];
    const float p[4] = {pv.x, pv.y, pv.z, pv.w}, t[4] = {tv.x, tv.y, tv.z, tv.w};
    float iou, l, g[4];
    giou_fwd_bwd(p, t, eps, iou, l, g);
    const float w = weight ? weight[r] : 1.f;
    if (d_pred) reinterpret_cast<float4*>(d_pred)[r] = make_float4(scale * w * g[0], scale * w * g[1], scale * w * g[2], scale * w * g[3]);
    if (row_loss) row_loss[r] = scale * w * l;
    acc += w * l;
  }
  if (loss) block_accumulate(acc * scale, loss);
}

}  // namespace ldb

extern "C" int ldb200_qfl_fwd_bwd(const float* pred, const int32_t* label, const float* score, const float* weight,
                                  int64_t rows, int C, float beta, int use_sigmoid, float scale, float* row_loss,
                                  float* loss, float* d_pred, void* stream_) {
  LDB_REQUIRE(rows >= 0 && C >= 1 && rows < (1LL << 31), LDB_EINVAL, "qfl_fwd_bwd: bad shape");
  if (rows == 0) return 0;
  LDB_REQUIRE(pred && label && score && (row_loss || loss), LDB_EINVAL, "qfl_fwd_bwd: null pointer");
  long long blocks = (rows + 7) / 8;
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  ldb::qfl_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(pred, label, score, weight, (int)rows, C, beta,
                                                                           use_sigmoid, scale, row_loss, loss, d_pred);
  LDB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ldb200_dfl_fwd_bwd(const float* pred, const float* label, const float* weight, int64_t rows, int bins,
                                  float scale, float* row_loss, float* loss, float* d_pred, void* stream_) {
  LDB_REQUIRE(rows >= 0 && bins >= 2 && bins <= 32 && rows < (1LL << 31), LDB_EINVAL, "dfl_fwd_bwd: 2 <= bins <= 32");
  if (rows == 0) return 0;
  LDB_REQUIRE(pred && label && (row_loss || loss), LDB_EINVAL, "dfl_fwd_bwd: null pointer");
  long long blocks = (rows + 7) / 8;
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  ldb::dfl_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(pred, label, weight, (int)rows, bins, scale,
                                                                           row_loss, loss, d_pred);
  LDB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ldb200_giou_fwd_bwd(const float* pred, const float* target, const float* weight, int64_t rows, float eps,
                                   float scale, float* row_loss, float* loss, float* d_pred, void* stream_) {
  LDB_REQUIRE(rows >= 0 && rows < (1LL << 31), LDB_EINVAL, "giou_fwd_bwd: bad shape");
  if (rows == 0) return 0;
  LDB_REQUIRE(pred && target && (row_loss || loss), LDB_EINVAL, "giou_fwd_bwd: null pointer");
  LDB_REQUIRE((((uintptr_t)pred | (uintptr_t)target | (uintptr_t)d_pred) & 15) == 0, LDB_EINVAL, "giou_fwd_bwd: 16-byte alignment");
  long long blocks = (rows + 255) / 256;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  ldb::giou_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(pred, target, weight, (int)rows, eps, scale,
                                                                            row_loss, loss, d_pred);
  LDB_CUDA(cudaGetLastError());
  return 0;
}
