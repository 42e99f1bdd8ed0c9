// GFLv2 Distribution-Guided Quality Predictor on row-space head outputs:
// `GFocalHead.forward_single` (mmdet/models/dense_heads/gfocal_head.py:198-215)
//     prob  = softmax(bbox_pred.reshape(N, 4, reg_max+1, H, W), dim=2);  stat = cat(topk(prob, 4), mean(topk))
//     q     = sigmoid(conv1x1(relu(conv1x1(stat))))            (reg_conf, gfocal_head.py:147-152)
//     score = sigmoid(cls_feat) * q
// and its backward (autograd in the reference).  The per-row arithmetic lives in v2_math.cuh (host+device, checked
// against torch autograd on CPU by tests/test_v2_math_host.py); this file adds the row-space indexing, the fused
// d_score -> (d_cls_feat, dq) step and the block-level reduction of the 1x1-conv weight gradients.
//
// STATUS (round 1): compiled for sm_100a and exported, not yet exercised on a GPU -- the LDv2 loss variant that
// consumes `score` lands next round; nothing on the LDHead path calls into this file.
#include "common.cuh"
#include "v2_math.cuh"

namespace ldb {
namespace {

constexpr int DG_ROWS = 64;   // rows per block = threads per block

struct DgqpParams {
  DgqpShape s;
  int rows, C;
  int reg_ld, cls_ld;          // row strides (elements) of reg / cls_feat
  const float* reg;
  const float* cls_feat;
  const float* W1;             // [hid][nstat]   reg_conf.0.weight
  const float* b1;             // [hid]          reg_conf.0.bias
  const float* W2;             // [hid]          reg_conf.2.weight
  const float* b2;             // [1]            reg_conf.2.bias
  float* quality;              // [rows]
  float* score;                // [rows, C] (ld = C)
};

__global__ void __launch_bounds__(DG_ROWS) dgqp_forward_kernel(DgqpParams P) {
  __shared__ float s_q[DG_ROWS];
  const int r0 = blockIdx.x * DG_ROWS, r = r0 + threadIdx.x;
  float q = 0.f;
  if (r < P.rows) {
    q = dgqp_row_forward(P.reg + (size_t)r * P.reg_ld, P.s, P.W1, P.b1, P.W2, __ldg(P.b2));
    P.quality[r] = q;
  }
  s_q[threadIdx.x] = q;
  __syncthreads();
  const int nrows = min(DG_ROWS, P.rows - r0);
  for (int e = threadIdx.x; e < nrows * P.C; e += DG_ROWS) {
    const int rr = e / P.C, c = e - rr * P.C;
    const float f = __ldg(P.cls_feat + (size_t)(r0 + rr) * P.cls_ld + c);
    P.score[(size_t)(r0 + rr) * P.C + c] = s_q[rr] / (1.f + expf(-f));
  }
}

template <typename GT> __device__ __forceinline__ float load_grad(const GT* p);
template <> __device__ __forceinline__ float load_grad<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float load_grad<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename GT> __device__ __forceinline__ void store_grad2(GT* p, float v);
template <> __device__ __forceinline__ void store_grad2<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void store_grad2<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

struct DgqpBwdParams {
  DgqpParams f;
  const void* d_score;     // [rows, dscore_ld] grad dtype: d loss / d joint score
  const void* d_kd;        // optional [rows, dscore_ld] grad dtype: extra gradient on cls_feat (KD term), may be null
  void* d_cls_feat;        // [rows, dscore_ld] grad dtype (written)
  void* d_reg;             // [rows, dreg_ld]   grad dtype (the DGQP contribution is ADDED)
  int dscore_ld, dreg_ld;
  float* dW1;              // [hid][nstat]  accumulated (atomics)
  float* db1;              // [hid]
  float* dW2;              // [hid]
  float* db2;              // [1]
};

template <typename GT>
__global__ void __launch_bounds__(DG_ROWS) dgqp_backward_kernel(DgqpBwdParams B) {
  const DgqpParams& P = B.f;
  __shared__ float s_dq[DG_ROWS], s_do[DG_ROWS];
  __shared__ float s_dh[DG_ROWS][DGQP_MAX_HID + 1];
  __shared__ float s_h[DG_ROWS][DGQP_MAX_HID + 1];
  __shared__ float s_stat[DG_ROWS][DGQP_MAX_STAT + 1];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int r0 = blockIdx.x * DG_ROWS;
  const int nrows = min(DG_ROWS, P.rows - r0);
  const GT* dS = (const GT*)B.d_score;
  const GT* dK = (const GT*)B.d_kd;
  GT* dF = (GT*)B.d_cls_feat;
  // 1. joint score = sigmoid(f) * q:  d f = d_score * q * s (1 - s) (+ KD gradient),  dq = sum_c d_score * s
  for (int rr = wid; rr < nrows; rr += DG_ROWS / 32) {
    const int r = r0 + rr;
    const float q = P.quality[r];
    float acc = 0.f;
    for (int c = lane; c < P.C; c += 32) {
      const float f = __ldg(P.cls_feat + (size_t)r * P.cls_ld + c);
      const float sg = 1.f / (1.f + expf(-f));
      const float g = load_grad<GT>(dS + (size_t)r * B.dscore_ld + c);
      acc += g * sg;
      float df = g * q * sg * (1.f - sg);
      if (dK) df += load_grad<GT>(dK + (size_t)r * B.dscore_ld + c);
      store_grad2<GT>(dF + (size_t)r * B.dscore_ld + c, df);
    }
    acc = warp_sum(acc);
    if (lane == 0) s_dq[rr] = acc;
  }
  __syncthreads();
  // 2. one thread per row through the predictor
  const int ns = P.s.nstat(), hid = P.s.hid;
  if (tid < nrows) {
    const int r = r0 + tid;
    float d_reg[4 * DGQP_MAX_BINS];
    for (int i = 0; i < 4 * P.s.bins; ++i) d_reg[i] = 0.f;
    float dh[DGQP_MAX_HID], st[DGQP_MAX_STAT], h[DGQP_MAX_HID];
    const float d_o = dgqp_row_backward(P.reg + (size_t)r * P.reg_ld, P.s, P.W1, P.b1, P.W2, __ldg(P.b2), s_dq[tid], d_reg,
                                        dh, st, h);
    s_do[tid] = d_o;
    for (int j = 0; j < hid; ++j) { s_dh[tid][j] = dh[j]; s_h[tid][j] = h[j]; }
    for (int i = 0; i < ns; ++i) s_stat[tid][i] = st[i];
    GT* dr = (GT*)B.d_reg + (size_t)r * B.dreg_ld;
    for (int i = 0; i < 4 * P.s.bins; ++i) store_grad2<GT>(dr + i, load_grad<GT>(dr + i) + d_reg[i]);
  } else {
    s_do[tid] = 0.f;
    for (int j = 0; j < hid; ++j) { s_dh[tid][j] = 0.f; s_h[tid][j] = 0.f; }
    for (int i = 0; i < ns; ++i) s_stat[tid][i] = 0.f;
  }
  __syncthreads();
  // 3. weight gradients of the two 1x1 convs: outer products reduced over the block's rows, then one atomic each
  for (int e = tid; e < hid * ns; e += DG_ROWS) {
    const int j = e / ns, i = e - j * ns;
    float a = 0.f;
    for (int rr = 0; rr < DG_ROWS; ++rr) a += s_dh[rr][j] * s_stat[rr][i];
    if (a != 0.f) atomicAdd(B.dW1 + e, a);
  }
  for (int j = tid; j < hid; j += DG_ROWS) {
    float a = 0.f, b = 0.f;
    for (int rr = 0; rr < DG_ROWS; ++rr) { a += s_dh[rr][j]; b += s_do[rr] * s_h[rr][j]; }
    if (a != 0.f) atomicAdd(B.db1 + j, a);
    if (b != 0.f) atomicAdd(B.dW2 + j, b);
  }
  if (tid == 0) {
    float a = 0.f;
    for (int rr = 0; rr < DG_ROWS; ++rr) a += s_do[rr];
    if (a != 0.f) atomicAdd(B.db2, a);
  }
}

int fill(DgqpParams* P, const float* reg, int reg_ld, const float* cls_feat, int cls_ld, int rows, int C, int bins, int topk,
         int add_mean, int hid, const float* W1, const float* b1, const float* W2, const float* b2) {
  LDB_REQUIRE(reg && cls_feat && W1 && b1 && W2 && b2, LDB_EINVAL, "dgqp: null pointer");
  LDB_REQUIRE(rows >= 0 && C >= 1, LDB_EINVAL, "dgqp: shape");
  LDB_REQUIRE(bins >= 2 && bins <= DGQP_MAX_BINS && topk >= 1 && topk <= DGQP_MAX_TOPK && topk <= bins && hid >= 1 &&
              hid <= DGQP_MAX_HID, LDB_ENOSUP, "dgqp: bins <= %d, reg_topk <= %d, reg_channels <= %d", DGQP_MAX_BINS,
              DGQP_MAX_TOPK, DGQP_MAX_HID);
  P->s.bins = bins; P->s.topk = topk; P->s.add_mean = add_mean ? 1 : 0; P->s.hid = hid;
  P->rows = rows; P->C = C;
  P->reg_ld = reg_ld > 0 ? reg_ld : 4 * bins;
  P->cls_ld = cls_ld > 0 ? cls_ld : C;
  P->reg = reg; P->cls_feat = cls_feat; P->W1 = W1; P->b1 = b1; P->W2 = W2; P->b2 = b2;
  return 0;
}

}  // namespace
}  // namespace ldb

using namespace ldb;

extern "C" int ldb200_dgqp_forward(const float* reg, int reg_ld, const float* cls_feat, int cls_ld, int rows, int C,
                                   int bins, int topk, int add_mean, int hid, const float* W1, const float* b1,
                                   const float* W2, const float* b2, float* quality, float* score, void* stream) {
  DgqpParams P;
  int e = fill(&P, reg, reg_ld, cls_feat, cls_ld, rows, C, bins, topk, add_mean, hid, W1, b1, W2, b2);
  if (e) return e;
  LDB_REQUIRE(quality && score, LDB_EINVAL, "dgqp_forward: null output");
  if (rows == 0) return 0;
  P.quality = quality; P.score = score;
  dgqp_forward_kernel<<<(rows + DG_ROWS - 1) / DG_ROWS, DG_ROWS, 0, (cudaStream_t)stream>>>(P);
  LDB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ldb200_dgqp_backward(const float* reg, int reg_ld, const float* cls_feat, int cls_ld, int rows, int C,
                                    int bins, int topk, int add_mean, int hid, const float* W1, const float* b1,
                                    const float* W2, const float* b2, const float* quality, const void* d_score,
                                    const void* d_kd, int dscore_ld, void* d_cls_feat, void* d_reg, int dreg_ld,
                                    int grad_dtype, float* dW1, float* db1, float* dW2, float* db2, void* stream) {
  DgqpBwdParams B;
  int e = fill(&B.f, reg, reg_ld, cls_feat, cls_ld, rows, C, bins, topk, add_mean, hid, W1, b1, W2, b2);
  if (e) return e;
  LDB_REQUIRE(quality && d_score && d_cls_feat && d_reg && dW1 && db1 && dW2 && db2, LDB_EINVAL, "dgqp_backward: null pointer");
  LDB_REQUIRE(grad_dtype == LDB_F32 || grad_dtype == LDB_BF16, LDB_EINVAL, "dgqp_backward: grad_dtype");
  if (rows == 0) return 0;
  B.f.quality = const_cast<float*>(quality);
  B.f.score = nullptr;
  B.d_score = d_score; B.d_kd = d_kd; B.d_cls_feat = d_cls_feat; B.d_reg = d_reg;
  B.dscore_ld = dscore_ld > 0 ? dscore_ld : C;
  B.dreg_ld = dreg_ld > 0 ? dreg_ld : 4 * bins;
  B.dW1 = dW1; B.db1 = db1; B.dW2 = dW2; B.db2 = db2;
  const int grid = (rows + DG_ROWS - 1) / DG_ROWS;
  if (grad_dtype == LDB_F32) dgqp_backward_kernel<float><<<grid, DG_ROWS, 0, (cudaStream_t)stream>>>(B);
  else dgqp_backward_kernel<__nv_bfloat16><<<grid, DG_ROWS, 0, (cudaStream_t)stream>>>(B);
  LDB_CUDA(cudaGetLastError());
  return 0;
}
