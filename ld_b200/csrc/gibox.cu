// General-Instance (GI) imitation region: `LDHead.get_gi_region` (mmdet/models/dense_heads/ld_head.py:613-638),
// the row selection behind `imitation_method='gibox'` (the LDHead constructor default, ld_head.py:55, 170-176).
//
// Per FPN level, over ALL locations of the level (the images of the batch share one box frame, as in the reference):
//   GI score  = max_c |sigmoid(teacher_cls) - sigmoid(student_cls)|                      (ld_head.py:627-628)
//   GI box    = the teacher's decoded box where the teacher scores higher at that class,
//               else the student's (Integral + distance2bbox in grid units)             (ld_head.py:620-634)
//   rows      = first `topk` keeps of torchvision.ops.nms(GI boxes, GI scores, iou_thr) (ld_head.py:637)
// Three launches: a warp-per-row scoring pass over the R rows, one block per level running the `topk` greedy NMS
// rounds (arg-max over the unsuppressed rows, ties to the lower index = stable descending sort, then one
// suppression sweep with torchvision's CPU arithmetic `inter / (area_i + area_j - inter) > thr`, contraction-free),
// and a finalize step that compacts the per-level picks into the `fg_rows` / `counts` form the fused loss reads.
#include "common.cuh"

namespace ldb {
namespace {

constexpr int GI_THREADS = 256;
constexpr int GI_WARPS = GI_THREADS / 32;
constexpr int NMS_THREADS = 1024;

struct GiParams {
  Pyr pyr;
  int C, bins;
  const float* cls;
  const float* t_cls;
  const float* reg;
  const float* t_reg;
  float* score;          // [R]
  float4* box;           // [R]
  unsigned char* flags;  // [R] 1 = picked or suppressed
};

__device__ __forceinline__ float gi_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// Integral (gfl_head.py:32-44): expectation of softmax(z) over bins 0..bins-1, lanes = bins
__device__ __forceinline__ float warp_integral(float z, bool active, int lane) {
  const float zz = active ? z : -INFINITY;
  const float m = warp_max(zz);
  const float e = active ? expf(zz - m) : 0.f;
  const float S = warp_sum(e);
  return warp_sum(active ? (float)lane * (e / S) : 0.f);
}

// RAW: LDv2Head.get_gi_region (ld_gflv2.py) compares the teacher's cls_feat LOGITS with the student's joint scores
// without any sigmoid
template <bool RAW>
__global__ void __launch_bounds__(GI_THREADS) gi_score_kernel(GiParams P) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int R = P.pyr.off[P.pyr.L] * P.pyr.N;
  for (int r = blockIdx.x * GI_WARPS + wid; r < R; r += gridDim.x * GI_WARPS) {
    // ---- GI score: first arg-max over classes of |sigmoid(t) - sigmoid(s)| (torch.max(dim) keeps the first maximum)
    const float* s = P.cls + (size_t)r * P.C;
    const float* t = P.t_cls + (size_t)r * P.C;
    float best = -1.f, bz = 0.f;
    int bidx = 0x7fffffff;
    for (int c = lane; c < P.C; c += 32) {
      const float z = RAW ? __ldg(t + c) - __ldg(s + c) : gi_sigmoid(__ldg(t + c)) - gi_sigmoid(__ldg(s + c));
      const float a = fabsf(z);
      if (a > best) { best = a; bidx = c; bz = z; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(FULL, best, o), oz = __shfl_xor_sync(FULL, bz, o);
      const int oi = __shfl_xor_sync(FULL, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; bz = oz; }
    }
    const bool teacher_bigger = bz >= 0.f;   // ld_head.py:629-631
    // ---- GI box: decoded box of the chosen network, grid units (anchor centre / stride = (ix, iy))
    const float* zr = (teacher_bigger ? P.t_reg : P.reg) + (size_t)r * 4 * P.bins;
    const bool active = lane < P.bins;
    float d[4];
#pragma unroll
    for (int sd = 0; sd < 4; ++sd) d[sd] = warp_integral(active ? __ldg(zr + sd * P.bins + lane) : 0.f, active, lane);
    if (lane == 0) {
      const int l = level_of_row(P.pyr, r);
      const int q = r - P.pyr.off[l] * P.pyr.N;
      const int a = q % (P.pyr.h[l] * P.pyr.w[l]);
      const float cy = (float)(a / P.pyr.w[l]), cx = (float)(a % P.pyr.w[l]);
      P.score[r] = best;
      P.box[r] = make_float4(cx - d[0], cy - d[1], cx + d[2], cy + d[3]);
      P.flags[r] = 0;
    }
  }
}

// one block per level: `topk` rounds of (arg-max over unsuppressed rows) + (suppression sweep)
__global__ void __launch_bounds__(NMS_THREADS) gi_nms_kernel(Pyr pyr, const float* __restrict__ score,
                                                            const float4* __restrict__ box, unsigned char* flags,
                                                            float thr, int topk, int* __restrict__ lvl_rows,
                                                            int* __restrict__ lvl_cnt) {
  __shared__ unsigned long long s_key[NMS_THREADS / 32];
  __shared__ unsigned long long s_pick;
  const int l = blockIdx.x;
  const int r0 = pyr.off[l] * pyr.N, r1 = pyr.off[l + 1] * pyr.N;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int kept = 0;
  for (int it = 0; it < topk; ++it) {
    // key: score bits (scores are >= 0, so the IEEE bit pattern orders like the value) | inverted level-local index
    unsigned long long key = 0;
    for (int r = r0 + tid; r < r1; r += NMS_THREADS) {
      if (flags[r]) continue;
      const float sc = score[r];
      const unsigned long long k =
          ((unsigned long long)__float_as_uint(sc >= 0.f ? sc : 0.f) << 32) | (unsigned long long)(0xffffffffu - (unsigned)(r - r0));
      key = k > key ? k : key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(FULL, key, o);
      key = other > key ? other : key;
    }
    if (lane == 0) s_key[wid] = key;
    __syncthreads();
    if (wid == 0) {
      unsigned long long k = lane < NMS_THREADS / 32 ? s_key[lane] : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(FULL, k, o);
        k = other > k ? other : k;
      }
      if (lane == 0) s_pick = k;
    }
    __syncthreads();
    const unsigned long long pick = s_pick;
    if (pick == 0ull) break;   // every row of the level is suppressed (block-uniform)
    const int p = r0 + (int)(0xffffffffu - (unsigned)(pick & 0xffffffffull));
    if (tid == 0) {
      lvl_rows[l * topk + it] = p;
      flags[p] = 1;
    }
    ++kept;
    const float4 pb = box[p];
    const float iarea = __fmul_rn(__fsub_rn(pb.z, pb.x), __fsub_rn(pb.w, pb.y));
    for (int r = r0 + tid; r < r1; r += NMS_THREADS) {
      if (flags[r]) continue;
      const float4 b = box[r];
      const float w = fmaxf(0.f, __fsub_rn(fminf(pb.z, b.z), fmaxf(pb.x, b.x)));
      const float h = fmaxf(0.f, __fsub_rn(fminf(pb.w, b.w), fmaxf(pb.y, b.y)));
      const float inter = __fmul_rn(w, h);
      const float area = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
      const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, area), inter));
      if (ovr > thr) flags[r] = 1;
    }
    __syncthreads();   // flags[p] and the sweep are visible to the next round's arg-max
  }
  if (tid == 0) lvl_cnt[l] = kept;
}

// picks of the levels back to back + the counts vector of the fused loss with the imitation entries replaced
__global__ void gi_finalize_kernel(int L, int topk, const int* __restrict__ lvl_rows, const int* __restrict__ lvl_cnt,
                                   const int* __restrict__ counts_in, int* __restrict__ gi_rows,
                                   int* __restrict__ counts_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int i = 0; i < 4 + 3 * LDB_MAX_LEVELS; ++i) counts_out[i] = counts_in[i];
  int n = 0;
  for (int l = 0; l < L; ++l) {
    const int c = lvl_cnt[l];
    for (int i = 0; i < c; ++i) gi_rows[n++] = lvl_rows[l * topk + i];
    counts_out[4 + 3 * l + 2] = c;
  }
  for (int l = L; l < LDB_MAX_LEVELS; ++l) counts_out[4 + 3 * l + 2] = 0;
  counts_out[1] = n;
  for (int i = n; i < L * topk; ++i) gi_rows[i] = 0;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace
}  // namespace ldb

using namespace ldb;

extern "C" size_t ldb200_gi_region_workspace_bytes(const ldb_pyramid* pyr, int topk) {
  Pyr p;
  if (make_pyr(pyr, &p) || topk < 1) return 0;
  const size_t R = (size_t)p.off[p.L] * p.N;
  return align_up(R * sizeof(float4), 256) + align_up(R * sizeof(float), 256) + align_up(R, 256) +
         align_up((size_t)LDB_MAX_LEVELS * topk * sizeof(int), 256) + align_up(LDB_MAX_LEVELS * sizeof(int), 256);
}

static int gi_region_impl(const ldb_pyramid* pyr, const ldb_loss_cfg* cfg, const float* cls, const float* reg,
                          const float* t_cls, const float* t_reg, float iou_thr, int topk, const int32_t* counts_in,
                          int32_t* gi_rows, int32_t* counts_out, void* workspace, void* stream_, bool raw) {
  GiParams P;
  int e = make_pyr(pyr, &P.pyr);
  LDB_REQUIRE(e == 0, LDB_EINVAL, "gi_region: bad pyramid");
  LDB_REQUIRE(cfg && cls && reg && t_cls && t_reg && counts_in && gi_rows && counts_out && workspace, LDB_EINVAL,
              "gi_region: null pointer");
  LDB_REQUIRE(topk >= 1 && topk <= 64, LDB_EINVAL, "gi_region: topk=%d out of range [1, 64]", topk);
  LDB_REQUIRE(cfg->num_classes >= 1 && cfg->reg_max >= 1 && cfg->reg_max + 1 <= 32, LDB_ENOSUP,
              "gi_region: reg_max + 1 must fit one warp");
  P.C = cfg->num_classes;
  P.bins = cfg->reg_max + 1;
  P.cls = cls; P.t_cls = t_cls; P.reg = reg; P.t_reg = t_reg;
  const size_t R = (size_t)P.pyr.off[P.pyr.L] * P.pyr.N;
  LDB_REQUIRE(R < (1u << 31), LDB_ENOSUP, "gi_region: too many rows");
  unsigned char* w = (unsigned char*)workspace;
  LDB_REQUIRE(((uintptr_t)w & 15) == 0, LDB_EINVAL, "gi_region: workspace must be 16-byte aligned");
  P.box = (float4*)w;          w += align_up(R * sizeof(float4), 256);
  P.score = (float*)w;         w += align_up(R * sizeof(float), 256);
  P.flags = w;                 w += align_up(R, 256);
  int* lvl_rows = (int*)w;     w += align_up((size_t)LDB_MAX_LEVELS * topk * sizeof(int), 256);
  int* lvl_cnt = (int*)w;
  cudaStream_t stream = (cudaStream_t)stream_;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long blocks = ((long long)R + GI_WARPS - 1) / GI_WARPS;
  if (blocks > 8LL * sms) blocks = 8LL * sms;
  if (raw) gi_score_kernel<true><<<(int)blocks, GI_THREADS, 0, stream>>>(P);
  else gi_score_kernel<false><<<(int)blocks, GI_THREADS, 0, stream>>>(P);
  LDB_CUDA(cudaGetLastError());
  gi_nms_kernel<<<P.pyr.L, NMS_THREADS, 0, stream>>>(P.pyr, P.score, P.box, P.flags, iou_thr, topk, lvl_rows, lvl_cnt);
  LDB_CUDA(cudaGetLastError());
  gi_finalize_kernel<<<1, 32, 0, stream>>>(P.pyr.L, topk, lvl_rows, lvl_cnt, counts_in, gi_rows, counts_out);
  LDB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ldb200_gi_region(const ldb_pyramid* pyr, const ldb_loss_cfg* cfg, const float* cls, const float* reg,
                                const float* t_cls, const float* t_reg, float iou_thr, int topk,
                                const int32_t* counts_in, int32_t* gi_rows, int32_t* counts_out, void* workspace,
                                void* stream_) {
  return gi_region_impl(pyr, cfg, cls, reg, t_cls, t_reg, iou_thr, topk, counts_in, gi_rows, counts_out, workspace, stream_,
                        false);
}

// LDv2Head.get_gi_region (ld_gflv2.py): `cls` = student joint scores, `t_cls` = teacher cls_feat logits, no sigmoid.
extern "C" int ldb200_gi_region_v2(const ldb_pyramid* pyr, const ldb_loss_cfg* cfg, const float* score, const float* reg,
                                   const float* t_cls_feat, const float* t_reg, float iou_thr, int topk,
                                   const int32_t* counts_in, int32_t* gi_rows, int32_t* counts_out, void* workspace,
                                   void* stream_) {
  return gi_region_impl(pyr, cfg, score, reg, t_cls_feat, t_reg, iou_thr, topk, counts_in, gi_rows, counts_out, workspace,
                        stream_, true);
}
