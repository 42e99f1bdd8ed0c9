// LD target generation on the GPU: ATSS assignment + Valuable Localization Region + imitation
// region for a whole batch in three launches (no host sync, no A x G temporaries).
//
// Replaces (reference, paths under /root/reference/mmdet):
//   models/dense_heads/ld_head.py:377-577   LDHead.get_targets / _get_target_single
//   core/anchor/anchor_generator.py:233-328 grid anchors + valid flags (computed analytically)
//   core/bbox/assigners/atss_assigner.py:33-181   ATSSAssigner.assign
//   core/bbox/assigners/atss_assigner.py:183-298  ATSSAssigner.get_vlr_region
//   models/dense_heads/ld_head.py:580-611   LDHead.get_im_region
//
// Bit-exactness: all float arithmetic that feeds a comparison is written
// with the explicit round-to-nearest intrinsics (__fadd_rn/__fmul_rn/__fdiv_rn/__fsqrt_rn) in the
// reference's operation order, so nvcc can neither contract to FMA nor re-associate.  The ATSS
// threshold mean is summed in the order torch's CPU `sum` kernel uses (see thr_mean()) and the
// unbiased std with Welford in double like torch's CPU `std`.
#include "common.cuh"

namespace ldb {

constexpr int KMAX = 16;        // max supported topk
constexpr int TGT_THREADS = 256;
constexpr float NEG_INF_F = -100000000.0f;  // atss_assigner.py:66 INF = 100000000

__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

struct Box {
  float x1, y1, x2, y2;
};

// iou2d_calculator.py:111-161 (mode 'iou'): (area1 + area2) - overlap, max(union, eps), divide.
__device__ __forceinline__ float iou_pair(const Box& a, float area_a, const Box& b, float area_b) {
  float w = fmaxf(fsub(fminf(a.x2, b.x2), fmaxf(a.x1, b.x1)), 0.f);
  float h = fmaxf(fsub(fminf(a.y2, b.y2), fmaxf(a.y1, b.y1)), 0.f);
  float ov = fmul(w, h);
  float uni = fmaxf(fsub(fadd(area_a, area_b), ov), 1e-6f);
  return fdiv(ov, uni);
}

// iou2d_calculator.py:146-188 (mode 'diou', a fork addition): base term is IoF = ov/area(a).
__device__ __forceinline__ float diou_pair(const Box& a, float area_a, const Box& b) {
  float w = fmaxf(fsub(fminf(a.x2, b.x2), fmaxf(a.x1, b.x1)), 0.f);
  float h = fmaxf(fsub(fminf(a.y2, b.y2), fmaxf(a.y1, b.y1)), 0.f);
  float ov = fmul(w, h);
  float base = fdiv(ov, fmaxf(area_a, 1e-6f));
  float dx = fsub(fadd(b.x1, b.x2), fadd(a.x1, a.x2));
  float dy = fsub(fadd(b.y1, b.y2), fadd(a.y1, a.y2));
  float rho2 = fadd(fdiv(fmul(dx, dx), 4.f), fdiv(fmul(dy, dy), 4.f));
  float ew = fmaxf(fsub(fmaxf(a.x2, b.x2), fminf(a.x1, b.x1)), 0.f);
  float eh = fmaxf(fsub(fmaxf(a.y2, b.y2), fminf(a.y1, b.y1)), 0.f);
  float c2 = fmaxf(fadd(fmul(ew, ew), fmul(eh, eh)), 1e-6f);
  return fsub(base, fdiv(rho2, c2));
}

__device__ __forceinline__ Box anchor_box(int ix, int iy, int stride) {
  float cx = (float)(ix * stride), cy = (float)(iy * stride), half = 4.f * (float)stride;
  Box b;
  b.x1 = cx - half; b.y1 = cy - half; b.x2 = cx + half; b.y2 = cy + half;  // exact integers
  return b;
}

// ---- torch CPU float `sum` orders (aten/src/ATen/native/cpu/SumKernel.cpp), Vec<float>::size()=16.
// multi_row_sum: 4-level cascade with 16-row chunks.
__device__ float sum_cascade(const float* v, int n, int stride) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int lp = 4;
  {
    int lg = 0;
    while ((1 << lg) < n) ++lg;
    if (lg / 4 > lp) lp = lg / 4;
  }
  const int step = 1 << lp, mask = step - 1;
  int i = 0;
  for (; i + step <= n;) {
    for (int j = 0; j < step; ++j, ++i) acc[0] = fadd(acc[0], v[i * stride]);
    for (int j = 1; j < 4; ++j) {
      acc[j] = fadd(acc[j], acc[j - 1]);
      acc[j - 1] = 0.f;
      if ((i & (mask << (j * lp))) != 0) break;
    }
  }
  for (; i < n; ++i) acc[0] = fadd(acc[0], v[i * stride]);
  for (int j = 1; j < 4; ++j) acc[0] = fadd(acc[0], acc[j]);
  return acc[0];
}
// row_sum: view as (n/4, 4), cascade per column, tail into column 0, then fold columns.
__device__ float sum_rowsum4(const float* v, int n) {
  const int q = n / 4;
  float p[4];
  for (int k = 0; k < 4; ++k) p[k] = sum_cascade(v + k, q, 4);
  for (int i = q * 4; i < n; ++i) p[0] = fadd(p[0], v[i]);
  for (int k = 1; k < 4; ++k) p[0] = fadd(p[0], p[k]);
  return p[0];
}
// vectorized_inner_sum with 16 lanes: lane-wise row_sum over n/16 vector rows, scalar tail first.
__device__ float sum_inner16(const float* v, int n) {
  const int vrows = n / 16;
  const int q = vrows / 4;
  float fin = 0.f;
  for (int k = vrows * 16; k < n; ++k) fin = fadd(fin, v[k]);
  for (int lane = 0; lane < 16; ++lane) {
    float p[4];
    for (int k = 0; k < 4; ++k) p[k] = sum_cascade(v + lane + 16 * k, q, 64);
    for (int i = q * 4; i < vrows; ++i) p[0] = fadd(p[0], v[lane + 16 * i]);
    for (int k = 1; k < 4; ++k) p[0] = fadd(p[0], p[k]);
    fin = fadd(fin, p[0]);
  }
  return fin;
}
// candidate_overlaps.mean(0) for column g of G columns, n rows (atss_assigner.py:132).
__device__ float thr_mean(const float* v, int n, int g, int G) {
  float s;
  if (G == 1)
    s = (n >= 16) ? sum_inner16(v, n) : sum_rowsum4(v, n);
  else if (G < 16)
    s = (g < 4 * (G / 4)) ? sum_cascade(v, n, 1) : sum_rowsum4(v, n);
  else
    s = (g < 64 * (G / 64)) ? sum_cascade(v, n, 1) : sum_rowsum4(v, n);
  return fdiv(s, (float)n);
}
// candidate_overlaps.std(0): Welford in double, unbiased (ReduceOpsKernel.cpp std_var_kernel_impl)
__device__ float thr_std(const float* v, int n) {
  double mean = 0.0, m2 = 0.0;
  for (int i = 0; i < n; ++i) {
    double x = (double)v[i];
    double delta = x - mean;
    mean = mean + delta / (double)(i + 1);
    m2 = m2 + delta * (x - mean);
  }
  double div = (double)n - 1.0;
  double var = m2 / (div > 0.0 ? div : 0.0);
  return (float)sqrt(var);
}

struct TgtWork {            // workspace carve-up
  unsigned long long* key;  // [N*A] (iou_bits << 32) | (~gt_idx), 0 = unassigned
  float* thr;               // [sumG]
  float* colmax;            // [sumG]
  int* pos_per_img;         // [N]
};

// ---------------------------------------------------------------------------------------------
// K1: one block per (image, gt).  Per-level top-k nearest anchor centres (atss_assigner.py:115-127),
// IoU threshold mean+std (:131-134), positive test (:136-154) and arg-max merge across gts via a
// packed atomicMax (:158-170); also the per-gt max IoU over anchors for 'finegrained'
// (ld_head.py:595).
__global__ void __launch_bounds__(TGT_THREADS)
atss_gt_kernel(Pyr pyr, const float* __restrict__ gt_boxes, const int* __restrict__ gt_offsets,
               const int* __restrict__ pad_hw, int topk, TgtWork ws) {
  const int j = blockIdx.x;  // global gt index
  int n = 0;
  while (n + 1 < pyr.N && j >= gt_offsets[n + 1]) ++n;
  const int g = j - gt_offsets[n];
  const int G = gt_offsets[n + 1] - gt_offsets[n];
  const int A = pyr.off[pyr.L];
  Box gt;
  gt.x1 = gt_boxes[4 * j + 0]; gt.y1 = gt_boxes[4 * j + 1];
  gt.x2 = gt_boxes[4 * j + 2]; gt.y2 = gt_boxes[4 * j + 3];
  const float gcx = fdiv(fadd(gt.x1, gt.x2), 2.0f), gcy = fdiv(fadd(gt.y1, gt.y2), 2.0f);
  const float area_g = fmul(fsub(gt.x2, gt.x1), fsub(gt.y2, gt.y1));
  const int pad_h = pad_hw[2 * n], pad_w = pad_hw[2 * n + 1];

  __shared__ unsigned long long s_red[TGT_THREADS / 32];
  __shared__ unsigned long long s_win;
  __shared__ int s_cand[LDB_MAX_LEVELS * KMAX];   // level-local anchor index, -1 = none
  __shared__ float s_iou[LDB_MAX_LEVELS * KMAX];
  __shared__ float s_thr;
  __shared__ float s_fred[TGT_THREADS / 32];
  __shared__ int s_ncand;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float my_max = 0.f;  // IoU >= 0
  if (tid == 0) s_ncand = 0;
  __syncthreads();

  for (int l = 0; l < pyr.L; ++l) {
    const int W = pyr.w[l], H = pyr.h[l], s = pyr.stride[l];
    const int vh = min((pad_h + s - 1) / s, H), vw = min((pad_w + s - 1) / s, W);
    unsigned long long best[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) best[k] = ~0ull;
    const float area_a = fmul(8.f * s, 8.f * s);
    for (int a = tid; a < H * W; a += TGT_THREADS) {
      const int iy = a / W, ix = a - iy * W;
      if (iy >= vh || ix >= vw) continue;
      // distance between centres: sqrt(dx^2 + dy^2) (atss_assigner.py:102-103)
      float dx = fsub((float)(ix * s), gcx), dy = fsub((float)(iy * s), gcy);
      float d = __fsqrt_rn(fadd(fmul(dx, dx), fmul(dy, dy)));
      unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)a;
      if (key < best[KMAX - 1]) {  // sorted insert, lowest (d, index) first; static indices only
        best[KMAX - 1] = key;
#pragma unroll
        for (int k = KMAX - 1; k > 0; --k) {
          if (best[k] < best[k - 1]) {
            unsigned long long tmp = best[k];
            best[k] = best[k - 1];
            best[k - 1] = tmp;
          }
        }
      }
      Box ab = anchor_box(ix, iy, s);
      my_max = fmaxf(my_max, iou_pair(ab, area_a, gt, area_g));
    }
    // block merge: topk rounds of arg-min over the per-thread heads
    const int kk = min(topk, vh * vw);
    for (int r = 0; r < kk; ++r) {
      unsigned long long m = warp_min_u64(best[0]);
      if (lane == 0) s_red[wid] = m;
      __syncthreads();
      if (wid == 0) {
        unsigned long long t = (lane < TGT_THREADS / 32) ? s_red[lane] : ~0ull;
        t = warp_min_u64(t);
        if (lane == 0) s_win = t;
      }
      __syncthreads();
      const unsigned long long win = s_win;
      if (best[0] == win) {  // unique owner (the key holds the anchor index)
#pragma unroll
        for (int m = 0; m < KMAX - 1; ++m) best[m] = best[m + 1];
        best[KMAX - 1] = ~0ull;
        s_cand[s_ncand + r] = (l << 28) | (int)(win & 0xffffffffu);
      }
      __syncthreads();
    }
    if (tid == 0) s_ncand += kk;
    __syncthreads();
  }
  const int ncand = s_ncand;
  // IoU of each candidate with this gt (atss_assigner.py:131)
  if (tid < ncand) {
    const int enc = s_cand[tid];
    const int l = enc >> 28, a = enc & 0x0fffffff;
    const int W = pyr.w[l], s = pyr.stride[l];
    const int iy = a / W, ix = a - iy * W;
    Box ab = anchor_box(ix, iy, s);
    s_iou[tid] = iou_pair(ab, fmul(8.f * s, 8.f * s), gt, area_g);
  }
  // per-gt max IoU over all valid anchors
  float wm = warp_max(my_max);
  if (lane == 0) s_fred[wid] = wm;
  __syncthreads();
  if (tid == 0) {
    float m = 0.f;
    for (int i = 0; i < TGT_THREADS / 32; ++i) m = fmaxf(m, s_fred[i]);
    ws.colmax[j] = m;
    float thr = fadd(thr_mean(s_iou, ncand, g, G), thr_std(s_iou, ncand));
    s_thr = thr;
    ws.thr[j] = thr;
  }
  __syncthreads();
  if (tid < ncand) {
    const float thr = s_thr;
    const int enc = s_cand[tid];
    const int l = enc >> 28, a = enc & 0x0fffffff;
    const int W = pyr.w[l], s = pyr.stride[l];
    const int iy = a / W, ix = a - iy * W;
    const float cx = (float)(ix * s), cy = (float)(iy * s);
    const float iou = s_iou[tid];
    const float mn = fminf(fminf(fsub(cx, gt.x1), fsub(cy, gt.y1)),
                           fminf(fsub(gt.x2, cx), fsub(gt.y2, cy)));
    if (iou >= thr && mn > 0.01f) {
      unsigned long long key = ((unsigned long long)__float_as_uint(iou) << 32) |
                               (unsigned)(0xffffffffu - (unsigned)g);
      atomicMax(&ws.key[(size_t)n * A + pyr.off[l] + a], key);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K2: one thread per (image, anchor): VLR band + imitation region over the image's gts, decode the
// assignment, write every per-anchor target in row space and append to the compact row lists.
__global__ void __launch_bounds__(256)
anchor_targets_kernel(Pyr pyr, const float* __restrict__ gt_boxes, const int* __restrict__ gt_labels,
                      const int* __restrict__ gt_offsets, const int* __restrict__ pad_hw, int im_mode,
                      int num_classes, TgtWork ws, int* __restrict__ labels,
                      float* __restrict__ label_w, float* __restrict__ bbox_t, float* __restrict__ vlr_out,
                      float* __restrict__ im_out, int* __restrict__ gt_inds, int* __restrict__ sparse_rows,
                      int* __restrict__ fg_rows, int* __restrict__ counts) {
  const int A = pyr.off[pyr.L];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)pyr.N * A) return;
  const int n = (int)(t / A), ag = (int)(t - (long long)n * A);
  const int l = level_of_anchor(pyr, ag);
  const int a = ag - pyr.off[l];
  const int W = pyr.w[l], H = pyr.h[l], s = pyr.stride[l];
  const int iy = a / W, ix = a - iy * W;
  const int row = pyr.off[l] * pyr.N + n * (H * W) + a;
  const int pad_h = pad_hw[2 * n], pad_w = pad_hw[2 * n + 1];
  const bool valid = iy < min((pad_h + s - 1) / s, H) && ix < min((pad_w + s - 1) / s, W);

  int lab = num_classes, gi = 0;
  float lw = 0.f, v = 0.f, im = 0.f;
  float4 bt = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid) {
    lw = 1.f;
    const Box ab = anchor_box(ix, iy, s);
    const float area_a = fmul(8.f * s, 8.f * s);
    const float cx = (float)(ix * s), cy = (float)(iy * s);
    const int g0 = gt_offsets[n], g1 = gt_offsets[n + 1];
    float vm = NEG_INF_F;
    bool fg = false;
    for (int j = g0; j < g1; ++j) {
      Box gt;
      gt.x1 = __ldg(gt_boxes + 4 * j); gt.y1 = __ldg(gt_boxes + 4 * j + 1);
      gt.x2 = __ldg(gt_boxes + 4 * j + 2); gt.y2 = __ldg(gt_boxes + 4 * j + 3);
      const float area_g = fmul(fsub(gt.x2, gt.x1), fsub(gt.y2, gt.y1));
      const float iou = iou_pair(ab, area_a, gt, area_g);
      const float thr = __ldg(ws.thr + j);
      const float di = diou_pair(ab, area_a, gt);
      // atss_assigner.py:272-273
      if (di < thr && di >= fmul(0.25f, thr)) vm = fmaxf(vm, iou);
      if (im_mode == LDB_IM_FINEGRAINED) {
        fg = fg || (iou > fmul(0.5f, __ldg(ws.colmax + j)));      // ld_head.py:595
      } else {
        fg = fg || (cx > gt.x1 && cx < gt.x2 && cy > gt.y1 && cy < gt.y2);  // ld_head.py:599-606
      }
    }
    v = (vm != NEG_INF_F) ? vm : 0.f;
    im = fg ? 1.f : 0.f;
    const unsigned long long key = ws.key[(size_t)n * A + ag];
    if (key != 0ull) {
      const int g = (int)(0xffffffffu - (unsigned)(key & 0xffffffffu));
      gi = g + 1;
      const int j = g0 + g;
      lab = __ldg(gt_labels + j);
      bt = make_float4(__ldg(gt_boxes + 4 * j), __ldg(gt_boxes + 4 * j + 1),
                       __ldg(gt_boxes + 4 * j + 2), __ldg(gt_boxes + 4 * j + 3));
      atomicAdd(ws.pos_per_img + n, 1);
      atomicAdd(counts + 4 + 3 * l + 0, 1);
    }
    if (v > 0.f) atomicAdd(counts + 4 + 3 * l + 1, 1);
    if (im > 0.f) {
      atomicAdd(counts + 4 + 3 * l + 2, 1);
      fg_rows[atomicAdd(counts + 1, 1)] = row;
    }
    if (gi > 0 || v > 0.f) sparse_rows[atomicAdd(counts + 0, 1)] = row;
  }
  labels[row] = lab;
  label_w[row] = lw;
  reinterpret_cast<float4*>(bbox_t)[row] = bt;
  vlr_out[row] = v;
  im_out[row] = im;
  if (gt_inds) gt_inds[row] = gi;
}

__global__ void finalize_counts_kernel(int N, const int* __restrict__ pos_per_img, int* counts) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int tot = 0, raw = 0;
    for (int n = 0; n < N; ++n) {
      raw += pos_per_img[n];
      tot += max(pos_per_img[n], 1);   // ld_head.py:429
    }
    counts[2] = tot;
    counts[3] = raw;
  }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace ldb

using namespace ldb;

extern "C" size_t ldb200_targets_workspace_bytes(const ldb_pyramid* p, int sum_gt) {
  Pyr pyr;
  if (make_pyr(p, &pyr)) return 0;
  size_t A = pyr.off[pyr.L];
  size_t b = align_up((size_t)pyr.N * A * 8, 256);
  b += align_up((size_t)sum_gt * 4, 256) * 2;
  b += align_up((size_t)pyr.N * 4, 256);
  return b;
}

extern "C" int ldb200_ld_targets(const ldb_pyramid* p, const float* gt_boxes, const int32_t* gt_labels,
                                 const int32_t* gt_offsets, int sum_gt, const int32_t* pad_hw, int topk,
                                 int im_mode, int num_classes, int32_t* labels, float* label_weights,
                                 float* bbox_targets, float* vlr, float* im_region, int32_t* gt_inds,
                                 int32_t* sparse_rows, int32_t* fg_rows, int32_t* counts,
                                 void* workspace, void* stream_) {
  Pyr pyr;
  LDB_REQUIRE(make_pyr(p, &pyr) == 0, LDB_EINVAL, "ld_targets: bad pyramid");
  LDB_REQUIRE(topk >= 1 && topk <= KMAX, LDB_ENOSUP, "ld_targets: topk %d not in [1,%d]", topk, KMAX);
  LDB_REQUIRE(sum_gt >= pyr.N, LDB_EINVAL, "ld_targets: every image needs >= 1 gt (sum_gt=%d, N=%d)", sum_gt, pyr.N);
  LDB_REQUIRE(gt_boxes && gt_labels && gt_offsets && pad_hw && labels && label_weights && bbox_targets &&
              vlr && im_region && sparse_rows && fg_rows && counts && workspace, LDB_EINVAL,
              "ld_targets: null pointer");
  LDB_REQUIRE(im_mode == LDB_IM_FINEGRAINED || im_mode == LDB_IM_CENTER, LDB_EINVAL, "ld_targets: im_mode");
  for (int l = 0; l < pyr.L; ++l)
    LDB_REQUIRE(pyr.h[l] * pyr.w[l] < (1 << 28), LDB_ENOSUP, "ld_targets: level too large");
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t A = pyr.off[pyr.L];
  char* w = (char*)workspace;
  TgtWork ws;
  ws.key = (unsigned long long*)w;  w += align_up((size_t)pyr.N * A * 8, 256);
  ws.thr = (float*)w;               w += align_up((size_t)sum_gt * 4, 256);
  ws.colmax = (float*)w;            w += align_up((size_t)sum_gt * 4, 256);
  ws.pos_per_img = (int*)w;
  LDB_CUDA(cudaMemsetAsync(ws.key, 0, (size_t)pyr.N * A * 8, stream));
  LDB_CUDA(cudaMemsetAsync(ws.pos_per_img, 0, (size_t)pyr.N * 4, stream));
  LDB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (4 + 3 * LDB_MAX_LEVELS), stream));
  atss_gt_kernel<<<sum_gt, TGT_THREADS, 0, stream>>>(pyr, gt_boxes, gt_offsets, pad_hw, topk, ws);
  LDB_CUDA(cudaGetLastError());
  const long long total = (long long)pyr.N * A;
  anchor_targets_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      pyr, gt_boxes, gt_labels, gt_offsets, pad_hw, im_mode, num_classes, ws, labels, label_weights,
      bbox_targets, vlr, im_region, gt_inds, sparse_rows, fg_rows, counts);
  LDB_CUDA(cudaGetLastError());
  finalize_counts_kernel<<<1, 32, 0, stream>>>(pyr.N, ws.pos_per_img, counts);
  LDB_CUDA(cudaGetLastError());
  return 0;
}
