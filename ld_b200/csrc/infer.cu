// Inference step after the training path: GFLHead._get_bboxes (gfl_head.py:354-451) +
// multiclass_nms with nms_cfg type 'nms' (core/post_processing/bbox_nms.py:70-193 -> mmcv.ops.batched_nms, not vendored
// in the reference: its class-offset trick and greedy IoU rule are restated from mmcv 1.2.x ops/nms.py).
//
// For every image of the batch, on row-space head outputs (cls [R, C], bbox_pred [R, 4*bins], fp32, levels back to back):
//   1. per anchor: max_c sigmoid(cls)                                                   infer_maxscore_kernel
//   2. per (image, level): the nms_pre anchors of highest max score (all, if the level has no more than nms_pre):
//      exact top-k by a 4-pass radix select over the score bits, ties at the cut broken by anchor index
//                                                                                       infer_select_kernel
//   3. per candidate: Integral (softmax over the bins . [0..bins-1]) * stride, distance2bbox from the anchor centre,
//      clip to the image shape, optional rescale; every (candidate, class) with sigmoid score > score_thr is appended to
//      the image's entry list                                                           infer_decode_kernel
//   4. per image: entries sorted by score (descending, ties by entry key), mmcv's batched_nms offset boxes
//      (box + class * (max coordinate + 1), fp32 as the reference computes them), greedy NMS (IoU > thr suppressed), stop
//      after max_per_img survivors                                                      infer_nms_kernel
// Integer outputs (labels, which anchors survive) are the reference's as long as no two scores tie exactly.
#include "common.cuh"

namespace ldb {

constexpr int INF_MAX_ENTRIES = 16384;    // valid (candidate, class) pairs per image the NMS stage can hold in smem

struct InferParams {
  Pyr pyr;
  int C, bins;
  int use_sigmoid;         // 1: cls holds logits (GFLHead); 0: probabilities (GFocalHead joint scores)
  const float* cls;        // [R, C]
  const float* reg;        // [R, 4*bins]
  const int* img_hw;       // [N, 2] (h, w) for clipping
  const float* scale;      // [N, 4] or null: boxes are divided by it (rescale=True)
  int nms_pre;
  float score_thr, iou_thr;
  int max_per_img;
  // workspace
  float* maxscore;         // [R]
  int* cand;               // [N, L, nms_pre] anchor index inside the level
  int* ncand;              // [N, L]
  float* boxes;            // [N, L*nms_pre, 4]
  unsigned long long* entries;   // [N, cap] (score bits << 32) | (cand slot << 8 | class)   (class < 256)
  int* nentries;           // [N]
  int cap;
  // outputs
  float* dets;             // [N, max_per_img, 5]
  int* labels;             // [N, max_per_img]
  int* counts;             // [N]: detections, or -1 when the image had more than INF_MAX_ENTRIES valid pairs
};

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256) infer_maxscore_kernel(InferParams P, int R) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < R; r += gridDim.x * wpb) {
    float m = -INFINITY;
    for (int c = lane; c < P.C; c += 32) m = fmaxf(m, __ldg(P.cls + (size_t)r * P.C + c));
    m = warp_max(m);
    if (lane == 0) P.maxscore[r] = P.use_sigmoid ? sigm(m) : m;   // sigmoid is monotone: max of sigmoids = sigmoid of the max
  }
}

// one block per (image, level): radix select of the k largest positive floats (bit pattern order = value order)
__global__ void __launch_bounds__(1024) infer_select_kernel(InferParams P) {
  const int l = blockIdx.x % P.pyr.L, n = blockIdx.x / P.pyr.L;
  const int Al = P.pyr.off[l + 1] - P.pyr.off[l];
  const float* sc = P.maxscore + (size_t)P.pyr.off[l] * P.pyr.N + (size_t)n * Al;
  int* out = P.cand + ((size_t)n * P.pyr.L + l) * P.nms_pre;
  __shared__ int hist[256];
  __shared__ unsigned s_prefix, s_need;
  __shared__ int s_count;
  const int k = P.nms_pre;
  if (k <= 0 || Al <= k) {
    for (int a = threadIdx.x; a < Al; a += blockDim.x) out[a] = a;
    if (threadIdx.x == 0) P.ncand[n * P.pyr.L + l] = Al;
    return;
  }
  // find the k-th largest key: fix 8 bits per pass from the top
  unsigned prefix = 0, mask = 0;
  int need = k;     // how many keys >= the final threshold are still to be found inside the current prefix bucket
  for (int pass = 3; pass >= 0; --pass) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int a = threadIdx.x; a < Al; a += blockDim.x) {
      const unsigned key = __float_as_uint(sc[a]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 255], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int acc = 0, b = 255;
      for (; b >= 0; --b) {
        if (acc + hist[b] >= need) break;
        acc += hist[b];
      }
      s_prefix = prefix | ((unsigned)b << (8 * pass));
      s_need = (unsigned)(need - acc);
    }
    __syncthreads();
    prefix = s_prefix;
    need = (int)s_need;
    mask |= 255u << (8 * pass);
    __syncthreads();
  }
  // prefix = key of the k-th largest; `need` keys equal to it are taken (lowest anchor index first), all larger keys too
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  for (int a = threadIdx.x; a < Al; a += blockDim.x)
    if (__float_as_uint(sc[a]) > prefix) out[atomicAdd(&s_count, 1)] = a;
  __syncthreads();
  if (threadIdx.x == 0) {     // ties at the cut: sequential, in index order (rare; deterministic)
    int c = s_count, taken = 0;
    for (int a = 0; a < Al && taken < need; ++a)
      if (__float_as_uint(sc[a]) == prefix) { out[c++] = a; ++taken; }
    P.ncand[n * P.pyr.L + l] = c;
  }
}

// one warp per candidate
__global__ void __launch_bounds__(256) infer_decode_kernel(InferParams P) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int L = P.pyr.L, N = P.pyr.N;
  const int total = N * L * P.nms_pre;
  for (int t = blockIdx.x * wpb + (threadIdx.x >> 5); t < total; t += gridDim.x * wpb) {
    const int slot = t % P.nms_pre, l = (t / P.nms_pre) % L, n = t / (P.nms_pre * L);
    if (slot >= P.ncand[n * L + l]) continue;
    const int a = P.cand[((size_t)n * L + l) * P.nms_pre + slot];
    const int Al = P.pyr.off[l + 1] - P.pyr.off[l];
    const size_t r = (size_t)P.pyr.off[l] * N + (size_t)n * Al + a;
    const float stride = (float)P.pyr.stride[l];
    // Integral (gfl_head.py:32-44): expectation of softmax over the bins, per side
    const bool active = lane < P.bins;
    float d[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const float z = active ? __ldg(P.reg + r * 4 * P.bins + s * P.bins + lane) : -INFINITY;
      const float m = warp_max(z);
      const float e = active ? expf(z - m) : 0.f;
      const float S = warp_sum(e);
      d[s] = warp_sum(active ? (e / S) * (float)lane : 0.f) * stride;
    }
    // anchor centre (gfl_head.py:185-194) of the square anchor centred at (x*s, y*s); distance2bbox + clamp
    const int ix = a % P.pyr.w[l], iy = a / P.pyr.w[l];
    const float cx = (float)ix * stride, cy = (float)iy * stride;
    const float ih = (float)P.img_hw[2 * n], iw = (float)P.img_hw[2 * n + 1];
    float x1 = fminf(fmaxf(cx - d[0], 0.f), iw), y1 = fminf(fmaxf(cy - d[1], 0.f), ih);
    float x2 = fminf(fmaxf(cx + d[2], 0.f), iw), y2 = fminf(fmaxf(cy + d[3], 0.f), ih);
    if (P.scale) {
      x1 /= P.scale[4 * n]; y1 /= P.scale[4 * n + 1]; x2 /= P.scale[4 * n + 2]; y2 /= P.scale[4 * n + 3];
    }
    const int cslot = l * P.nms_pre + slot;
    if (lane == 0) reinterpret_cast<float4*>(P.boxes)[(size_t)n * L * P.nms_pre + cslot] = make_float4(x1, y1, x2, y2);
    for (int c = lane; c < P.C; c += 32) {
      const float raw = __ldg(P.cls + r * P.C + c);
      const float s = P.use_sigmoid ? sigm(raw) : raw;
      if (s > P.score_thr) {
        const int e = atomicAdd(&P.nentries[n], 1);
        if (e < P.cap)
          P.entries[(size_t)n * P.cap + e] = ((unsigned long long)__float_as_uint(s) << 32) |
                                             (unsigned long long)(((unsigned)cslot << 8) | (unsigned)c);
      }
    }
  }
}

// one block per image: bitonic sort (descending) of the entries in shared memory, then greedy NMS
__global__ void __launch_bounds__(1024) infer_nms_kernel(InferParams P) {
  extern __shared__ unsigned long long sh_e[];     // [npow2]
  __shared__ float s_max;
  __shared__ int s_kept, s_cur;
  __shared__ float s_box[4];
  const int n = blockIdx.x;
  const int ne = P.nentries[n];
  float* dets = P.dets + (size_t)n * P.max_per_img * 5;
  int* labels = P.labels + (size_t)n * P.max_per_img;
  if (ne > INF_MAX_ENTRIES || ne > P.cap) {
    if (threadIdx.x == 0) P.counts[n] = -1;
    return;
  }
  int np2 = 1;
  while (np2 < ne) np2 <<= 1;
  for (int i = threadIdx.x; i < np2; i += blockDim.x) sh_e[i] = i < ne ? P.entries[(size_t)n * P.cap + i] : 0ull;
  __syncthreads();
  for (int k = 2; k <= np2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = sh_e[i], b = sh_e[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? (a < b) : (a > b)) { sh_e[i] = b; sh_e[ixj] = a; }
        }
      }
      __syncthreads();
    }
  // mmcv batched_nms: max_coordinate over the valid boxes, offsets = class * (max_coordinate + 1)
  const float4* boxes = reinterpret_cast<const float4*>(P.boxes) + (size_t)n * P.pyr.L * P.nms_pre;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < ne; i += blockDim.x) {
    const float4 b = boxes[(unsigned)(sh_e[i] & 0xffffffffu) >> 8];
    m = fmaxf(m, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
  }
  m = warp_max(m);
  if (threadIdx.x == 0) { s_max = -INFINITY; s_kept = 0; s_cur = 0; }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(&s_max), __float_as_int(m));   // coordinates are >= 0
  __syncthreads();
  const float off1 = s_max + 1.f;
  // greedy: sh_e entries with bit 63.. cleared stay; a suppressed entry gets its score bits zeroed (score > 0 always)
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) {
      int i = s_cur;
      while (i < ne && (sh_e[i] >> 32) == 0ull) ++i;
      s_cur = i;
      if (i < ne && s_kept < P.max_per_img) {
        const unsigned pay = (unsigned)(sh_e[i] & 0xffffffffu);
        const int c = pay & 255;
        const float4 b = boxes[pay >> 8];
        const float o = (float)c * off1;
        s_box[0] = b.x + o; s_box[1] = b.y + o; s_box[2] = b.z + o; s_box[3] = b.w + o;
        dets[s_kept * 5 + 0] = b.x; dets[s_kept * 5 + 1] = b.y; dets[s_kept * 5 + 2] = b.z; dets[s_kept * 5 + 3] = b.w;
        dets[s_kept * 5 + 4] = __uint_as_float((unsigned)(sh_e[i] >> 32));
        labels[s_kept] = c;
        ++s_kept;
      } else {
        s_cur = ne;
      }
    }
    __syncthreads();
    const int cur = s_cur;
    if (cur >= ne) break;
    const float ax1 = s_box[0], ay1 = s_box[1], ax2 = s_box[2], ay2 = s_box[3];
    const float aarea = (ax2 - ax1) * (ay2 - ay1);
    for (int j = cur + 1 + threadIdx.x; j < ne; j += blockDim.x) {
      const unsigned long long e = sh_e[j];
      if ((e >> 32) == 0ull) continue;
      const unsigned pay = (unsigned)(e & 0xffffffffu);
      const float4 b = boxes[pay >> 8];
      const float o = (float)(pay & 255) * off1;
      const float bx1 = b.x + o, by1 = b.y + o, bx2 = b.z + o, by2 = b.w + o;
      const float w = fmaxf(fminf(ax2, bx2) - fmaxf(ax1, bx1), 0.f), h = fmaxf(fminf(ay2, by2) - fmaxf(ay1, by1), 0.f);
      const float inter = w * h;
      const float iou = inter / (aarea + (bx2 - bx1) * (by2 - by1) - inter);
      if (iou > P.iou_thr) sh_e[j] = e & 0xffffffffull;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_cur = cur + 1;
  }
  if (threadIdx.x == 0) P.counts[n] = s_kept;
}

}  // namespace ldb

using namespace ldb;

static size_t inf_align(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" size_t ldb200_get_bboxes_workspace_bytes(const ldb_pyramid* pyr, int num_classes, int nms_pre) {
  Pyr p;
  if (make_pyr(pyr, &p) || nms_pre <= 0) return 0;
  const size_t N = p.N, L = p.L, R = (size_t)p.off[p.L] * N;
  const size_t cap = L * nms_pre * (size_t)num_classes;
  return inf_align(R * 4) + inf_align(N * L * nms_pre * 4) + inf_align(N * L * 4) + inf_align(N * L * nms_pre * 16) +
         inf_align(N * (cap < (size_t)INF_MAX_ENTRIES * 4 ? cap : (size_t)INF_MAX_ENTRIES * 4) * 8) + inf_align(N * 4);
}

extern "C" int ldb200_gfl_get_bboxes(const ldb_pyramid* pyr, const float* cls, const float* bbox_pred, int num_classes,
                                     int bins, int use_sigmoid, const int32_t* img_hw, const float* scale_factors, int nms_pre,
                                     float score_thr, float iou_thr, int max_per_img, float* dets, int32_t* labels,
                                     int32_t* counts, void* workspace, void* stream_) {
  InferParams P;
  memset(&P, 0, sizeof(P));
  LDB_REQUIRE(make_pyr(pyr, &P.pyr) == 0, LDB_EINVAL, "get_bboxes: bad pyramid");
  LDB_REQUIRE(cls && bbox_pred && img_hw && dets && labels && counts && workspace, LDB_EINVAL, "get_bboxes: null pointer");
  LDB_REQUIRE(num_classes >= 1 && num_classes <= 256 && bins >= 2 && bins <= 32, LDB_EINVAL, "get_bboxes: classes / bins");
  LDB_REQUIRE(nms_pre >= 1 && max_per_img >= 1 && max_per_img <= 4096, LDB_EINVAL, "get_bboxes: nms_pre / max_per_img");
  LDB_REQUIRE((size_t)P.pyr.L * nms_pre < (1u << 24), LDB_EINVAL, "get_bboxes: too many candidates per image");
  const size_t N = P.pyr.N, L = P.pyr.L, R = (size_t)P.pyr.off[P.pyr.L] * N;
  size_t cap = L * nms_pre * (size_t)num_classes;
  if (cap > (size_t)INF_MAX_ENTRIES * 4) cap = (size_t)INF_MAX_ENTRIES * 4;
  P.C = num_classes; P.bins = bins; P.use_sigmoid = use_sigmoid; P.cls = cls; P.reg = bbox_pred; P.img_hw = img_hw; P.scale = scale_factors;
  P.nms_pre = nms_pre; P.score_thr = score_thr; P.iou_thr = iou_thr; P.max_per_img = max_per_img;
  P.dets = dets; P.labels = labels; P.counts = counts; P.cap = (int)cap;
  unsigned char* w = (unsigned char*)workspace;
  P.maxscore = (float*)w; w += inf_align(R * 4);
  P.cand = (int*)w; w += inf_align(N * L * nms_pre * 4);
  P.ncand = (int*)w; w += inf_align(N * L * 4);
  P.boxes = (float*)w; w += inf_align(N * L * nms_pre * 16);
  P.entries = (unsigned long long*)w; w += inf_align(N * cap * 8);
  P.nentries = (int*)w;
  cudaStream_t s = (cudaStream_t)stream_;
  LDB_CUDA(cudaMemsetAsync(P.nentries, 0, N * 4, s));
  LDB_CUDA(cudaMemsetAsync(dets, 0, N * max_per_img * 5 * sizeof(float), s));
  LDB_CUDA(cudaMemsetAsync(labels, 0, N * max_per_img * sizeof(int32_t), s));
  long long blocks = ((long long)R + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  infer_maxscore_kernel<<<(unsigned)blocks, 256, 0, s>>>(P, (int)R);
  LDB_CUDA(cudaGetLastError());
  infer_select_kernel<<<(unsigned)(N * L), 1024, 0, s>>>(P);
  LDB_CUDA(cudaGetLastError());
  blocks = ((long long)(N * L * nms_pre) + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  infer_decode_kernel<<<(unsigned)blocks, 256, 0, s>>>(P);
  LDB_CUDA(cudaGetLastError());
  static bool attr = false;
  const size_t smem = (size_t)INF_MAX_ENTRIES * 8;
  if (!attr) {
    LDB_CUDA(cudaFuncSetAttribute(infer_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  infer_nms_kernel<<<(unsigned)N, 1024, smem, s>>>(P);
  LDB_CUDA(cudaGetLastError());
  return 0;
}
