// HBM-bound glue kernels around the tensor-core convolutions (NHWC bf16 activations):
// stem im2col, max-pool, GroupNorm(+ReLU) forward/backward, nearest-upsample-add backward,
// ReLU-mask, per-channel column sums (bias / BN-beta gradients) and the flat SGD update.
//
// Replaces (reference, under /root/reference/mmdet): backbones/resnet.py:626-630 (stem BN/ReLU/
// maxpool after the 7x7 conv), dense_heads/gfl_head.py:177-180 via mmcv ConvModule (GroupNorm(32) +
// ReLU, torch native_group_norm / native_group_norm_backward), necks/fpn.py:182-191 backward,
// torch.optim.SGD (configs/ld/ld_r50_gflv1_r101_fpn_coco_1x.py:53).
#include "common.cuh"

namespace ldb {

__device__ __forceinline__ void unpack8(const uint4& q, float v[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = __bfloat1622float2(h[e]);
    v[2 * e] = f.x;
    v[2 * e + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float v[8]) {
  uint4 q;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
  return q;
}

// 8 consecutive elements of a bf16 (one 16-byte access) or fp32 (two) tensor; `p` is 8-element aligned.
// fp32 activations are the "parity" mode (split-bf16 convolutions): same kernels, wider elements.
template <typename T> __device__ __forceinline__ void ld8(const T* p, float v[8]);
template <> __device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float v[8]) {
  unpack8(__ldg(reinterpret_cast<const uint4*>(p)), v);
}
template <> __device__ __forceinline__ void ld8<float>(const float* p, float v[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <typename T> __device__ __forceinline__ void st8(T* p, const float v[8]);
template <> __device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, const float v[8]) {
  *reinterpret_cast<uint4*>(p) = pack8(v);
}
template <> __device__ __forceinline__ void st8<float>(float* p, const float v[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// ---------------------------------------------------------------------------------------------
// Stem: 7x7 stride-2 pad-3 conv on a 3-channel NCHW fp32 image as a GEMM: A[pixel, (r*7+s)*3+c],
// K = 147 zero-padded to 192 (3 x 64) so the tcgen05 kernel can consume it in flat mode.
// One block = 64 consecutive output pixels of one output row: the 3 x 7 x 133 input patch is staged in
// shared memory with coalesced loads, every thread then emits 16-byte (8-column) pieces of the rows.
constexpr int IM_PX = 64;
constexpr int IM_COLS = 2 * IM_PX + 5;   // input columns touched by 64 outputs
template <typename T>
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const float* __restrict__ img, T* __restrict__ out, int N, int H, int W, int HO, int WO) {
  __shared__ float tile[3][7][IM_COLS + 1];
  const int xo0 = blockIdx.x * IM_PX, yo = blockIdx.y, n = blockIdx.z;
  const int xi0 = xo0 * 2 - 3, yi0 = yo * 2 - 3;
  for (int i = threadIdx.x; i < 3 * 7 * IM_COLS; i += 256) {
    const int xx = i % IM_COLS, r = (i / IM_COLS) % 7, c = i / (IM_COLS * 7);
    const int y = yi0 + r, x = xi0 + xx;
    float v = 0.f;
    if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg(img + (((size_t)n * 3 + c) * H + y) * W + x);
    tile[c][r][xx] = v;
  }
  __syncthreads();
  const int npx = min(IM_PX, WO - xo0);
  T* orow = out + (((size_t)n * HO + yo) * WO + xo0) * 192;
  for (int j = threadIdx.x; j < npx * 24; j += 256) {
    const int px = j / 24, grp = j % 24;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = grp * 8 + e;
      v[e] = (k < 147) ? tile[k % 3][k / 21][px * 2 + (k / 3) % 7] : 0.f;
    }
    st8<T>(orow + (size_t)j * 8, v);
  }
}

// 3x3 stride-2 pad-1 max pool, NHWC, 8 channels per thread.
template <typename T>
__global__ void maxpool3x3s2_kernel(const T* __restrict__ in, T* __restrict__ out, int N,
                                    int H, int W, int C, int HO, int WO) {
  const int cg = C / 8;
  const long long total = (long long)N * HO * WO * cg;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % cg);
    const long long pix = t / cg;
    const int xo = (int)(pix % WO), yo = (int)((pix / WO) % HO), n = (int)(pix / ((long long)WO * HO));
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
    for (int r = 0; r < 3; ++r) {
      const int y = yo * 2 - 1 + r;
      if (y < 0 || y >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int x = xo * 2 - 1 + s;
        if (x < 0 || x >= W) continue;
        float v[8];
        ld8<T>(in + (((size_t)n * H + y) * W + x) * C + g * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], v[e]);
      }
    }
    st8<T>(out + (size_t)t * 8, m);
  }
}

// ---------------------------------------------------------------------------------------------
// GroupNorm with 8 channels per group (256 ch / 32 groups).  x: [N, HW, C] bf16.  One thread owns one group of
// one pixel per step (a 16-byte load); a block is 32 groups x 8 pixel lanes and walks GN_PIX pixels of ONE image, so
// mean / rstd / gamma / beta are loop invariants and a warp always reads one contiguous 512-byte pixel row.
// stats: [N, G, 2] fp32 (sum, sum of squares), accumulated with one atomic pair per (block, group).
constexpr int GN_CPG = 8;
constexpr int GN_PIX = 256;       // pixels per block
constexpr int GN_THREADS = 256;

// Up to LDB_GN_MAX_SEG tensors (the pyramid levels of one level-shared tower layer) per launch: the 1-D grid is the
// concatenation of every segment's (image, 256-pixel block) list, so the small levels no longer pay a launch each.
struct GnSegs {
  int count, C;
  int blk0[LDB_GN_MAX_SEG + 1];          // first block of each segment; [count] = grid size
  int HW[LDB_GN_MAX_SEG], bpi[LDB_GN_MAX_SEG];   // pixels per image, blocks per image
  const void* x[LDB_GN_MAX_SEG];
  const void* y[LDB_GN_MAX_SEG];
  const void* dy[LDB_GN_MAX_SEG];
  void* out[LDB_GN_MAX_SEG];             // y (apply) or dx (backward)
  const float* stats[LDB_GN_MAX_SEG];    // [N, G, 2] sum / sum of squares of x
  float* acc[LDB_GN_MAX_SEG];            // [N, G, 2]: stats being accumulated, or the backward reductions
};
struct GnBlock { int sg, n, p0, p1; };
__device__ __forceinline__ GnBlock gn_block(const GnSegs& S) {
  GnBlock b;
  b.sg = 0;
#pragma unroll
  for (int i = 1; i < LDB_GN_MAX_SEG; ++i) b.sg += (i < S.count && (int)blockIdx.x >= S.blk0[i]) ? 1 : 0;
  const int local = blockIdx.x - S.blk0[b.sg];
  b.n = local / S.bpi[b.sg];
  b.p0 = (local - b.n * S.bpi[b.sg]) * GN_PIX;
  b.p1 = min(b.p0 + GN_PIX, S.HW[b.sg]);
  return b;
}

template <typename T>
__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const __grid_constant__ GnSegs S) {
  const GnBlock B = gn_block(S);
  const int C = S.C, HW = S.HW[B.sg], n = B.n, p0 = B.p0, p1 = B.p1;
  const T* __restrict__ x = (const T*)S.x[B.sg];
  float* __restrict__ stats = S.acc[B.sg];
  const int G = C / GN_CPG;
  const int g = threadIdx.x % G, lanes = GN_THREADS / G, r0 = threadIdx.x / G;
  const T* base = x + (size_t)n * HW * C + g * 8;
  float s = 0.f, ss = 0.f;
  int p = p0 + r0;
  for (; p + 3 * lanes < p1; p += 4 * lanes) {
    float v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) ld8<T>(base + (size_t)(p + u * lanes) * C, v[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { s += v[u][e]; ss += v[u][e] * v[u][e]; }
    }
  }
  for (; p < p1; p += lanes) {
    float v[8];
    ld8<T>(base + (size_t)p * C, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) { s += v[e]; ss += v[e] * v[e]; }
  }
  __shared__ float sh[2][GN_THREADS];
  sh[0][threadIdx.x] = s;
  sh[1][threadIdx.x] = ss;
  __syncthreads();
  if (threadIdx.x < G) {
    float a = 0.f, b = 0.f;
    for (int r = 0; r < lanes; ++r) { a += sh[0][r * G + threadIdx.x]; b += sh[1][r * G + threadIdx.x]; }
    atomicAdd(stats + ((size_t)n * G + threadIdx.x) * 2, a);
    atomicAdd(stats + ((size_t)n * G + threadIdx.x) * 2 + 1, b);
  }
}

// y = relu((x - mean) * rstd * gamma + beta)   (relu optional)
template <typename T>
__global__ void __launch_bounds__(GN_THREADS)
gn_apply_kernel(const __grid_constant__ GnSegs S, const float* __restrict__ gamma, const float* __restrict__ beta,
                float eps, int relu) {
  const GnBlock B = gn_block(S);
  const int C = S.C, HW = S.HW[B.sg], n = B.n, p0 = B.p0, p1 = B.p1;
  const T* __restrict__ x = (const T*)S.x[B.sg];
  const float* __restrict__ stats = S.stats[B.sg];
  T* __restrict__ y = (T*)S.out[B.sg];
  const int G = C / GN_CPG;
  const int g = threadIdx.x % G, lanes = GN_THREADS / G, r0 = threadIdx.x / G;
  const float inv_m = 1.f / ((float)HW * GN_CPG);
  const float mean = __ldg(stats + ((size_t)n * G + g) * 2) * inv_m;
  const float rstd = rsqrtf(fmaxf(__ldg(stats + ((size_t)n * G + g) * 2 + 1) * inv_m - mean * mean, 0.f) + eps);
  float a[8], b[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    a[e] = rstd * __ldg(gamma + g * GN_CPG + e);
    b[e] = __ldg(beta + g * GN_CPG + e) - mean * a[e];
  }
  const T* xin = x + (size_t)n * HW * C + g * 8;
  T* yout = y + (size_t)n * HW * C + g * 8;
  int p = p0 + r0;
  for (; p + 3 * lanes < p1; p += 4 * lanes) {
    float v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) ld8<T>(xin + (size_t)(p + u * lanes) * C, v[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float o = v[u][e] * a[e] + b[e];
        v[u][e] = relu ? fmaxf(o, 0.f) : o;
      }
      st8<T>(yout + (size_t)(p + u * lanes) * C, v[u]);
    }
  }
  for (; p < p1; p += lanes) {
    float v[8];
    ld8<T>(xin + (size_t)p * C, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float o = v[e] * a[e] + b[e];
      v[e] = relu ? fmaxf(o, 0.f) : o;
    }
    st8<T>(yout + (size_t)p * C, v);
  }
}

// GroupNorm+ReLU backward, pass 1: per-(n,group) sums of dxh and dxh*xh, per-channel dgamma/dbeta.
//   g = dy * (y > 0), xh = (x - mean)*rstd, dxh = g*gamma
// red: [N, G, 2] fp32 accumulated; dgamma/dbeta: [C] fp32 accumulated.
// y == null: the ReLU mask is recomputed from x (y > 0 <=> xh*gamma + beta > 0: bf16 / fp32 rounding of a positive value
// stays positive), which saves reading y in both backward passes (14 -> 10 B / element).
// Register diet (ncu r02: 120 registers, 2 blocks / SM, 28 % of HBM peak): three pixels in flight as raw vectors,
// unpacked one at a time; <= 85 registers -> 3 blocks / SM.
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> {
  uint4 q;
  __device__ __forceinline__ void ld(const __nv_bfloat16* p) { q = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void get(float v[8]) const { unpack8(q, v); }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void ld(const float* p) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
  __device__ __forceinline__ void get(float v[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

template <typename T, bool REMASK>
__global__ void __launch_bounds__(GN_THREADS, 3)
gn_bwd_stats_kernel(const __grid_constant__ GnSegs S, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, float eps) {
  const GnBlock B = gn_block(S);
  const int C = S.C, HW = S.HW[B.sg], n = B.n, p0 = B.p0, p1 = B.p1;
  const T* __restrict__ dy = (const T*)S.dy[B.sg];
  const T* __restrict__ y = (const T*)S.y[B.sg];
  const T* __restrict__ x = (const T*)S.x[B.sg];
  const float* __restrict__ stats = S.stats[B.sg];
  float* __restrict__ red = S.acc[B.sg];
  const int G = C / GN_CPG;
  const int g = threadIdx.x % G, lanes = GN_THREADS / G, r0 = threadIdx.x / G;
  const float inv_m = 1.f / ((float)HW * GN_CPG);
  const float mean = __ldg(stats + ((size_t)n * G + g) * 2) * inv_m;
  const float rstd = rsqrtf(fmaxf(__ldg(stats + ((size_t)n * G + g) * 2 + 1) * inv_m - mean * mean, 0.f) + eps);
  constexpr bool remask = REMASK;
  float gam[8], bet[8], acc[18];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    gam[e] = __ldg(gamma + g * GN_CPG + e);
    bet[e] = remask ? __ldg(beta + g * GN_CPG + e) : 0.f;
  }
#pragma unroll
  for (int e = 0; e < 18; ++e) acc[e] = 0.f;   // [0]=sum dxh, [1]=sum dxh*xh, [2..9]=dgamma, [10..17]=dbeta
  const size_t img = (size_t)n * HW * C + g * 8;
  const T* pdy = dy + img;
  const T* py = remask ? nullptr : y + img;
  const T* px = x + img;
  constexpr int U = 3;
  for (int p = p0 + r0; p < p1; p += U * lanes) {
    Raw8<T> rd[U], rx[U], ry[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (p + u * lanes < p1) {
        rd[u].ld(pdy + (size_t)(p + u * lanes) * C);
        rx[u].ld(px + (size_t)(p + u * lanes) * C);
        if (!remask) ry[u].ld(py + (size_t)(p + u * lanes) * C);
      }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (p + u * lanes >= p1) break;
      float vd[8], vx[8], vy[8];
      rd[u].get(vd);
      rx[u].get(vx);
      if (!remask) ry[u].get(vy);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = (vx[e] - mean) * rstd;
        const bool on = remask ? (xh * gam[e] + bet[e] > 0.f) : (vy[e] > 0.f);
        const float gg = on ? vd[e] : 0.f;
        const float dxh = gg * gam[e];
        acc[0] += dxh;
        acc[1] += dxh * xh;
        acc[2 + e] += gg * xh;
        acc[10 + e] += gg;
      }
    }
  }
  __shared__ float sh[18][GN_THREADS];
#pragma unroll
  for (int e = 0; e < 18; ++e) sh[e][threadIdx.x] = acc[e];
  __syncthreads();
  for (int o = threadIdx.x; o < 18 * G; o += GN_THREADS) {
    const int e = o / G, gg = o % G;
    float t = 0.f;
    for (int r = 0; r < lanes; ++r) t += sh[e][r * G + gg];
    if (e < 2) atomicAdd(red + ((size_t)n * G + gg) * 2 + e, t);
    else if (e < 10) atomicAdd(dgamma + gg * GN_CPG + (e - 2), t);
    else atomicAdd(dbeta + gg * GN_CPG + (e - 10), t);
  }
}

// pass 2: dx = rstd * (dxh - mean(dxh) - xh * mean(dxh*xh))
template <typename T, bool REMASK>
__global__ void __launch_bounds__(GN_THREADS, 3)
gn_bwd_apply_kernel(const __grid_constant__ GnSegs S, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float eps) {
  const GnBlock B = gn_block(S);
  const int C = S.C, HW = S.HW[B.sg], n = B.n, p0 = B.p0, p1 = B.p1;
  const T* __restrict__ dy = (const T*)S.dy[B.sg];
  const T* __restrict__ y = (const T*)S.y[B.sg];
  const T* __restrict__ x = (const T*)S.x[B.sg];
  const float* __restrict__ stats = S.stats[B.sg];
  const float* __restrict__ red = S.acc[B.sg];
  T* __restrict__ dx = (T*)S.out[B.sg];
  const int G = C / GN_CPG;
  const int g = threadIdx.x % G, lanes = GN_THREADS / G, r0 = threadIdx.x / G;
  const float inv_m = 1.f / ((float)HW * GN_CPG);
  const float mean = __ldg(stats + ((size_t)n * G + g) * 2) * inv_m;
  const float rstd = rsqrtf(fmaxf(__ldg(stats + ((size_t)n * G + g) * 2 + 1) * inv_m - mean * mean, 0.f) + eps);
  const float m1 = __ldg(red + ((size_t)n * G + g) * 2) * inv_m, m2 = __ldg(red + ((size_t)n * G + g) * 2 + 1) * inv_m;
  constexpr bool remask = REMASK;
  float gam[8], bet[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    gam[e] = __ldg(gamma + g * GN_CPG + e);
    bet[e] = remask ? __ldg(beta + g * GN_CPG + e) : 0.f;
  }
  const size_t img = (size_t)n * HW * C + g * 8;
  const T* pdy = dy + img;
  const T* py = remask ? nullptr : y + img;
  const T* px = x + img;
  T* pdx = dx + img;
  constexpr int U = 3;
  for (int p = p0 + r0; p < p1; p += U * lanes) {
    Raw8<T> rd[U], rx[U], ry[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (p + u * lanes < p1) {
        rd[u].ld(pdy + (size_t)(p + u * lanes) * C);
        rx[u].ld(px + (size_t)(p + u * lanes) * C);
        if (!remask) ry[u].ld(py + (size_t)(p + u * lanes) * C);
      }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (p + u * lanes >= p1) break;
      float vd[8], vx[8], vy[8];
      rd[u].get(vd);
      rx[u].get(vx);
      if (!remask) ry[u].get(vy);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = (vx[e] - mean) * rstd;
        const bool on = remask ? (xh * gam[e] + bet[e] > 0.f) : (vy[e] > 0.f);
        const float gg = on ? vd[e] : 0.f;
        vd[e] = rstd * (gg * gam[e] - m1 - xh * m2);
      }
      st8<T>(pdx + (size_t)(p + u * lanes) * C, vd);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Backward of `fine += nearest_upsample(coarse)` (fpn.py:182-191): coarse_grad[n,yc,xc,:] +=
// sum of fine_grad over the fine pixels that read it (src index = floor(dst * Hc / Hf)).
template <typename T>
__global__ void upsample_add_bwd_kernel(const T* __restrict__ dfine, T* __restrict__ dcoarse,
                                        int N, int Hf, int Wf, int Hc, int Wc, int C) {
  const int cg = C / 8;
  const long long total = (long long)N * Hc * Wc * cg;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % cg);
    const long long pix = t / cg;
    const int xc = (int)(pix % Wc), yc = (int)((pix / Wc) % Hc), n = (int)(pix / ((long long)Wc * Hc));
    // fine rows y with floor(y*Hc/Hf) == yc  <=>  y in [ceil(yc*Hf/Hc), ceil((yc+1)*Hf/Hc))
    const int y0 = (yc * Hf + Hc - 1) / Hc, y1 = min(((yc + 1) * Hf + Hc - 1) / Hc, Hf);
    const int x0 = (xc * Wf + Wc - 1) / Wc, x1 = min(((xc + 1) * Wf + Wc - 1) / Wc, Wf);
    float acc[8];
    ld8<T>(dcoarse + (size_t)t * 8, acc);
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) {
        float v[8];
        ld8<T>(dfine + (((size_t)n * Hf + y) * Wf + x) * C + g * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += v[e];
      }
    st8<T>(dcoarse + (size_t)t * 8, acc);
  }
}

// out = dy * (y > 0), 8 elements per thread
template <typename T>
__global__ void relu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ out, long long n8) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n8;
       t += (long long)gridDim.x * blockDim.x) {
    float a[8], b[8];
    ld8<T>(dy + (size_t)t * 8, a);
    ld8<T>(y + (size_t)t * 8, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = b[e] > 0.f ? a[e] : 0.f;
    st8<T>(out + (size_t)t * 8, a);
  }
}

// x [rows, ld_in] fp32 -> `planes` stacked bf16 tensors [planes][rows, ld_out] with x ~= sum of the planes:
// p0 = bf16(x), p1 = bf16(x - p0), p2 = bf16(x - p0 - p1) (each residual is exact in fp32).  Columns [C, ld_out)
// are zero-filled.  Operand preparation of the split-bf16 ("parity") convolutions.
__global__ void split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long rows, int C,
                                  int ld_in, int ld_out, int planes) {
  const int cg = ld_out / 8;
  const long long total = rows * cg;
  const size_t plane = (size_t)rows * ld_out;
  const bool vec = (ld_in % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % cg);
    const long long r = t / cg;
    float v[8];
    const float* src = x + (size_t)r * ld_in + g * 8;
    if (vec && g * 8 + 8 <= C) {
      ld8<float>(src, v);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = (g * 8 + e < C) ? __ldg(src + e) : 0.f;
    }
    for (int pl = 0; pl < planes; ++pl) {
      float h[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        h[e] = __bfloat162float(__float2bfloat16_rn(v[e]));
        v[e] -= h[e];
      }
      *reinterpret_cast<uint4*>(out + pl * plane + (size_t)r * ld_out + g * 8) = pack8(h);
    }
  }
}

// out[c] += sum_rows x[row, c]   (x: [M, ld] bf16 or f32; C % 8 == 0 for bf16, C % 4 == 0 for f32)
// block = (channel groups) x (row lanes); each block owns a band of rows, 4 independent 16-byte loads in flight
// per thread, ONE shared-memory pass to fold the row lanes, one atomic per (block, channel).
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, float* __restrict__ out, long long M,
                                                      int C, int ld, int rows_per_block) {
  constexpr int VW = (sizeof(T) == 2) ? 8 : 4;
  const int cg = C / VW;                       // vector columns
  const int cgb = cg < 256 ? cg : 256;          // vector columns handled per pass by this block
  const int lanes = 256 / cgb;
  const int tx = threadIdx.x % cgb, ty = threadIdx.x / cgb;
  const long long m0 = (long long)blockIdx.x * rows_per_block;
  const long long m1 = (m0 + rows_per_block < M) ? m0 + rows_per_block : M;
  __shared__ float sh[VW][256];
  for (int g = tx; g < cg; g += cgb) {          // cg > 256 only for C > 2048 bf16
    float acc[VW];
#pragma unroll
    for (int e = 0; e < VW; ++e) acc[e] = 0.f;
    if (ty < lanes) {
      long long m = m0 + ty;
      for (; m + 3LL * lanes < m1; m += 4LL * lanes) {
        if (sizeof(T) == 2) {
          uint4 q[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            q[u] = __ldg(reinterpret_cast<const uint4*>((const __nv_bfloat16*)x + (size_t)(m + (long long)u * lanes) * ld) + g);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float v[8];
            unpack8(q[u], v);
#pragma unroll
            for (int e = 0; e < VW; ++e) acc[e] += v[e];
          }
        } else {
          float4 q[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            q[u] = __ldg(reinterpret_cast<const float4*>((const float*)x + (size_t)(m + (long long)u * lanes) * ld) + g);
#pragma unroll
          for (int u = 0; u < 4; ++u) { acc[0] += q[u].x; acc[1] += q[u].y; acc[2] += q[u].z; acc[3] += q[u].w; }
        }
      }
      for (; m < m1; m += lanes) {
        if (sizeof(T) == 2) {
          float v[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>((const __nv_bfloat16*)x + (size_t)m * ld) + g), v);
#pragma unroll
          for (int e = 0; e < VW; ++e) acc[e] += v[e];
        } else {
          const float4 v = __ldg(reinterpret_cast<const float4*>((const float*)x + (size_t)m * ld) + g);
          acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < VW; ++e) sh[e][threadIdx.x] = acc[e];
    __syncthreads();
    // one thread per vector column folds the row lanes and issues 16-byte vector atomics (same-address atomic
    // contention between blocks, not bandwidth, bounds this kernel)
    if (threadIdx.x < cgb && threadIdx.x + (g - tx) < cg) {
      float t[VW];
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        t[e] = 0.f;
        for (int r = 0; r < lanes; ++r) t[e] += sh[e][r * cgb + threadIdx.x];
      }
      float* o = out + (size_t)(g - tx + threadIdx.x) * VW;
      if ((reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
        for (int e = 0; e < VW; e += 4) atomicAdd(reinterpret_cast<float4*>(o + e), make_float4(t[e], t[e + 1], t[e + 2], t[e + 3]));
      } else {
#pragma unroll
        for (int e = 0; e < VW; ++e) atomicAdd(o + e, t[e]);
      }
    }
  }
}

// SGD with momentum and weight decay on a flat fp32 buffer (torch.optim.SGD semantics, dampening 0):
//   g += wd*w ; buf = first ? g : mom*buf + g ; w -= lr*buf
// grad_scale folds the data-parallel mean (1/world) into the update.
__global__ void sgd_flat_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ buf,
                                long long n, float lr, float mom, float wd, float grad_scale, int first) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale + wd * w[i];
    float b = first ? gi : mom * buf[i] + gi;
    buf[i] = b;
    w[i] -= lr * b;
  }
}

// graph-capturable variant: hyper-parameters live in device memory [lr, momentum, wd, grad_scale, first(0/1)]
__global__ void sgd_flat_dev_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ buf,
                                    long long n, const float* __restrict__ hyper) {
  const float lr = hyper[0], mom = hyper[1], wd = hyper[2], gs = hyper[3];
  const bool first = hyper[4] != 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * gs + wd * w[i];
    float b = first ? gi : mom * buf[i] + gi;
    buf[i] = b;
    w[i] -= lr * b;
  }
}

// Eval-mode BatchNorm folded to a per-channel affine (resnet.py:639-649 norm_eval=True):
//   scale = gamma * rsqrt(var + eps), shift = beta - mean * scale      and its adjoint
//   dgamma = (d_scale - d_shift * mean) * rsqrt(var + eps), dbeta = d_shift
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, float eps,
                               float* __restrict__ scale, float* __restrict__ shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const float s = gamma[c] * rsqrtf(var[c] + eps);
    scale[c] = s;
    shift[c] = beta[c] - mean[c] * s;
  }
}
__global__ void bn_fold_bwd_kernel(const float* __restrict__ d_scale, const float* __restrict__ d_shift,
                                   const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                   float* __restrict__ dgamma, float* __restrict__ dbeta, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const float ds = d_scale ? d_scale[c] : 0.f, db = d_shift ? d_shift[c] : 0.f;
    dgamma[c] += (ds - db * mean[c]) * rsqrtf(var[c] + eps);   // accumulates into the parameter gradients
    dbeta[c] += db;
  }
}

// IMLoss (kd_loss.py:91-120): loss_weight * mean((x - t)^2) over all elements, and d/dx.
template <typename T>
__global__ void __launch_bounds__(256) mse_fwd_bwd_kernel(const T* __restrict__ x, const T* __restrict__ t, long long n,
                                                           float scale, float* __restrict__ loss, T* __restrict__ dx) {
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float a, b;
    if (sizeof(T) == 4) { a = ((const float*)x)[i]; b = ((const float*)t)[i]; }
    else { a = __bfloat162float(((const __nv_bfloat16*)x)[i]); b = __bfloat162float(((const __nv_bfloat16*)t)[i]); }
    const float d = a - b;
    acc += d * d;
    if (dx) {
      if (sizeof(T) == 4) ((float*)dx)[i] = 2.f * scale * d;
      else ((__nv_bfloat16*)dx)[i] = __float2bfloat16_rn(2.f * scale * d);
    }
  }
  acc = warp_sum(acc);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += sh[i];
    atomicAdd(loss, s * scale);
  }
}

static int grid_for(long long total, int threads = 256, int cap = 148 * 16) {
  long long b = (total + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace ldb

using namespace ldb;

extern "C" int ldb200_stem_im2col(const float* img, void* out, int N, int H, int W, int out_dtype, void* stream_) {
  LDB_REQUIRE(img && out && N > 0 && H > 0 && W > 0, LDB_EINVAL, "stem_im2col: bad argument");
  const int HO = (H + 6 - 7) / 2 + 1, WO = (W + 6 - 7) / 2 + 1;
  dim3 grid((WO + IM_PX - 1) / IM_PX, HO, N);
  if (out_dtype == LDB_F32)
    stem_im2col_kernel<float><<<grid, 256, 0, (cudaStream_t)stream_>>>(img, (float*)out, N, H, W, HO, WO);
  else
    stem_im2col_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream_>>>(img, (__nv_bfloat16*)out, N, H, W, HO, WO);
  LDB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ldb200_maxpool3x3s2(const void* in, void* out, int N, int H, int W, int C, int dtype, void* stream_) {
  LDB_REQUIRE(in && out && C % 8 == 0, LDB_EINVAL, "maxpool: bad argument (C %% 8)");
  const int HO = (H + 2 - 3) / 2 + 1, WO = (W + 2 - 3) / 2 + 1;
  const long long total = (long long)N * HO * WO * (C / 8);
  if (dtype == LDB_F32)
    maxpool3x3s2_kernel<float><<<grid_for(total), 256, 0, (cudaStream_t)stream_>>>((const float*)in, (float*)out, N, H, W, C, HO, WO);
  else
    maxpool3x3s2_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, (cudaStream_t)stream_>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out,
                                                                                          N, H, W, C, HO, WO);
  LDB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ldb200_split_bf16(const float* x, int64_t rows, int C, int ld_in, int planes, void* out, int ld_out,
                                 void* stream_) {
  LDB_REQUIRE(x && out && rows >= 0 && C > 0 && ld_in >= C && ld_out >= C && ld_out % 8 == 0 && planes >= 1 && planes <= 3,
              LDB_EINVAL, "split_bf16: bad argument (ld_out %% 8, 1 <= planes <= 3)");
  if (rows == 0) return 0;
  split_bf16_kernel<<<grid_for(rows * (ld_out / 8)), 256, 0, (cudaStream_t)stream_>>>(x, (__nv_bfloat16*)out, rows, C, ld_in,
                                                                                      ld_out, planes);
  LDB_CUDA(cudaGetLastError());
  return 0;
}

static int gn_check(int C, int G) {
  LDB_REQUIRE(G > 0 && C == G * GN_CPG && G <= 256 && 256 % G == 0, LDB_ENOSUP,
              "group norm kernels support 8 channels per group and a group count dividing 256 (C=%d, G=%d)", C, G);
  return 0;
}

// segment table of a launch: `segs[i]` describes one tensor of [N, HW, C]
static int gn_make_segs(const ldb_gn_seg* segs, int count, int C, int G, GnSegs* S) {
  int e = gn_check(C, G);
  if (e) return e;
  LDB_REQUIRE(segs && count >= 1 && count <= LDB_GN_MAX_SEG, LDB_EINVAL, "group norm: 1..%d segments", LDB_GN_MAX_SEG);
  memset(S, 0, sizeof(*S));
  S->count = count;
  S->C = C;
  int blk = 0;
  for (int i = 0; i < count; ++i) {
    LDB_REQUIRE(segs[i].N >= 1 && segs[i].HW >= 1, LDB_EINVAL, "group norm: segment %d shape", i);
    S->blk0[i] = blk;
    S->HW[i] = segs[i].HW;
    S->bpi[i] = (segs[i].HW + GN_PIX - 1) / GN_PIX;
    blk += S->bpi[i] * segs[i].N;
    S->x[i] = segs[i].x; S->y[i] = segs[i].y; S->dy[i] = segs[i].dy;
    S->stats[i] = segs[i].stats;
  }
  for (int i = count; i <= LDB_GN_MAX_SEG; ++i) S->blk0[i] = blk;
  return 0;
}

extern "C" int ldb200_gn_stats(const void* x, float* stats, int N, int HW, int C, int G, int dtype, void* stream_) {
  LDB_REQUIRE(x && stats, LDB_EINVAL, "gn_stats: null pointer");
  ldb_gn_seg seg;
  memset(&seg, 0, sizeof(seg));
  seg.x = x; seg.N = N; seg.HW = HW;
  GnSegs S;
  int e = gn_make_segs(&seg, 1, C, G, &S);
  if (e) return e;
  S.acc[0] = stats;
  cudaStream_t s = (cudaStream_t)stream_;
  LDB_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * N * G, s));
  if (dtype == LDB_F32) gn_stats_kernel<float><<<S.blk0[1], GN_THREADS, 0, s>>>(S);
  else gn_stats_kernel<__nv_bfloat16><<<S.blk0[1], GN_THREADS, 0, s>>>(S);
  LDB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ldb200_gn_apply_multi(const ldb_gn_seg* segs, int count, const float* gamma, const float* beta, int C,
                                     int G, float eps, int relu, int dtype, void* stream_) {
  GnSegs S;
  int e = gn_make_segs(segs, count, C, G, &S);
  if (e) return e;
  LDB_REQUIRE(gamma && beta, LDB_EINVAL, "gn_apply: null pointer");
  for (int i = 0; i < count; ++i) {
    LDB_REQUIRE(segs[i].x && segs[i].stats && segs[i].out, LDB_EINVAL, "gn_apply: null pointer in segment %d", i);
    S.out[i] = segs[i].out;
  }
  if (dtype == LDB_F32)
    gn_apply_kernel<float><<<S.blk0[count], GN_THREADS, 0, (cudaStream_t)stream_>>>(S, gamma, beta, eps, relu);
  else
    gn_apply_kernel<__nv_bfloat16><<<S.blk0[count], GN_THREADS, 0, (cudaStream_t)stream_>>>(S, gamma, beta, eps, relu);
  LDB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ldb200_gn_apply(const void* x, const float* stats, const float* gamma, const float* beta, void* y,
                               int N, int HW, int C, int G, float eps, int relu, int dtype, void* stream_) {
  ldb_gn_seg seg;
  memset(&seg, 0, sizeof(seg));
  seg.x = x; seg.stats = stats; seg.out = y; seg.N = N; seg.HW = HW;
  return ldb200_gn_apply_multi(&seg, 1, gamma, beta, C, G, eps, relu, dtype, stream_);
}

extern "C" int ldb200_gn_relu_bwd_multi(const ldb_gn_seg* segs, int count, const float* gamma, const float* beta,
                                        float* dgamma, float* dbeta, int C, int G, float eps, int dtype, void* stream_) {
  GnSegs S;
  int e = gn_make_segs(segs, count, C, G, &S);
  if (e) return e;
  LDB_REQUIRE(gamma && dgamma && dbeta, LDB_EINVAL, "gn_relu_bwd: null pointer");
  cudaStream_t s = (cudaStream_t)stream_;
  for (int i = 0; i < count; ++i) {
    LDB_REQUIRE(segs[i].dy && segs[i].x && segs[i].stats && segs[i].red_ws && segs[i].out, LDB_EINVAL,
                "gn_relu_bwd: null pointer in segment %d", i);
    LDB_REQUIRE(segs[i].y || beta, LDB_EINVAL, "gn_relu_bwd: segment %d has no y: beta is needed to recompute the mask", i);
    S.out[i] = segs[i].out;
    S.acc[i] = segs[i].red_ws;
    // the per-segment scratch may be slices of one buffer the caller zeroed itself (red_zeroed)
    if (!segs[i].red_zeroed) LDB_CUDA(cudaMemsetAsync(segs[i].red_ws, 0, sizeof(float) * 2 * segs[i].N * G, s));
  }
  // all segments read y, or none does (the mask is then recomputed from x)
  bool remask = segs[0].y == nullptr;
  for (int i = 1; i < count; ++i)
    LDB_REQUIRE((segs[i].y == nullptr) == remask, LDB_EINVAL, "gn_relu_bwd: y must be given for all segments or for none");
  const int grid = S.blk0[count];
#define LDB_GN_BWD(T, R)                                                                               \
  do {                                                                                                 \
    gn_bwd_stats_kernel<T, R><<<grid, GN_THREADS, 0, s>>>(S, gamma, beta, dgamma, dbeta, eps);          \
    LDB_CUDA(cudaGetLastError());                                                                      \
    gn_bwd_apply_kernel<T, R><<<grid, GN_THREADS, 0, s>>>(S, gamma, beta, eps);                         \
  } while (0)
  if (dtype == LDB_F32) {
    if (remask) LDB_GN_BWD(float, true); else LDB_GN_BWD(float, false);
  } else {
    if (remask) LDB_GN_BWD(__nv_bfloat16, true); else LDB_GN_BWD(__nv_bfloat16, false);
  }
#undef LDB_GN_BWD
  LDB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ldb200_gn_relu_bwd(const void* dy, const void* y, const void* x, const float* stats, const float* gamma,
                                  float* red_ws, float* dgamma, float* dbeta, void* dx, int N, int HW, int C, int G,
                                  float eps, int dtype, void* stream_) {
  ldb_gn_seg seg;
  memset(&seg, 0, sizeof(seg));
  seg.dy = dy; seg.y = y; seg.x = x; seg.stats = stats; seg.red_ws = red_ws; seg.out = dx; seg.N = N; seg.HW = HW;
  return ldb200_gn_relu_bwd_multi(&seg, 1, gamma, nullptr, dgamma, dbeta, C, G, eps, dtype, stream_);
}

extern "C" int ldb200_upsample_add_bwd(const void* dfine, void* dcoarse, int N, int Hf, int Wf, int Hc, int Wc, int C,
                                       int dtype, void* stream_) {
  LDB_REQUIRE(dfine && dcoarse && C % 8 == 0 && Hc <= Hf && Wc <= Wf, LDB_EINVAL, "upsample_add_bwd: bad argument");
  const long long total = (long long)N * Hc * Wc * (C / 8);
  if (dtype == LDB_F32)
    upsample_add_bwd_kernel<float><<<grid_for(total), 256, 0, (cudaStream_t)stream_>>>((const float*)dfine, (float*)dcoarse, N, Hf, Wf, Hc, Wc, C);
  else
    upsample_add_bwd_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, (cudaStream_t)stream_>>>((const __nv_bfloat16*)dfine,
                                                                                               (__nv_bfloat16*)dcoarse, N, Hf, Wf, Hc, Wc, C);
  LDB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ldb200_relu_bwd(const void* dy, const void* y, void* out, int64_t numel, int dtype, void* stream_) {
  LDB_REQUIRE(dy && y && out && numel % 8 == 0, LDB_EINVAL, "relu_bwd: bad argument (numel %% 8)");
  if (dtype == LDB_F32)
    relu_bwd_kernel<float><<<grid_for(numel / 8), 256, 0, (cudaStream_t)stream_>>>((const float*)dy, (const float*)y, (float*)out, numel / 8);
  else
    relu_bwd_kernel<__nv_bfloat16><<<grid_for(numel / 8), 256, 0, (cudaStream_t)stream_>>>((const __nv_bfloat16*)dy,
                                                                                           (const __nv_bfloat16*)y, (__nv_bfloat16*)out, numel / 8);
  LDB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ldb200_colsum(const void* x, float* out, int64_t rows, int C, int ld, int dtype, void* stream_) {
  LDB_REQUIRE(x && out && rows >= 0 && C > 0 && ld >= C, LDB_EINVAL, "colsum: bad argument");
  if (rows == 0) return 0;
  // two blocks per SM (each extra block is one more contended atomic per channel), at least 32 rows per block
  long long rpb = (rows + 148 * 2 - 1) / (148 * 2);
  if (rpb < 32) rpb = 32;
  const int grid = (int)((rows + rpb - 1) / rpb);
  if (dtype == LDB_BF16) {
    LDB_REQUIRE(C % 8 == 0 && ld % 8 == 0, LDB_EINVAL, "colsum: bf16 needs C, ld %% 8 == 0");
    colsum_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream_>>>((const __nv_bfloat16*)x, out, rows, C, ld, (int)rpb);
  } else {
    LDB_REQUIRE(C % 4 == 0 && ld % 4 == 0, LDB_EINVAL, "colsum: f32 needs C, ld %% 4 == 0");
    colsum_kernel<float><<<grid, 256, 0, (cudaStream_t)stream_>>>((const float*)x, out, rows, C, ld, (int)rpb);
  }
  LDB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ldb200_sgd_flat(float* w, const float* g, float* buf, int64_t n, float lr, float momentum,
                               float weight_decay, float grad_scale, int first_step, void* stream_) {
  LDB_REQUIRE(w && g && buf && n >= 0, LDB_EINVAL, "sgd_flat: bad argument");
  if (n == 0) return 0;
  sgd_flat_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream_>>>(w, g, buf, n, lr, momentum, weight_decay, grad_scale,
                                                                  first_step);
  LDB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ldb200_sgd_flat_dev(float* w, const float* g, float* buf, int64_t n, const float* hyper_dev,
                                   void* stream_) {
  LDB_REQUIRE(w && g && buf && hyper_dev && n >= 0, LDB_EINVAL, "sgd_flat_dev: bad argument");
  if (n == 0) return 0;
  sgd_flat_dev_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream_>>>(w, g, buf, n, hyper_dev);
  LDB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ldb200_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                              float* scale, float* shift, int C, void* stream_) {
  LDB_REQUIRE(gamma && beta && mean && var && scale && shift && C > 0, LDB_EINVAL, "bn_fold: bad argument");
  bn_fold_kernel<<<(C + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(gamma, beta, mean, var, eps, scale, shift, C);
  LDB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ldb200_bn_fold_bwd(const float* d_scale, const float* d_shift, const float* mean, const float* var,
                                  float eps, float* dgamma, float* dbeta, int C, void* stream_) {
  LDB_REQUIRE(mean && var && dgamma && dbeta && C > 0, LDB_EINVAL, "bn_fold_bwd: bad argument");
  bn_fold_bwd_kernel<<<(C + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(d_scale, d_shift, mean, var, eps, dgamma, dbeta, C);
  LDB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ldb200_mse_fwd_bwd(const void* x, const void* target, int64_t numel, float loss_weight, int dtype,
                                  float* loss, void* dx, void* stream_) {
  LDB_REQUIRE(numel >= 0 && loss, LDB_EINVAL, "mse_fwd_bwd: bad argument");
  if (numel == 0) return 0;
  LDB_REQUIRE(x && target, LDB_EINVAL, "mse_fwd_bwd: null pointer");
  const float scale = loss_weight / (float)numel;
  if (dtype == LDB_F32)
    mse_fwd_bwd_kernel<float><<<grid_for(numel), 256, 0, (cudaStream_t)stream_>>>((const float*)x, (const float*)target, numel, scale, loss, (float*)dx);
  else if (dtype == LDB_BF16)
    mse_fwd_bwd_kernel<__nv_bfloat16><<<grid_for(numel), 256, 0, (cudaStream_t)stream_>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)target, numel, scale, loss, (__nv_bfloat16*)dx);
  else
    LDB_REQUIRE(false, LDB_EINVAL, "mse_fwd_bwd: dtype");
  LDB_CUDA(cudaGetLastError());
  return 0;
}
