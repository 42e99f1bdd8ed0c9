// Fused ResNet stem for student AND teacher in one launch (the default route of the detector; LDB200_NO_FUSED_STEM=1
// selects the im2col + GEMM route):
// conv1 7x7/2 pad 3 on the 3-channel fp32 NCHW image + folded eval-BN + ReLU (mmdet/models/backbones/resnet.py:
// 626-629) for up to two networks that share the input (kd_one_stage.py:60-66: the student and the frozen teacher see
// the same batch), written as NHWC bf16.
//
// Why: the im2col route writes a [N*HO*WO, 192] bf16 operand (1.65 GB at 16 x 800 x 1344) that two GEMMs read back --
// 2.9 % of the step for the im2col kernel alone plus ~5 GB of HBM traffic.  Here the A operand never leaves the SM:
//   * producer warps (256 threads, two per operand row) stage the 3 x 21 x 37 input patch of an 8 x 16 output tile in shared memory and
//     expand it into the K-major SWIZZLE_128B UMMA layout [128 pixels][192] (K = (r*7+s)*3+c, 147 zero-padded to 192:
//     the same order `stem_im2col_kernel` / `ResNet.stem` use), one thread per pixel row, 16-byte swizzled stores;
//   * the [64*nets][192] weight operand is TMA-loaded once per CTA and stays resident;
//   * one thread issues 12 tcgen05.mma (M=128, N=64*nets, K=16) per tile into one of two TMEM accumulator stages;
//   * 8 epilogue warps apply each network's scale/shift + ReLU and store NHWC pixel rows (64 bytes per warp half).
// Persistent CTAs (one per SM) walk the tile list; generic-proxy writes of the operand are published to the tensor
// core's async proxy with fence.proxy.async before the full-barrier arrive.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ldb {
namespace {

constexpr int ST_TH = 8, ST_TW = 16;                 // output tile (ST_TH * ST_TW = 128 = UMMA M)
constexpr int ST_PH = 2 * ST_TH + 5, ST_PW = 2 * ST_TW + 5;   // input patch 21 x 37
constexpr int ST_PATCH = 3 * ST_PH * ST_PW;          // 2331 floats
constexpr int ST_K = 192, ST_KB = 3;                 // 3 k-blocks of 64
constexpr int ST_A_KB_BYTES = 128 * 128;             // one k-block of A: 128 rows x 128 B
constexpr int ST_A_BYTES = ST_KB * ST_A_KB_BYTES;    // 48 KB per stage
constexpr int ST_STAGES = 2;
constexpr int ST_B_KB_BYTES = 128 * 128;             // one k-block of B: up to 128 rows x 128 B
constexpr int ST_B_BYTES = ST_KB * ST_B_KB_BYTES;    // 48 KB
// ncu r02 (128 producers + 4 epilogue warps): 1.28 ms, 0.21 issued warps / scheduler with 2.25 active warps -- latency
// of two long dependent instruction chains; DRAM traffic already the algorithmic 0.21 + 1.05 GB.  Twice the warps on
// both sides: two threads build one operand row (even / odd 16-byte chunks), two warp groups drain the accumulator.
constexpr int ST_PRODUCERS = 256;
constexpr int ST_MMA_WARP = ST_PRODUCERS / 32;       // warp 8
constexpr int ST_EPI_WARPS = 8;
constexpr int ST_THREADS = ST_PRODUCERS + 32 + 32 * ST_EPI_WARPS;   // producers, MMA warp, epilogue warps
constexpr size_t ST_SMEM = (size_t)ST_STAGES * ST_A_BYTES + ST_B_BYTES + 2 * ((ST_PATCH + 3) / 4 * 4) * sizeof(float) +
                           1024 + 256;

struct StemParams {
  const float* img;          // [N,3,H,W] fp32
  int N, H, W, HO, WO;
  int tiles_y, tiles_x, tiles;
  int nets;                  // 1 or 2
  const float* scale[2];     // folded BN per network [64]
  const float* shift[2];
  __nv_bfloat16* out[2];     // [N,HO,WO,64] bf16
};

__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__global__ void __launch_bounds__(ST_THREADS, 1)
stem_tc_kernel(const __grid_constant__ CUtensorMap tmB, const StemParams p) {
  extern __shared__ unsigned char stem_smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)stem_smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;                                      // ST_STAGES x 48 KB
  unsigned char* sB = smem + (size_t)ST_STAGES * ST_A_BYTES;     // 48 KB
  float* patch = (float*)(sB + ST_B_BYTES);                      // 2 x ST_PATCH (double buffer)
  constexpr int PATCH_LD = (ST_PATCH + 3) / 4 * 4;
  uint64_t* bars = (uint64_t*)((unsigned char*)patch + 2 * PATCH_LD * sizeof(float));
  uint64_t* full = bars;                 // [ST_STAGES] producers -> MMA
  uint64_t* empty = bars + ST_STAGES;    // [ST_STAGES] MMA -> producers
  uint64_t* tfull = empty + ST_STAGES;   // [2] MMA -> epilogue
  uint64_t* tempty = tfull + 2;          // [2] epilogue -> MMA
  uint64_t* bbar = tempty + 2;           // weights landed
  uint32_t* tmem_slot = (uint32_t*)(bbar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NCOLS = 64 * p.nets;
  if (threadIdx.x == 0) {
    for (int s = 0; s < ST_STAGES; ++s) {
      mbarrier_init(&full[s], ST_PRODUCERS);
      mbarrier_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbarrier_init(&tfull[a], 1);
      mbarrier_init(&tempty[a], ST_EPI_WARPS);
    }
    mbarrier_init(bbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == ST_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < ST_MMA_WARP) {
    // ===================================== producers: im2col into the UMMA layout ==========================
    const int tid = threadIdx.x;               // load_patch: all 256 threads
    const int t = tid & 127;                   // = tile row: pixel (t / 16, t % 16)
    const int part = tid >> 7;                 // this thread writes the 16-byte chunks grp with (grp & 1) == part
    const int ty = t >> 4, tx = t & 15;
    // the zero padding k = 147..191 never changes: chunks 19..23 of the last k-block, written once per stage
    if (part == 0)
      for (int s = 0; s < ST_STAGES; ++s) {
        unsigned char* row = sA + (size_t)s * ST_A_BYTES + 2 * ST_A_KB_BYTES + t * 128;
        for (int c = 3; c < 8; ++c) *reinterpret_cast<uint4*>(row + ((c ^ (t & 7)) << 4)) = make_uint4(0, 0, 0, 0);
      }
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    auto load_patch = [&](int tile, float* dst) {
      const int txi = tile % p.tiles_x;
      const int r = tile / p.tiles_x;
      const int tyi = r % p.tiles_y, n = r / p.tiles_y;
      const int yi0 = tyi * ST_TH * 2 - 3, xi0 = txi * ST_TW * 2 - 3;
      for (int i = tid; i < ST_PATCH; i += ST_PRODUCERS) {
        const int xx = i % ST_PW, rr = (i / ST_PW) % ST_PH, c = i / (ST_PW * ST_PH);
        const int y = yi0 + rr, x = xi0 + xx;
        float v = 0.f;
        if (y >= 0 && y < p.H && x >= 0 && x < p.W) v = __ldg(p.img + (((size_t)n * 3 + c) * p.H + y) * p.W + x);
        dst[i] = v;
      }
    };
    if ((int)blockIdx.x < p.tiles) load_patch(blockIdx.x, patch);
    named_bar_sync(1, ST_PRODUCERS);
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      const float* pt = patch + (it & 1) * PATCH_LD;
      // next tile's patch goes into the other buffer while this one is expanded (its readers finished two
      // iterations ago: the named barrier at the end of every iteration orders them)
      const int next = tile + gridDim.x;
      if (next < p.tiles) load_patch(next, patch + ((it + 1) & 1) * PATCH_LD);
      mbarrier_wait(&empty[stage], phase ^ 1);
      unsigned char* a0 = sA + (size_t)stage * ST_A_BYTES + t * 128;
      const float* base = pt + (2 * ty) * ST_PW + 2 * tx;
#pragma unroll
      for (int grp = 0; grp < 19; ++grp) {       // 19 chunks hold k = 0..151 (k >= 147 zero); constant offsets
        if ((grp & 1) != part) continue;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int k = grp * 8 + e;
          const int c = k % 3, kk = k / 3;
          v[e] = k < 147 ? base[c * (ST_PH * ST_PW) + (kk / 7) * ST_PW + (kk % 7)] : 0.f;
        }
        uint4 q;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
        for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
        *reinterpret_cast<uint4*>(a0 + (grp >> 3) * ST_A_KB_BYTES + (((grp & 7) ^ (t & 7)) << 4)) = q;
      }
      fence_proxy_async();                        // generic-proxy smem writes -> visible to tcgen05 (async proxy)
      mbarrier_arrive(&full[stage]);
      if (++stage == ST_STAGES) { stage = 0; phase ^= 1; }
      named_bar_sync(1, ST_PRODUCERS);            // patch double buffer hand-over
    }
  } else if (warp == ST_MMA_WARP) {
    // ===================================== MMA issuer =====================================================
    if (lane == 0) {
      mbarrier_expect_tx(bbar, (uint32_t)ST_B_BYTES);
      for (int kb = 0; kb < ST_KB; ++kb) tma_load_2d(sB + kb * ST_B_KB_BYTES, &tmB, bbar, kb * 64, 0);
      mbarrier_wait(bbar, 0);
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NCOLS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbarrier_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        mbarrier_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 128);
        const uint32_t a_addr = s2u(sA + (size_t)stage * ST_A_BYTES), b_addr = s2u(sB);
#pragma unroll
        for (int kb = 0; kb < ST_KB; ++kb) {
          const uint64_t da = smem_desc_k128(a_addr + kb * ST_A_KB_BYTES), db = smem_desc_k128(b_addr + kb * ST_B_KB_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(&empty[stage]);
        umma_commit(&tfull[acc]);
        if (++stage == ST_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================================== epilogue: scale/shift + ReLU, NHWC bf16 rows ====================
    const int ew = warp & 3;                    // TMEM lane quadrant of this warp (warps 9..16 -> 1,2,3,0,1,2,3,0)
    const int grp_e = (warp - ST_MMA_WARP - 1) >> 2;   // warp group 0 / 1: the 32-column half of every network it drains
    const int m = ew * 32 + lane;               // tile row
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int txi = tile % p.tiles_x;
      const int r = tile / p.tiles_x;
      const int tyi = r % p.tiles_y, n = r / p.tiles_y;
      const int y = tyi * ST_TH + (m >> 4), x = txi * ST_TW + (m & 15);
      const bool ok = y < p.HO && x < p.WO;
      const size_t pix = ((size_t)n * p.HO + y) * p.WO + x;
      mbarrier_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * 128);
      for (int net = 0; net < p.nets; ++net) {
        {
          const int half = grp_e;
          uint32_t rr[32];
          tmem_ld32(tbase + net * 64 + half * 32, rr);
          if (ok) {
            __nv_bfloat16* o = p.out[net] + pix * 64 + half * 32;
            const float* sc = p.scale[net] + half * 32;
            const float* sh = p.shift[net] + half * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const float4 s0 = __ldg(reinterpret_cast<const float4*>(sc + j)), s1 = __ldg(reinterpret_cast<const float4*>(sc + j + 4));
              const float4 h0 = __ldg(reinterpret_cast<const float4*>(sh + j)), h1 = __ldg(reinterpret_cast<const float4*>(sh + j + 4));
              const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
              const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
              uint4 q;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float v0 = fmaxf(__uint_as_float(rr[j + 2 * e]) * sv[2 * e] + hv[2 * e], 0.f);
                const float v1 = fmaxf(__uint_as_float(rr[j + 2 * e + 1]) * sv[2 * e + 1] + hv[2 * e + 1], 0.f);
                h[e] = __floats2bfloat162_rn(v0, v1);
              }
              *reinterpret_cast<uint4*>(o + j) = q;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbarrier_arrive(&tempty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == ST_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}

typedef CUresult (*PFN_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encode stem_get_encode() {
  static PFN_encode fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = (PFN_encode)sym;
  }
  return fn;
}

}  // namespace
}  // namespace ldb

using namespace ldb;

// weights: [64*nets][192] bf16, row = output channel (network 0 first), K = (r*7+s)*3+c zero-padded from 147.
extern "C" int ldb200_stem_tc(const float* img, int N, int H, int W, const void* weights, int nets,
                              const float* scale0, const float* shift0, void* out0, const float* scale1,
                              const float* shift1, void* out1, void* stream_) {
  LDB_REQUIRE(img && weights && scale0 && shift0 && out0, LDB_EINVAL, "stem_tc: null pointer");
  LDB_REQUIRE(nets == 1 || (nets == 2 && scale1 && shift1 && out1), LDB_EINVAL, "stem_tc: nets must be 1 or 2");
  LDB_REQUIRE(N >= 1 && H >= 1 && W >= 1, LDB_EINVAL, "stem_tc: shape");
  LDB_REQUIRE(((uintptr_t)weights & 15) == 0 && ((uintptr_t)out0 & 15) == 0 && (!out1 || ((uintptr_t)out1 & 15) == 0),
              LDB_EINVAL, "stem_tc: weights / outputs must be 16-byte aligned");
  LDB_REQUIRE((((uintptr_t)scale0 | (uintptr_t)shift0 | (uintptr_t)scale1 | (uintptr_t)shift1) & 15) == 0, LDB_EINVAL,
              "stem_tc: scale / shift must be 16-byte aligned");
  StemParams p;
  memset(&p, 0, sizeof(p));
  p.img = img; p.N = N; p.H = H; p.W = W;
  p.HO = (H - 1) / 2 + 1; p.WO = (W - 1) / 2 + 1;       // floor((H + 6 - 7) / 2) + 1
  p.tiles_y = (p.HO + ST_TH - 1) / ST_TH; p.tiles_x = (p.WO + ST_TW - 1) / ST_TW;
  const long long tiles = (long long)N * p.tiles_y * p.tiles_x;
  LDB_REQUIRE(tiles < (1ll << 31), LDB_ENOSUP, "stem_tc: too many tiles");
  p.tiles = (int)tiles;
  p.nets = nets;
  p.scale[0] = scale0; p.shift[0] = shift0; p.out[0] = (__nv_bfloat16*)out0;
  p.scale[1] = scale1; p.shift[1] = shift1; p.out[1] = (__nv_bfloat16*)out1;
  PFN_encode enc = stem_get_encode();
  LDB_REQUIRE(enc, LDB_ENOSUP, "cuTensorMapEncodeTiled not available from the driver");
  CUtensorMap tmB;
  {
    cuuint64_t dims[2] = {(cuuint64_t)ST_K, (cuuint64_t)(64 * nets)};
    cuuint64_t strides[1] = {(cuuint64_t)ST_K * 2};
    cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};     // rows beyond 64*nets are out of bounds -> zero fill
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(weights), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LDB_REQUIRE(r == CUDA_SUCCESS, LDB_EINVAL, "stem_tc: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  }
  static bool attr = false;
  if (!attr) {
    LDB_CUDA(cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM));
    attr = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.tiles < sms ? p.tiles : sms;
  stem_tc_kernel<<<grid, ST_THREADS, ST_SMEM, (cudaStream_t)stream_>>>(tmB, p);
  LDB_CUDA(cudaGetLastError());
  return 0;
}
