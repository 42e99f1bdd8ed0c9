// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM), operands staged by TMA.
//
// Replaces every `nn.Conv2d` the reference's hot path runs through cuDNN
// (mmdet/models/backbones/resnet.py:65-92,260-299 bottleneck/basic-block convs, necks/fpn.py:170-221
// lateral/output/extra convs, dense_heads/gfl_head.py:161-183 tower + gfl_cls/gfl_reg convs) and,
// with rotated weights, their data-gradient (autograd's conv backward-data).
//
//   out[n, y, x, co] = epilogue( sum_{tap} sum_{ci} A[n, y*sy + dy_tap, x*sx + dx_tap, ci] * W[co, tap, ci] )
//
// * activations are NHWC bf16; one CTA tile = TH x TW output pixels (TH*TW = 128 = UMMA M) of one
//   image x BN output channels.  For every tap and every 64-channel slice the A operand is ONE 4-D
//   TMA box {64 ch, TW, TH, 1} at the shifted coordinate: the box lands in shared memory as 128 rows
//   x 128 B, which is exactly the K-major SWIZZLE_128B canonical UMMA layout, padding comes for free
//   from TMA out-of-bounds zero fill, stride-2 convs use the tensor map's element strides.  1x1 convs
//   use the same map with the pixels flattened ({C, N*H*W, 1, 1}).
// * weights are [Cout][taps*Cin] bf16 (K-major), one 2-D TMA box {64, BN} per k-block.
// * warp-specialised persistent kernels: warp 0 = TMA producer, warp 1 = single-thread tcgen05.mma issuer
//   (kind::f16, bf16 x bf16 -> fp32 in TMEM, K=16 per instruction), warp 2 = TMEM allocator, warps 4-11 = epilogue
//   (tcgen05.ld 32x32b -> registers -> fused scale/bias (folded BN, conv bias, Scale), residual / nearest-upsample
//   add, ReLU or ReLU-mask, optional accumulate -> NHWC).  Two TMEM accumulator stages so the epilogue of tile i
//   overlaps the main loop of tile i+1.
// * three kernels share that skeleton:
//     conv2_tc_kernel          2-CTA clusters (cta_group::2, M=256 x N=256 pair tiles): 256-wide output tiles
//     conv_tc_kernel<BN,false> 1-CTA (M=128 x N=BN): narrower layers
//     conv_tc_kernel<BN,true>  1-CTA, two CTAs per SM: reductions of <= 4 k-blocks (HBM-bound 1x1 "expand" convs)
//   and one epilogue (`epilogue_tile`) with two output paths: TMA-staged (32 pixel x 64 channel SWIZZLE_128B boxes,
//   residual / mask TMA-loaded, result TMA-stored) or direct 16-byte stores (fp32 / strided / accumulating outputs).
// * wgrad2_tc_kernel / wgrad_tc_kernel<BN>: weight gradient, both operands MN-major from the same TMA boxes,
//   split-K over pixel tiles, fp32 vector reductions into the parameter-gradient buffer.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace ldb {

constexpr int CONV_THREADS = 384;       // wgrad kernels: 4 control warps + 8 epilogue warps
constexpr int TILE_M = 128;
constexpr int KCH = 64;                        // channels per k-block (128 B of bf16)
constexpr int A_TILE_BYTES = TILE_M * KCH * 2;  // 16 KB

constexpr int CONV_MAX_SEG = 5;   // pyramid levels batched into one launch (shared weights, one tile list)

struct ConvSeg {            // geometry of one segment (= one FPN level of a level-shared head conv, or the whole conv)
  int N, HO, WO;            // tile space = logical output pixels
  int TH, TW, tiles_y, tiles_x, tw_shift;
  uint32_t dx_mul, dx_sh, dy_mul, dy_sh;   // fastdiv constants of tiles_x / tiles_y (host: make_div)
  // output addressing: element offset of pixel (n,y,x) = ((n*OH + y*oys + oyo)*OW + x*oxs + oxo)*ldc
  int OH, OW, oys, oyo, oxs, oxo;
  void* out;
  const void* add_src;      // optional add (bf16, or fp32 when p.add_f32), nearest-resampled from [N, add_H, add_W, add_ld]
  int add_H, add_W;
  const void* mask_src;     // optional: multiply by (mask_src > 0), same addressing as out (bf16 / fp32 when p.mask_f32, ld = mask_ld)
  const float* pre_src;     // optional fp32 tensor addressed like `out`: added to the raw accumulator BEFORE scale / bias
                            // (second pass of a two-pass split-bf16 conv: main plane product first, corrections after)
  float* gn_stats;          // optional [N, Cout/8, 2]: the epilogue accumulates sum / sum of squares of the output per
                            // (image, group of 8 channels) -- GroupNorm(32) statistics without a pass over the output
  int pf_add, pf_mask;      // L2-prefetch the residual / mask tile with TMA while the main loop runs
  int tl_add, tl_mask;      // TMA epilogue: residual / mask boxes are TMA-loaded into the warp's staging buffer
};

struct ConvParams {
  int nseg;
  int seg_tile0[CONV_MAX_SEG + 1];   // first tile of each segment; [nseg] = total tiles
  int seg_mt0[CONV_MAX_SEG + 1];     // the same in pixel tiles (seg_tile0 / tiles_n)
  uint32_t dn_mul, dn_sh;            // fastdiv constants of tiles_n
  uint32_t dh_mul, dh_sh;            // ... of the pair count per output-channel tile of the 2-CTA kernel, ceil(pixel tiles / 2)
  ConvSeg seg[CONV_MAX_SEG];
  int tiles_n;
  int Cout, Cin;
  int ntaps;
  int tap_dx[9], tap_dy[9], tap_wk[9];
  int sx, sy;               // input coordinate = out coordinate * s + tap offset
  int ldc;
  int out_f32;
  const float* scale;       // [Cout] or null (1)
  const float* bias;        // [Cout] or null (0)
  int add_ld, mask_ld;
  int relu;
  int accumulate;           // result += previous out (before ReLU / mask)
  int vec_ok;               // leading dimensions allow 16-byte vector accesses
  int tma_epi;              // epilogue stages 32-pixel x 64-channel boxes in shared memory and TMA-stores them
  int nt_inner;             // 1-CTA kernel: a CTA walks all output-channel tiles of one pixel tile back to back
  // split-bf16 (fp32-accurate) mode: the input is `planes` stacked bf16 tensors [planes][N,H,W,ld] whose sum is the fp32
  // activation, the weights [Cout][planes][w_plane]; the reduction runs over ncombo (activation plane, weight plane)
  // pairs per tap and channel slice, all into the same fp32 TMEM accumulator.  planes = 1, ncombo = 1: plain bf16.
  int planes, ncombo, w_plane;
  int comb_a[6], comb_w[6];
  int add_f32, mask_f32;    // residual / mask tensors are fp32 (fp32 activations of the split mode)
  // flat 1x1 convs with wide input rows: the A operand arrives as 128-byte pieces of Cin*2-byte rows, one piece per
  // k-block -- DRAM sees 128 scattered 128-byte reads per TMA box (ncu r02: 1x1 1024->256 at 25 % of HBM peak with
  // exactly the algorithmic bytes).  a_pf: one tile AHEAD, the producer pulls the next tile's rows into L2 in
  // 512-byte-per-row boxes (maps.ap), so DRAM sees long bursts and the operand loads hit L2.
  int a_pf;
};

struct ConvMaps {           // per-segment TMA descriptors: A operand, residual prefetch, mask prefetch,
  CUtensorMap a[CONV_MAX_SEG], r[CONV_MAX_SEG], m[CONV_MAX_SEG];
  // TMA epilogue: output store, residual load, mask load (boxes of EPI_BOX_PIX pixels x 64 channels, SWIZZLE_128B)
  CUtensorMap o[CONV_MAX_SEG], rl[CONV_MAX_SEG], ml[CONV_MAX_SEG];
  CUtensorMap ap[CONV_MAX_SEG];   // A-operand L2 prefetch: SWIZZLE_NONE boxes of APF_CH channels x 128 pixels (flat convs)
};
constexpr int APF_CH = 256;
constexpr int EPI_BOX_PIX = 32;                       // one epilogue warp = 32 TMEM lanes = 32 output pixels
constexpr int EPI_BOX_BYTES = EPI_BOX_PIX * 64 * 2;   // 4 KB staging buffer per epilogue warp

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t s2u(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarrier_init(uint64_t* b, int cnt) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2u(b)), "r"(cnt));
}
__device__ __forceinline__ void mbarrier_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t}" ::"r"(s2u(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbarrier_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2u(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarrier_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s2u(b)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(s2u(dst)), "l"(map), "r"(s2u(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(s2u(dst)), "l"(map), "r"(s2u(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1),
               "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
               "r"(s2u(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// Programmatic dependent launch: let the next kernel in the stream start its prologue (barrier init, TMEM
// allocation, descriptor prefetch) while this grid drains, and wait for the previous grid's memory before the
// first global access (no-ops when the kernel was not launched with the PDL attribute).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s2u(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// explicit shared-window vector accesses (the generic-pointer form compiles to LD.E / ST.E)
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// rows of 128 B, 8-row swizzle atoms 1024 B apart (SBO), version 1, layout type 2.
__device__ __forceinline__ uint64_t smem_desc_k128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)1 << 16;                 // LBO (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;       // SBO
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}

// SHORTK: convolutions whose whole reduction is <= 4 k-blocks (1x1 convs with Cin <= 256: the bottleneck "expand"
// convs and their dgrads).  They are bound by the epilogue's memory traffic, not by the MMAs: run TWO CTAs per SM
// (2 pipeline stages, BN <= 128 so that both fit the 512 TMEM columns) so that one CTA's epilogue loads/stores
// overlap the other's.
template <int BN, bool SHORTK = false> struct ConvCfg {
  static constexpr int B_TILE_BYTES = BN * KCH * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int STAGES = SHORTK ? 2 : ((BN == 256) ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int MIN_BLOCKS = SHORTK ? 2 : 1;
  static_assert(!SHORTK || BN <= 128, "two resident CTAs need 2 * (2*BN) <= 512 TMEM columns");
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int STAGING_BYTES = BN >= 128 ? 4 * 2 * EPI_BOX_BYTES : 0;   // TMA-epilogue staging, one box per warp
  // epilogue warp groups of 4 warps (one per TMEM lane quadrant); each group owns BN/EPI_GROUPS columns.
  // Four groups (4 warps per SM sub-partition) hide the latency of the residual / mask loads.
  static constexpr int EPI_GROUPS = 2;   // 8 epilogue warps (16 measured slower: 67.6 vs 65.2 ms/step)
  static constexpr int THREADS = 128 + 128 * EPI_GROUPS;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static_assert((SMEM + 1024) * MIN_BLOCKS <= 228 * 1024 && SMEM <= 227 * 1024, "shared memory budget");
  // kind::f16 instruction descriptor: D=f32, A=B=bf16, K-major both, N=BN, M=128
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                    ((uint32_t)(TILE_M >> 4) << 24);
};

__device__ __forceinline__ size_t out_offset(const ConvSeg& p, int n, int y, int x) {
  return (((size_t)n * p.OH + (size_t)(y * p.oys + p.oyo)) * p.OW + (size_t)(x * p.oxs + p.oxo));
}

// Tile schedule of the 1-CTA kernel: the i-th work item of CTA `b`.  nt_inner: the CTA takes pixel tile
// b + (i / tiles_n) * grid and walks its tiles_n output-channel tiles consecutively, so the A operand of the later ones
// comes out of L2 a few microseconds after the first read (with the channel tile fastest ACROSS CTAs, the CTAs sharing a
// pixel tile drift apart and big maps re-read A from HBM: 975 MB instead of 688 MB on the layer-1 expand conv).
struct TileCoord { int sg, nt, tx, ty, n; };
// n / d for 0 <= n < 2^31 with the host-made constants of d (make_div): three instructions instead of the ~25 of an
// integer division -- the tile decode runs once per tile in every warp role (ncu r02: 303 instructions per tile in the
// epilogue warps, three divisions among them)
__device__ __forceinline__ int fastdiv(int n, uint32_t mul, uint32_t sh) {
  return (int)((__umulhi((uint32_t)n, mul) + (uint32_t)n) >> sh);
}
// pixel tile mt of segment sg -> (tx, ty, n)
__device__ __forceinline__ void decode_pixel_tile(const ConvSeg& S, int sp, int& tx, int& ty, int& n) {
  const int q = fastdiv(sp, S.dx_mul, S.dx_sh);
  tx = sp - q * S.tiles_x;
  n = fastdiv(q, S.dy_mul, S.dy_sh);
  ty = q - n * S.tiles_y;
}
__device__ __forceinline__ bool decode_tile(const ConvParams& p, int b, int grid, int i, TileCoord& c) {
  int mt;
  if (p.nt_inner) {
    const int r = fastdiv(i, p.dn_mul, p.dn_sh);
    mt = b + r * grid;
    c.nt = i - r * p.tiles_n;
  } else {
    const int t = b + i * grid;
    mt = fastdiv(t, p.dn_mul, p.dn_sh);
    c.nt = t - mt * p.tiles_n;
  }
  if (mt >= p.seg_mt0[p.nseg]) return false;
  int sg = 0;
  while (sg + 1 < p.nseg && mt >= p.seg_mt0[sg + 1]) ++sg;
  const ConvSeg& S = p.seg[sg];
  c.sg = sg;
  decode_pixel_tile(S, mt - p.seg_mt0[sg], c.tx, c.ty, c.n);
  return true;
}

// ---------------------------------------------------------------------------------------------
// Epilogue of ONE output tile for one epilogue warp (shared by the 1-CTA and 2-CTA kernels): warp (ew, half) owns
// TMEM lanes [32*ew, +32) = 32 tile rows and the tile's columns [half*GCOLS, +GCOLS).  Per 32-column chunk:
// tcgen05.ld (software-pipelined one chunk ahead), per-channel scale/bias (float4 loads), residual / accumulate /
// ReLU / mask, then either
//   * direct path: 16-byte global stores along the thread's own output row (any layout / dtype), or
//   * TMA path (p.tma_epi; bf16, identity pixel mapping, Cout % 64 == 0): the warp's 32 pixels x 64 channels go
//     through a 4 KB SWIZZLE_128B staging box -- residual and ReLU-mask boxes arrive by TMA load, the result leaves by
//     TMA store -- so global memory only sees full 128-byte lines instead of 32 scattered 16-byte accesses per
//     warp instruction (the scattered form held the 1x1 "expand" convs at ~40 % of the HBM roofline).
// Leaves the TMEM accumulator drained (caller signals tmem-empty).
// EXT selects the epilogue features compiled in (the plain bf16 hot path carries none of them):
//   bit 0: GroupNorm statistics (S.gn_stats)      bit 1: fp32 extras of the split mode (pre_src, fp32 add / mask)
template <int BN, bool SHORTK, int EPI_GROUPS, int EXT>
__device__ __forceinline__ void epilogue_tile(const ConvParams& p, const ConvMaps& maps, const ConvSeg& S, int sg, int nt,
                                              int tx, int ty, int n, bool live, uint32_t tmem_acc, uint64_t* tfull_bar,
                                              uint32_t acc_phase, int ew, int half, int lane, unsigned char* ebuf,
                                              uint64_t* rbar, uint32_t& rph) {
  constexpr int GCOLS = BN / EPI_GROUPS;   // columns per warp
  constexpr int NCH = GCOLS / 32;          // 32-column chunks per warp
  const bool tma = (BN >= 128) && p.tma_epi;   // warp-uniform
  const int m = ew * 32 + lane;
  const int y = ty * S.TH + (m >> S.tw_shift), x = tx * S.TW + (m & (S.TW - 1));
  const bool pix_ok = live && (y < S.HO) && (x < S.WO);
  const size_t pix = pix_ok ? out_offset(S, n, y, x) : 0;
  const __nv_bfloat16* addp = nullptr;
  const float* addpf = nullptr;
  if (S.add_src && pix_ok && !(tma && S.tl_add)) {
    const int ay = (int)(((long long)y * S.add_H) / S.HO), ax = (int)(((long long)x * S.add_W) / S.WO);
    const size_t aoff = (((size_t)n * S.add_H + ay) * S.add_W + ax) * (size_t)p.add_ld;
    if ((EXT & 2) && p.add_f32) addpf = (const float*)S.add_src + aoff;
    else addp = (const __nv_bfloat16*)S.add_src + aoff;
  }
  const int col0 = nt * BN + half * GCOLS;
  const uint32_t tbase = tmem_acc + ((uint32_t)(ew * 32) << 16) + (uint32_t)(half * GCOLS);
  // TMA path: first pixel of this warp's box, this lane's 128-byte row in the staging box and its swizzle phase
  const int bx = tx * S.TW + ((ew * 32) & (S.TW - 1)), by = ty * S.TH + ((ew * 32) >> S.tw_shift);
  unsigned char* erow = ebuf + lane * 128;
  const int sw = lane & 7;
  uint32_t mb0 = 0xffffffffu, mb1 = 0xffffffffu;   // ReLU-mask bits of this lane's 64 box columns
  // staging buffer free again (the previous TMA store has read it), then mask bits and the residual box
  auto acquire = [&](int colb) {
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
    if (S.tl_mask) {
      if (lane == 0) {
        if (S.tl_add) tma_prefetch_4d(&maps.rl[sg], colb, bx, by, n);   // residual box follows: start its fetch now
        mbarrier_expect_tx(rbar, EPI_BOX_BYTES);
        tma_load_4d(ebuf, &maps.ml[sg], rbar, colb, bx, by, n);
      }
      mbarrier_wait(rbar, rph);
      rph ^= 1;
      mb0 = mb1 = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 mv = *reinterpret_cast<const uint4*>(erow + ((q ^ sw) << 4));
        const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&mv);
        uint32_t bits = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) bits |= (__bfloat162float(hb[e]) > 0.f ? 1u : 0u) << e;
        if (q < 4) mb0 |= bits << (8 * q);
        else mb1 |= bits << (8 * (q - 4));
      }
      fence_proxy_async();
      __syncwarp();
    }
    if (S.tl_add) {
      if (lane == 0) {
        mbarrier_expect_tx(rbar, EPI_BOX_BYTES);
        tma_load_4d(ebuf, &maps.rl[sg], rbar, colb, bx, by, n);
      }
      mbarrier_wait(rbar, rph);
      rph ^= 1;
    }
  };
  if (tma && col0 < p.Cout) acquire(col0);   // overlaps this tile's main loop
  mbarrier_wait(tfull_bar, acc_phase);
  tc_fence_after();
  // SHORTK (two CTAs per SM, 85-register budget): one TMEM buffer, chunks strictly one after the other -- the
  // second CTA provides the latency hiding.  Otherwise the next chunk's tcgen05.ld is issued one chunk ahead.
  uint32_t rbuf[SHORTK ? 1 : 2][32];
  if (!SHORTK && col0 < p.Cout) tmem_ld32_nowait(tbase, rbuf[0]);
#pragma unroll(SHORTK ? 1 : NCH)
  for (int ch = 0; ch < NCH; ++ch) {
    const int cbase = col0 + ch * 32;
    if (cbase >= p.Cout) break;   // warp-uniform
    if (tma && ch > 0 && (ch & 1) == 0) acquire(cbase);   // next 64-column box of this warp (BN = 256)
    if (SHORTK) tmem_ld32_nowait(tbase + ch * 32, rbuf[0]);
    tmem_ld_wait();
    if (!SHORTK && ch + 1 < NCH && cbase + 32 < p.Cout) tmem_ld32_nowait(tbase + (ch + 1) * 32, rbuf[(ch + 1) & 1]);
    const uint32_t* r = rbuf[SHORTK ? 0 : (ch & 1)];
    float gs[8];   // GroupNorm partials of this lane's pixel: {sum, sum of squares} of the chunk's four 8-channel groups
#pragma unroll
    for (int i = 0; i < 8; ++i) gs[i] = 0.f;
    if (pix_ok) {
      const bool full = (cbase + 32 <= p.Cout) && p.vec_ok;   // warp-uniform fast path
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if ((EXT & 2) && S.pre_src) {
        const float* pr = S.pre_src + pix * (size_t)p.ldc + cbase;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 o4 = __ldg(reinterpret_cast<const float4*>(pr + j));
            v[j] += o4.x; v[j + 1] += o4.y; v[j + 2] += o4.z; v[j + 3] += o4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (cbase + j < p.Cout) v[j] += __ldg(pr + j);
        }
      }
      if (full) {
        if (p.scale) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.scale + cbase + j));
            v[j] *= s4.x; v[j + 1] *= s4.y; v[j + 2] *= s4.z; v[j + 3] *= s4.w;
          }
        }
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + cbase + j));
            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
          }
        }
        if (tma && S.tl_add) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const uint4 q = *reinterpret_cast<const uint4*>(erow + ((((ch & 1) * 4 + (j >> 3)) ^ sw) << 4));
            const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(hh[e]);
              v[j + 2 * e] += f.x;
              v[j + 2 * e + 1] += f.y;
            }
          }
        } else if ((EXT & 2) && addpf) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 a4 = __ldg(reinterpret_cast<const float4*>(addpf + cbase + j));
            v[j] += a4.x; v[j + 1] += a4.y; v[j + 2] += a4.z; v[j + 3] += a4.w;
          }
        } else if (addp) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(addp + cbase + j));
            const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(hh[e]);
              v[j + 2 * e] += f.x;
              v[j + 2 * e + 1] += f.y;
            }
          }
        }
        if (p.accumulate) {
          if (p.out_f32) {
            const float* o = (const float*)S.out + pix * (size_t)p.ldc + cbase;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 o4 = *reinterpret_cast<const float4*>(o + j);
              v[j] += o4.x; v[j + 1] += o4.y; v[j + 2] += o4.z; v[j + 3] += o4.w;
            }
          } else {
            const __nv_bfloat16* o = (const __nv_bfloat16*)S.out + pix * (size_t)p.ldc + cbase;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const uint4 q = *reinterpret_cast<const uint4*>(o + j);
              const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(hh[e]);
                v[j + 2 * e] += f.x;
                v[j + 2 * e + 1] += f.y;
              }
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (tma && S.tl_mask) {
          const uint32_t word = (ch & 1) ? mb1 : mb0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (!((word >> j) & 1u)) v[j] = 0.f;
        } else if ((EXT & 2) && S.mask_src && p.mask_f32) {
          const float* mk = (const float*)S.mask_src + pix * (size_t)p.mask_ld + cbase;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 m4 = __ldg(reinterpret_cast<const float4*>(mk + j));
            if (!(m4.x > 0.f)) v[j] = 0.f;
            if (!(m4.y > 0.f)) v[j + 1] = 0.f;
            if (!(m4.z > 0.f)) v[j + 2] = 0.f;
            if (!(m4.w > 0.f)) v[j + 3] = 0.f;
          }
        } else if (S.mask_src) {
          const __nv_bfloat16* mk = (const __nv_bfloat16*)S.mask_src + pix * (size_t)p.mask_ld + cbase;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(mk + j));
            const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&q);
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (!(__bfloat162float(hb[e]) > 0.f)) v[j + e] = 0.f;
          }
        }
        if ((EXT & 1) && S.gn_stats) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              gs[2 * q] += v[8 * q + e];
              gs[2 * q + 1] += v[8 * q + e] * v[8 * q + e];
            }
          }
        }
        if (tma) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 q;
            __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
            for (int e = 0; e < 4; ++e) hh[e] = __floats2bfloat162_rn(v[j + 2 * e], v[j + 2 * e + 1]);
            *reinterpret_cast<uint4*>(erow + ((((ch & 1) * 4 + (j >> 3)) ^ sw) << 4)) = q;
          }
        } else if (p.out_f32) {
          float* o = (float*)S.out + pix * (size_t)p.ldc + cbase;
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          __nv_bfloat16* o = (__nv_bfloat16*)S.out + pix * (size_t)p.ldc + cbase;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 q;
            __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
            for (int e = 0; e < 4; ++e) hh[e] = __floats2bfloat162_rn(v[j + 2 * e], v[j + 2 * e + 1]);
            *reinterpret_cast<uint4*>(o + j) = q;
          }
        }
      } else {
        // ragged tail (Cout not a multiple of 32, or unaligned leading dimension): scalar path
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int c = cbase + j;
          if (c >= p.Cout) continue;
          float vv = v[j];
          if (p.scale) vv *= __ldg(p.scale + c);
          if (p.bias) vv += __ldg(p.bias + c);
          if (addp) vv += __bfloat162float(addp[c]);
          if ((EXT & 2) && addpf) vv += addpf[c];
          if (p.accumulate)
            vv += p.out_f32 ? ((const float*)S.out)[pix * (size_t)p.ldc + c]
                            : __bfloat162float(((const __nv_bfloat16*)S.out)[pix * (size_t)p.ldc + c]);
          if (p.relu) vv = fmaxf(vv, 0.f);
          if (S.mask_src) {
            const float mv = ((EXT & 2) && p.mask_f32) ? ((const float*)S.mask_src)[pix * (size_t)p.mask_ld + c]
                                        : __bfloat162float(((const __nv_bfloat16*)S.mask_src)[pix * (size_t)p.mask_ld + c]);
            if (!(mv > 0.f)) vv = 0.f;
          }
          if (p.out_f32) ((float*)S.out)[pix * (size_t)p.ldc + c] = vv;
          else ((__nv_bfloat16*)S.out)[pix * (size_t)p.ldc + c] = __float2bfloat16_rn(vv);
        }
      }
    }
    if ((EXT & 1) && S.gn_stats && live && cbase + 32 <= p.Cout) {   // warp-uniform
      // 8 partials x 32 pixels -> 8 sums with 9 shuffles: each step halves the values a lane carries (the half it
      // keeps is chosen by one lane bit), the last two steps are a plain butterfly; lanes with (lane & 3) == 0 end
      // up with value index bit4*4 + bit3*2 + bit2
      {
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float send = hi ? gs[i] : gs[i + 4], keep = hi ? gs[i + 4] : gs[i];
          gs[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
      }
      {
        const bool hi = lane & 8;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float send = hi ? gs[i] : gs[i + 2], keep = hi ? gs[i + 2] : gs[i];
          gs[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
      }
      {
        const bool hi = lane & 4;
        const float send = hi ? gs[0] : gs[1], keep = hi ? gs[1] : gs[0];
        gs[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      gs[0] += __shfl_xor_sync(0xffffffffu, gs[0], 2);
      gs[0] += __shfl_xor_sync(0xffffffffu, gs[0], 1);
      if ((lane & 3) == 0) {
        const int idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        atomicAdd(S.gn_stats + ((size_t)n * (p.Cout >> 3) + (cbase >> 3) + (idx >> 1)) * 2 + (idx & 1), gs[0]);
      }
    }
    if (tma && (ch & 1)) {   // 64 columns staged: hand the box to the TMA (rows outside the tensor are clipped)
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(&maps.o[sg], ebuf, cbase - 32, bx, by, n);
        bulk_commit();
      }
    }
  }
  tmem_ld_wait();
  tc_fence_before();
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// Lean epilogue for the launches that make up almost all of the LD step's epilogue work: TMA-staged bf16 output, every
// 64-column box complete (Cout % BN == 0), no accumulate / fp32 / GroupNorm extras, and the feature set known at
// compile time (FEAT): per-channel scale AND bias (one FFMA), ReLU (on the packed bf16 pair), residual box, ReLU-mask
// box.  ncu (r02, 1x1 256->1024 +residual): the generic epilogue issues 392 instructions per 32-column chunk and 303 per
// tile, ~2.5x the arithmetic it carries (runtime feature tests, generic-space LD/ST, FMUL+FADD), and its warps sit at
// 12 cycles per issued instruction: the HBM-heavy convs are bound by the epilogue's instruction count.
constexpr int EF_AFFINE = 1, EF_RELU = 2, EF_ADD = 4, EF_MASK = 8, EF_BIAS = 16;   // EF_BIAS: bias without scale

template <int BN, bool SHORTK, int EPI_GROUPS, int FEAT>
__device__ __forceinline__ void epilogue_tile_fast(const ConvParams& p, const ConvMaps& maps, const ConvSeg& S, int sg, int nt,
                                                   int tx, int ty, int n, uint32_t tmem_acc, uint64_t* tfull_bar,
                                                   uint32_t acc_phase, int ew, int half, int lane, unsigned char* ebuf,
                                                   uint64_t* rbar, uint32_t& rph) {
  constexpr int GCOLS = BN / EPI_GROUPS;   // columns per warp
  constexpr int NBOX = GCOLS / 64;         // 64-column staging boxes per warp
  static_assert(GCOLS % 64 == 0, "the fast epilogue moves whole 64-column boxes");
  const int col0 = nt * BN + half * GCOLS;
  const uint32_t tbase = tmem_acc + ((uint32_t)(ew * 32) << 16) + (uint32_t)(half * GCOLS);
  const int bx = tx * S.TW + ((ew * 32) & (S.TW - 1)), by = ty * S.TH + ((ew * 32) >> S.tw_shift);
  const uint32_t erow = s2u(ebuf) + (uint32_t)lane * 128u;
  const uint32_t sw = (uint32_t)lane & 7u;
  uint32_t rbuf[SHORTK ? 1 : 2][32];
#pragma unroll
  for (int b = 0; b < NBOX; ++b) {
    const int colb = col0 + b * 64;
    uint32_t mb0 = 0xffffffffu, mb1 = 0xffffffffu;   // ReLU-mask bits of this lane's 64 box columns
    // staging box free again (the previous TMA store has read it), then the mask bits and the residual box
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
    if (FEAT & EF_MASK) {
      if (lane == 0) {
        if (FEAT & EF_ADD) tma_prefetch_4d(&maps.rl[sg], colb, bx, by, n);   // residual box follows: start its fetch now
        mbarrier_expect_tx(rbar, EPI_BOX_BYTES);
        tma_load_4d(ebuf, &maps.ml[sg], rbar, colb, bx, by, n);
      }
      mbarrier_wait(rbar, rph);
      rph ^= 1;
      mb0 = mb1 = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 mv = lds128(erow + (((uint32_t)q ^ sw) << 4));
        const uint32_t w4[4] = {mv.x, mv.y, mv.z, mv.w};
        uint32_t bits = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {   // bf16 > 0: sign clear and not zero (NaN counts as positive: never produced by ReLU)
          const uint32_t lo = w4[e] & 0xffffu, hi = w4[e] >> 16;
          bits |= ((lo - 1u) < 0x7fffu ? 1u : 0u) << (2 * e);
          bits |= ((hi - 1u) < 0x7fffu ? 1u : 0u) << (2 * e + 1);
        }
        if (q < 4) mb0 |= bits << (8 * q);
        else mb1 |= bits << (8 * (q - 4));
      }
      fence_proxy_async();
      __syncwarp();
    }
    if (FEAT & EF_ADD) {
      if (lane == 0) {
        mbarrier_expect_tx(rbar, EPI_BOX_BYTES);
        tma_load_4d(ebuf, &maps.rl[sg], rbar, colb, bx, by, n);
      }
    }
    if (b == 0) {
      mbarrier_wait(tfull_bar, acc_phase);
      tc_fence_after();
    }
    if (SHORTK || b == 0) tmem_ld32_nowait(tbase + b * 64, rbuf[0]);   // (else: issued one chunk ahead, below)
    if (FEAT & EF_ADD) {
      mbarrier_wait(rbar, rph);
      rph ^= 1;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int ch = 2 * b + h;
      const int cbase = colb + 32 * h;
      if (SHORTK && h == 1) tmem_ld32_nowait(tbase + ch * 32, rbuf[0]);
      tmem_ld_wait();
      if (!SHORTK && ch + 1 < 2 * NBOX) tmem_ld32_nowait(tbase + (ch + 1) * 32, rbuf[(ch + 1) & 1]);
      const uint32_t* r = rbuf[SHORTK ? 0 : (ch & 1)];
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (FEAT & EF_AFFINE) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.scale + cbase + j));
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + cbase + j));
          v[j] = fmaf(v[j], s4.x, b4.x);
          v[j + 1] = fmaf(v[j + 1], s4.y, b4.y);
          v[j + 2] = fmaf(v[j + 2], s4.z, b4.z);
          v[j + 3] = fmaf(v[j + 3], s4.w, b4.w);
        }
      }
      if (FEAT & EF_BIAS) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + cbase + j));
          v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
        }
      }
      if (FEAT & EF_ADD) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const uint4 q = lds128(erow + ((((uint32_t)(h * 4 + (j >> 3))) ^ sw) << 4));
          const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            v[j + 2 * e] += __uint_as_float(w4[e] << 16);
            v[j + 2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u);
          }
        }
      }
      if (FEAT & EF_MASK) {
        const uint32_t word = h ? mb1 : mb0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (!((word >> j) & 1u)) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 q;
        uint32_t* qw = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 2 * e], v[j + 2 * e + 1]);
          if (FEAT & EF_RELU) h2 = __hmax2(h2, __floats2bfloat162_rn(0.f, 0.f));   // max(round(v), 0) == round(max(v, 0))
          qw[e] = *reinterpret_cast<uint32_t*>(&h2);
        }
        sts128(erow + ((((uint32_t)(h * 4 + (j >> 3))) ^ sw) << 4), q);
      }
    }
    // 64 columns staged: hand the box to the TMA (rows outside the tensor are clipped; a dummy tile is all outside)
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_4d(&maps.o[sg], ebuf, colb, bx, by, n);
      bulk_commit();
    }
  }
  tc_fence_before();
  __syncwarp();
}

template <int BN, bool SHORTK, int EXT, int FEAT>
__global__ void __launch_bounds__(ConvCfg<BN, SHORTK>::THREADS, ConvCfg<BN, SHORTK>::MIN_BLOCKS)
conv_tc_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<BN, SHORTK>;
  extern __shared__ unsigned char conv_smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)conv_smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* staging = smem + (size_t)Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = (uint64_t*)(staging + Cfg::STAGING_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* tfull = bars + 2 * Cfg::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
  uint64_t* rbars = tempty + 3;   // one per epilogue warp (TMA-epilogue residual / mask loads)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cchunks = p.Cin / KCH;
  const int kblocks = p.ncombo * p.ntaps * cchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbarrier_init(&full[s], 1);
      mbarrier_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbarrier_init(&tfull[a], 1);
      mbarrier_init(&tempty[a], 4 * Cfg::EPI_GROUPS);   // one arrive per epilogue warp
    }
    for (int w = 0; w < 4 * Cfg::EPI_GROUPS; ++w) mbarrier_init(&rbars[w], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)),
                 "n"(Cfg::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      pdl_wait();   // inputs are produced by the previous kernel
      int stage = 0;
      uint32_t phase = 0;
      TileCoord tc;
      auto prefetch_a = [&](const TileCoord& t) {     // flat mode: pixel tile t.tx of segment t.sg, all channels
        if (t.nt != 0) return;                        // the CTAs sharing a pixel tile: one of them prefetches
        for (int c = 0; c < p.Cin; c += APF_CH) tma_prefetch_4d(&maps.ap[t.sg], c, t.tx * TILE_M, 0, 0);
      };
      if (p.a_pf && decode_tile(p, blockIdx.x, gridDim.x, 0, tc)) prefetch_a(tc);
      for (int i = 0; decode_tile(p, blockIdx.x, gridDim.x, i, tc); ++i) {
        const int sg = tc.sg, nt = tc.nt, tx = tc.tx, ty = tc.ty, n = tc.n;
        const ConvSeg& S = p.seg[sg];
        const int x0 = tx * S.TW * p.sx, y0 = ty * S.TH * p.sy;
        if (p.a_pf) {
          TileCoord nx;
          if (decode_tile(p, blockIdx.x, gridDim.x, i + 1, nx)) prefetch_a(nx);
        }
        // the epilogue will read this tile's residual / ReLU-mask rows a whole main loop from now: pull them into
        // L2 with one TMA prefetch each so those loads see L2 latency instead of HBM latency
        if (S.pf_add) tma_prefetch_4d(&maps.r[sg], nt * BN, tx * S.TW, ty * S.TH, n);
        if (S.pf_mask) tma_prefetch_4d(&maps.m[sg], nt * BN, tx * S.TW, ty * S.TH, n);
        // (plane pair, tap, 64-channel slice): nested counters, no divisions on the producer's critical path
        for (int cmb = 0; cmb < p.ncombo; ++cmb) {
          const int na = n + p.comb_a[cmb] * S.N, wk0 = p.comb_w[cmb] * p.w_plane;
          for (int tap = 0; tap < p.ntaps; ++tap) {
            const int ax = x0 + p.tap_dx[tap], ay = y0 + p.tap_dy[tap], wk = wk0 + p.tap_wk[tap];
            for (int c0 = 0; c0 < p.Cin; c0 += KCH) {
              mbarrier_wait(&empty[stage], phase ^ 1);
              unsigned char* sa = smem + (size_t)stage * Cfg::STAGE_BYTES;
              mbarrier_expect_tx(&full[stage], Cfg::STAGE_BYTES);
              tma_load_4d(sa, &maps.a[sg], &full[stage], c0, ax, ay, na);
              tma_load_2d(sa + A_TILE_BYTES, &tmB, &full[stage], wk + c0, nt * BN);
              if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      TileCoord tc;
      for (; decode_tile(p, blockIdx.x, gridDim.x, it, tc); ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbarrier_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbarrier_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = s2u(smem + (size_t)stage * Cfg::STAGE_BYTES);
          const uint64_t da = smem_desc_k128(sa), db = smem_desc_k128(sa + A_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < KCH / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
            umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), Cfg::IDESC, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);                 // frees the smem slot when the MMAs retire
          if (kb == kblocks - 1) umma_commit(&tfull[acc]);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue ==========================================
    // 4*EPI_GROUPS warps (EPI_GROUPS per SM sub-partition, hiding each other's ALU / TMEM-load / residual-load
    // latency): warp w reads TMEM lanes [32*(w%4), +32) = 32 tile rows, and BN/EPI_GROUPS of the tile's columns.  Per 32-column
    // chunk: tcgen05.ld (software-pipelined one chunk ahead), per-channel scale/bias (float4 loads),
    // residual / accumulate / ReLU / mask, 16-byte stores along the thread's own output row.
    pdl_wait();   // residual / mask / accumulate reads and the output stores touch memory of earlier kernels
    const int ew = warp & 3, half = (warp - 4) >> 2;              // half = column group of this warp
    unsigned char* ebuf = staging + (size_t)(warp - 4) * EPI_BOX_BYTES;   // TMA-epilogue staging box of this warp
    uint64_t* rbar = rbars + (warp - 4);
    uint32_t rph = 0;
    int it = 0;
    TileCoord tc;
    for (; decode_tile(p, blockIdx.x, gridDim.x, it, tc); ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int sg = tc.sg, nt = tc.nt, tx = tc.tx, ty = tc.ty, n = tc.n;
      const ConvSeg& S = p.seg[sg];
      if constexpr (FEAT >= 0)
        epilogue_tile_fast<BN, SHORTK, Cfg::EPI_GROUPS, FEAT>(p, maps, S, sg, nt, tx, ty, n, tmem_base + (uint32_t)(acc * BN),
                                                             &tfull[acc], acc_phase, ew, half, lane, ebuf, rbar, rph);
      else
        epilogue_tile<BN, SHORTK, Cfg::EPI_GROUPS, EXT>(p, maps, S, sg, nt, tx, ty, n, true, tmem_base + (uint32_t)(acc * BN),
                                                   &tfull[acc], acc_phase, ew, half, lane, ebuf, rbar, rph);
      if (lane == 0) mbarrier_arrive(&tempty[acc]);
    }
    if (lane == 0) bulk_wait_all();   // staging buffers must outlive the TMA stores that read them
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS));
  }
}

// ---------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2): a pair of SMs computes one 256 x 256 output tile; every CTA stages its own 128
// pixel rows of A and HALF of the weight tile, so shared-memory fill + operand reads drop from 192 to 128 B/cycle per
// SM (the measured limit of the 1-CTA kernel: 73-79 % tensor-pipe) and L2 weight traffic halves.
struct Conv2Cfg {
  static constexpr int BN = 256;
  static constexpr int B_TILE_BYTES = (BN / 2) * KCH * 2;        // this CTA's half of the weight tile
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;  // 32 KB
  static constexpr int STAGES = 6;
  static constexpr int TMEM_COLS = 512;
  static constexpr int STAGING_BYTES = 4 * 2 * EPI_BOX_BYTES;
  static constexpr int EPI_GROUPS = 2;
  static constexpr int THREADS = 128 + 128 * EPI_GROUPS;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + STAGING_BYTES + 1024 + 256;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                    ((uint32_t)(256 >> 4) << 24);
};
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> leader CTA

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                                int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(s2u(dst)), "l"(map), "r"(s2u(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(s2u(dst)), "l"(map), "r"(s2u(bar) & PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {   // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(s2u(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbarrier_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(s2u(bar) & PEER_MASK) : "memory");
}

template <int EXT, int FEAT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Conv2Cfg::THREADS, 1)
conv2_tc_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ ConvParams p) {
  using Cfg = Conv2Cfg;
  constexpr int BN = Conv2Cfg::BN;
  constexpr bool SHORTK = false;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  extern __shared__ unsigned char conv_smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)conv_smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* staging = smem + (size_t)Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = (uint64_t*)(staging + Cfg::STAGING_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* tfull = bars + 2 * Cfg::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
  uint64_t* rbars = tempty + 3;   // one per epilogue warp (TMA-epilogue residual / mask loads)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cchunks = p.Cin / KCH;
  const int kblocks = p.ncombo * p.ntaps * cchunks;
  // pair-tiles: CTA rank r of pair q owns M-tile 2*(q % half_m) + r of output-channel tile q / half_m; an odd
  // M-tile count is padded with a dummy tile (fully out-of-bounds loads, no stores)
  const int mt_total = p.seg_mt0[p.nseg];
  const int half_m = (mt_total + 1) >> 1;
  const int pairs = p.tiles_n * half_m;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbarrier_init(&full[s], 1);
      mbarrier_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbarrier_init(&tfull[a], 1);
      mbarrier_init(&tempty[a], 2 * 4 * Cfg::EPI_GROUPS);   // one arrive per epilogue warp of BOTH CTAs (leader's copy is used)
    }
    for (int w = 0; w < 4 * Cfg::EPI_GROUPS; ++w) mbarrier_init(&rbars[w], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)),
                 "n"(Cfg::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  cluster_sync_all();   // barriers of both CTAs initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      pdl_wait();   // inputs are produced by the previous kernel
      int stage = 0;
      uint32_t phase = 0;
      auto prefetch_a = [&](int qq) {     // flat mode (one segment): this CTA's 128 pixels of pair qq, all channels
        const int mtq = 2 * (qq % half_m) + (int)rank;
        if (mtq >= mt_total) return;
        for (int c = 0; c < p.Cin; c += APF_CH) tma_prefetch_4d(&maps.ap[0], c, mtq * TILE_M, 0, 0);
      };
      if (p.a_pf && (int)(blockIdx.x >> 1) < pairs) prefetch_a(blockIdx.x >> 1);
      for (int q = blockIdx.x >> 1; q < pairs; q += gridDim.x >> 1) {
        const int nt = fastdiv(q, p.dh_mul, p.dh_sh);
        const int mt = 2 * (q - nt * half_m) + (int)rank;
        const bool dummy = mt >= mt_total;
        int sg = 0;
        while (!dummy && sg + 1 < p.nseg && mt >= p.seg_mt0[sg + 1]) ++sg;
        const ConvSeg& S = p.seg[sg];
        int tx, ty, n;
        decode_pixel_tile(S, dummy ? 0 : mt - p.seg_mt0[sg], tx, ty, n);
        if (dummy) n = S.N * p.planes;   // image index beyond every plane: TMA zero-fills
        const int x0 = tx * S.TW * p.sx, y0 = ty * S.TH * p.sy;
        if (p.a_pf && q + (int)(gridDim.x >> 1) < pairs) prefetch_a(q + (gridDim.x >> 1));
        // the epilogue will read this tile's residual / ReLU-mask rows a whole main loop from now: pull them into
        // L2 with one TMA prefetch each so those loads see L2 latency instead of HBM latency
        if (S.pf_add) tma_prefetch_4d(&maps.r[sg], nt * BN, tx * S.TW, ty * S.TH, n);
        if (S.pf_mask) tma_prefetch_4d(&maps.m[sg], nt * BN, tx * S.TW, ty * S.TH, n);
        // (plane pair, tap, 64-channel slice): nested counters, no divisions on the producer's critical path
        const int wrow = nt * BN + (int)rank * (BN / 2);
        for (int cmb = 0; cmb < p.ncombo; ++cmb) {
          const int na = dummy ? n : n + p.comb_a[cmb] * S.N, wk0 = p.comb_w[cmb] * p.w_plane;
          for (int tap = 0; tap < p.ntaps; ++tap) {
            const int ax = x0 + p.tap_dx[tap], ay = y0 + p.tap_dy[tap], wk = wk0 + p.tap_wk[tap];
            for (int c0 = 0; c0 < p.Cin; c0 += KCH) {
              mbarrier_wait(&empty[stage], phase ^ 1);
              unsigned char* sa = smem + (size_t)stage * Cfg::STAGE_BYTES;
              // both CTAs' loads complete on the LEADER's barrier; only the leader arms it (bytes of the pair)
              if (rank == 0) mbarrier_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
              tma_load_4d_2sm(sa, &maps.a[sg], &full[stage], c0, ax, ay, na);
              tma_load_2d_2sm(sa + A_TILE_BYTES, &tmB, &full[stage], wk + c0, wrow);
              if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int q = blockIdx.x >> 1; q < pairs; q += gridDim.x >> 1, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbarrier_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbarrier_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = s2u(smem + (size_t)stage * Cfg::STAGE_BYTES);
          const uint64_t da = smem_desc_k128(sa), db = smem_desc_k128(sa + A_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < KCH / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
            umma_bf16_2cta(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), Cfg::IDESC, (kb | k) ? 1u : 0u);
          }
          umma_commit_2cta(&empty[stage]);            // frees the smem slot in BOTH CTAs when the MMAs retire
          if (kb == kblocks - 1) umma_commit_2cta(&tfull[acc]);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue ==========================================
    // 4*EPI_GROUPS warps (EPI_GROUPS per SM sub-partition, hiding each other's ALU / TMEM-load / residual-load
    // latency): warp w reads TMEM lanes [32*(w%4), +32) = 32 tile rows, and BN/EPI_GROUPS of the tile's columns.  Per 32-column
    // chunk: tcgen05.ld (software-pipelined one chunk ahead), per-channel scale/bias (float4 loads),
    // residual / accumulate / ReLU / mask, 16-byte stores along the thread's own output row.
    pdl_wait();   // residual / mask / accumulate reads and the output stores touch memory of earlier kernels
    const int ew = warp & 3, half = (warp - 4) >> 2;              // half = column group of this warp
    unsigned char* ebuf = staging + (size_t)(warp - 4) * EPI_BOX_BYTES;   // TMA-epilogue staging box of this warp
    uint64_t* rbar = rbars + (warp - 4);
    uint32_t rph = 0;
    int it = 0;
    for (int q = blockIdx.x >> 1; q < pairs; q += gridDim.x >> 1, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int nt = fastdiv(q, p.dh_mul, p.dh_sh);
      const int mt = 2 * (q - nt * half_m) + (int)rank;
      const bool dummy = mt >= mt_total;
      int sg = 0;
      while (!dummy && sg + 1 < p.nseg && mt >= p.seg_mt0[sg + 1]) ++sg;
      const ConvSeg& S = p.seg[sg];
      int tx, ty, n;
      decode_pixel_tile(S, dummy ? 0 : mt - p.seg_mt0[sg], tx, ty, n);
      if (dummy) n = S.N * p.planes;   // image index beyond every plane: TMA zero-fills
      if constexpr (FEAT >= 0)
        epilogue_tile_fast<BN, SHORTK, Cfg::EPI_GROUPS, FEAT>(p, maps, S, sg, nt, tx, ty, n, tmem_base + (uint32_t)(acc * BN),
                                                             &tfull[acc], acc_phase, ew, half, lane, ebuf, rbar, rph);
      else
        epilogue_tile<BN, SHORTK, Cfg::EPI_GROUPS, EXT>(p, maps, S, sg, nt, tx, ty, n, !dummy, tmem_base + (uint32_t)(acc * BN),
                                                   &tfull[acc], acc_phase, ew, half, lane, ebuf, rbar, rph);
      if (lane == 0) mbarrier_arrive_leader(&tempty[acc]);
    }
    if (lane == 0) bulk_wait_all();   // staging buffers must outlive the TMA stores that read them
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS));
  }
}

// =============================================================================================
// Weight gradient: dW[co, tap, ci] += sum_{n,y,x} dY[n,y,x,co] * X[n, y*sy+dy_tap, x*sx+dx_tap, ci]
// (autograd's conv backward-weight for every trainable conv of the student).  GEMM with the
// reduction over pixels: both operands are "MN-major" straight out of NHWC memory -- the same 4-D TMA
// boxes {64 ch, TW, TH, 1} as the forward pass land as [128 pixels][128 B] tiles, read by tcgen05 as
// MN-major SWIZZLE_128B (rows = K = pixels; 8-pixel groups 1024 B apart = SBO; 64-channel chunks one
// box = 16 KB apart = LBO).  One work unit = (128 output channels) x (BN input channels) x one tap x
// one slice of the pixel tiles (split-K); the fp32 partial leaves TMEM through vector red.global.add.
struct WgradParams {
  int N, TH, TW, tiles_y, tiles_x;
  int Cout, Cin, ntaps;
  int tap_dx[9], tap_dy[9], tap_wk[9];
  int sx, sy;
  int co_tiles, ci_tiles, splits;
  float* dw;
  int w_k;
  const float* scale;     // optional per-output-channel factor applied to the accumulated gradient
  const float* w_master;  // optional fp32 weights [Cout, w_k]: d_scale[co] += <w[co,:], G[co,:]>
  float* d_scale;
  // split-bf16 mode: dy and x are `planes` stacked bf16 tensors each; every pixel tile is reduced over ncombo
  // (dy plane, x plane) pairs into the same accumulator (planes = ncombo = 1: plain bf16)
  int planes, ncombo;
  int comb_a[6], comb_b[6];
};

// MN-major SWIZZLE_128B descriptor: LBO = 16 KB between 64-element chunks, SBO = 1 KB between 8-row groups
__device__ __forceinline__ uint64_t smem_desc_mn128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)(A_TILE_BYTES >> 4) << 16;   // LBO
  d |= (uint64_t)(1024 >> 4) << 32;           // SBO
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BN> struct WgradCfg {
  static constexpr int A_BYTES = 2 * A_TILE_BYTES;             // 128 output channels
  static constexpr int B_BYTES = (BN / 64) * A_TILE_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 2 : (BN == 128 ? 3 : 4);
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                    ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
};

template <int BN>
__global__ void __launch_bounds__(CONV_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradParams p) {
  using Cfg = WgradCfg<BN>;
  extern __shared__ unsigned char conv_smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)conv_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + (size_t)Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* tfull = bars + 2 * Cfg::STAGES;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ptiles = p.N * p.tiles_y * p.tiles_x;
  const int units = p.co_tiles * p.ci_tiles * p.ntaps * p.splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbarrier_init(&full[s], 1);
      mbarrier_init(&empty[s], 1);
    }
    mbarrier_init(tfull, 1);
    mbarrier_init(tempty, 8);   // one arrive per epilogue warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)),
                 "n"(Cfg::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int split = u % p.splits;
        int rest = u / p.splits;
        const int tap = rest % p.ntaps;
        rest /= p.ntaps;
        const int ci_t = rest % p.ci_tiles, co_t = rest / p.ci_tiles;
        const int pt0 = (int)(((long long)ptiles * split) / p.splits);
        const int pt1 = (int)(((long long)ptiles * (split + 1)) / p.splits);
        for (int pt = pt0; pt < pt1; ++pt) {
          const int tx = pt % p.tiles_x;
          const int ty = (pt / p.tiles_x) % p.tiles_y;
          const int n = pt / (p.tiles_x * p.tiles_y);
          for (int cmb = 0; cmb < p.ncombo; ++cmb) {
            mbarrier_wait(&empty[stage], phase ^ 1);
            unsigned char* sa = smem + (size_t)stage * Cfg::STAGE_BYTES;
            mbarrier_expect_tx(&full[stage], Cfg::STAGE_BYTES);
            const int na = n + p.comb_a[cmb] * p.N, nb = n + p.comb_b[cmb] * p.N;
#pragma unroll
            for (int j = 0; j < 2; ++j)
              tma_load_4d(sa + j * A_TILE_BYTES, &tmA, &full[stage], co_t * TILE_M + 64 * j, tx * p.TW, ty * p.TH, na);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_4d(sa + Cfg::A_BYTES + j * A_TILE_BYTES, &tmB, &full[stage], ci_t * BN + 64 * j,
                          tx * p.TW * p.sx + p.tap_dx[tap], ty * p.TH * p.sy + p.tap_dy[tap], nb);
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
        const int split = u % p.splits;
        const int pt0 = (int)(((long long)ptiles * split) / p.splits);
        const int pt1 = (int)(((long long)ptiles * (split + 1)) / p.splits);
        mbarrier_wait(tempty, (it & 1) ^ 1);
        tc_fence_after();
        for (int pt = pt0; pt < pt1; ++pt) {
          for (int cmb = 0; cmb < p.ncombo; ++cmb) {
            mbarrier_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t sa = s2u(smem + (size_t)stage * Cfg::STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < TILE_M / 16; ++k) {
              // 16 pixels = two 8-row groups = 2 KB further down the tile
              umma_bf16(tmem_base, smem_desc_mn128(sa + k * 2048), smem_desc_mn128(sa + Cfg::A_BYTES + k * 2048),
                        Cfg::IDESC, (pt > pt0 || cmb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty[stage]);
            if (pt == pt1 - 1 && cmb == p.ncombo - 1) umma_commit(tfull);
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
          }
        }
        if (pt1 == pt0) umma_commit(tfull);   // empty slice: nothing accumulated (epilogue skips it)
      }
    }
  } else if (warp >= 4) {
    pdl_wait();
    // 8 warps: lane quadrant ew (TMEM lanes 32*ew..+31 = output channels), column half `half` of the accumulator
    const int ew = warp & 3, half = (warp - 4) >> 2;
    int it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
      const int split = u % p.splits;
      int rest = u / p.splits;
      const int tap = rest % p.ntaps;
      rest /= p.ntaps;
      const int ci_t = rest % p.ci_tiles, co_t = rest / p.ci_tiles;
      const int pt0 = (int)(((long long)ptiles * split) / p.splits);
      const int pt1 = (int)(((long long)ptiles * (split + 1)) / p.splits);
      const int co = co_t * TILE_M + ew * 32 + lane;
      mbarrier_wait(tfull, it & 1);
      tc_fence_after();
      if (pt1 > pt0) {
#pragma unroll 1
        for (int ch = half * (BN / 64); ch < (half + 1) * (BN / 64); ++ch) {
          const int cbase = ci_t * BN + ch * 32;
          if (cbase >= p.Cin) break;
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(ch * 32), r);
          if (co < p.Cout) {
            const size_t woff = (size_t)co * p.w_k + p.tap_wk[tap] + cbase;
            float* o = p.dw + woff;
            const float sc = p.scale ? __ldg(p.scale + co) : 1.f;
            float dot = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (cbase + j + 4 <= p.Cin) {
                const float4 g4 = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                              __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                if (p.d_scale) {
                  const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.w_master + woff + j));
                  dot += g4.x * w4.x + g4.y * w4.y + g4.z * w4.z + g4.w * w4.w;
                }
                atomicAdd(reinterpret_cast<float4*>(o + j), make_float4(g4.x * sc, g4.y * sc, g4.z * sc, g4.w * sc));
              } else {
                for (int e = 0; e < 4; ++e)
                  if (cbase + j + e < p.Cin) {
                    const float gv = __uint_as_float(r[j + e]);
                    if (p.d_scale) dot += gv * __ldg(p.w_master + woff + j + e);
                    atomicAdd(o + j + e, gv * sc);
                  }
              }
            }
            if (p.d_scale) atomicAdd(p.d_scale + co, dot);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbarrier_arrive(tempty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS));
  }
}

// 2-CTA weight gradient (cta_group::2): 256 output channels x 256 input channels per cluster, every CTA stages its
// own 128 + 128 channels of the pixel tile.  The 1-CTA 128 x 256 tile needs (128 + 256) * 128 * 2 B of operands per
// 8.4 MFLOP = 131 GB/s per SM at tensor peak -- more than an SM can pull from L2; the pair tile needs 87 GB/s.
struct Wgrad2Cfg {
  static constexpr int BN = 256;
  static constexpr int A_BYTES = 2 * A_TILE_BYTES;     // this CTA's 128 output channels
  static constexpr int B_BYTES = 2 * A_TILE_BYTES;     // this CTA's 128 input channels
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 3;
  static constexpr int TMEM_COLS = 256;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                    ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CONV_THREADS, 1)
wgrad2_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradParams p) {
  using Cfg = Wgrad2Cfg;
  constexpr int BN = Wgrad2Cfg::BN;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  extern __shared__ unsigned char conv_smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)conv_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + (size_t)Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* tfull = bars + 2 * Cfg::STAGES;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ptiles = p.N * p.tiles_y * p.tiles_x;
  const int units = p.co_tiles * p.ci_tiles * p.ntaps * p.splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbarrier_init(&full[s], 1);
      mbarrier_init(&empty[s], 1);
    }
    mbarrier_init(tfull, 1);
    mbarrier_init(tempty, 2 * 8);   // epilogue warps of both CTAs arrive on the leader's copy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)),
                 "n"(Cfg::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x >> 1; u < units; u += gridDim.x >> 1) {
        const int split = u % p.splits;
        int rest = u / p.splits;
        const int tap = rest % p.ntaps;
        rest /= p.ntaps;
        const int ci_t = rest % p.ci_tiles, co_t = rest / p.ci_tiles;
        const int pt0 = (int)(((long long)ptiles * split) / p.splits);
        const int pt1 = (int)(((long long)ptiles * (split + 1)) / p.splits);
        for (int pt = pt0; pt < pt1; ++pt) {
          const int tx = pt % p.tiles_x;
          const int ty = (pt / p.tiles_x) % p.tiles_y;
          const int n = pt / (p.tiles_x * p.tiles_y);
          for (int cmb = 0; cmb < p.ncombo; ++cmb) {
            mbarrier_wait(&empty[stage], phase ^ 1);
            unsigned char* sa = smem + (size_t)stage * Cfg::STAGE_BYTES;
            // this CTA's 128 of the pair's 256 output channels and 128 of its 256 input channels; all four boxes of
            // both CTAs complete on the leader's barrier
            if (rank == 0) mbarrier_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
            const int na = n + p.comb_a[cmb] * p.N, nb = n + p.comb_b[cmb] * p.N;
#pragma unroll
            for (int j = 0; j < 2; ++j)
              tma_load_4d_2sm(sa + j * A_TILE_BYTES, &tmA, &full[stage], co_t * 2 * TILE_M + (int)rank * TILE_M + 64 * j,
                              tx * p.TW, ty * p.TH, na);
#pragma unroll
            for (int j = 0; j < 2; ++j)
              tma_load_4d_2sm(sa + Cfg::A_BYTES + j * A_TILE_BYTES, &tmB, &full[stage], ci_t * BN + (int)rank * (BN / 2) + 64 * j,
                              tx * p.TW * p.sx + p.tap_dx[tap], ty * p.TH * p.sy + p.tap_dy[tap], nb);
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = blockIdx.x >> 1; u < units; u += gridDim.x >> 1, ++it) {
        const int split = u % p.splits;
        const int pt0 = (int)(((long long)ptiles * split) / p.splits);
        const int pt1 = (int)(((long long)ptiles * (split + 1)) / p.splits);
        mbarrier_wait(tempty, (it & 1) ^ 1);
        tc_fence_after();
        for (int pt = pt0; pt < pt1; ++pt) {
          for (int cmb = 0; cmb < p.ncombo; ++cmb) {
            mbarrier_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t sa = s2u(smem + (size_t)stage * Cfg::STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < TILE_M / 16; ++k) {
              // 16 pixels = two 8-row groups = 2 KB further down the tile
              umma_bf16_2cta(tmem_base, smem_desc_mn128(sa + k * 2048), smem_desc_mn128(sa + Cfg::A_BYTES + k * 2048),
                        Cfg::IDESC, (pt > pt0 || cmb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit_2cta(&empty[stage]);
            if (pt == pt1 - 1 && cmb == p.ncombo - 1) umma_commit_2cta(tfull);
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
          }
        }
        if (pt1 == pt0) umma_commit_2cta(tfull);   // empty slice: nothing accumulated (epilogue skips it)
      }
    }
  } else if (warp >= 4) {
    pdl_wait();
    // 8 warps: lane quadrant ew (TMEM lanes 32*ew..+31 = output channels), column half `half` of the accumulator
    const int ew = warp & 3, half = (warp - 4) >> 2;
    int it = 0;
    for (int u = blockIdx.x >> 1; u < units; u += gridDim.x >> 1, ++it) {
      const int split = u % p.splits;
      int rest = u / p.splits;
      const int tap = rest % p.ntaps;
      rest /= p.ntaps;
      const int ci_t = rest % p.ci_tiles, co_t = rest / p.ci_tiles;
      const int pt0 = (int)(((long long)ptiles * split) / p.splits);
      const int pt1 = (int)(((long long)ptiles * (split + 1)) / p.splits);
      const int co = co_t * 2 * TILE_M + (int)rank * TILE_M + ew * 32 + lane;
      mbarrier_wait(tfull, it & 1);
      tc_fence_after();
      if (pt1 > pt0) {
#pragma unroll 1
        for (int ch = half * (BN / 64); ch < (half + 1) * (BN / 64); ++ch) {
          const int cbase = ci_t * BN + ch * 32;
          if (cbase >= p.Cin) break;
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(ch * 32), r);
          if (co < p.Cout) {
            const size_t woff = (size_t)co * p.w_k + p.tap_wk[tap] + cbase;
            float* o = p.dw + woff;
            const float sc = p.scale ? __ldg(p.scale + co) : 1.f;
            float dot = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (cbase + j + 4 <= p.Cin) {
                const float4 g4 = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                              __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                if (p.d_scale) {
                  const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.w_master + woff + j));
                  dot += g4.x * w4.x + g4.y * w4.y + g4.z * w4.z + g4.w * w4.w;
                }
                atomicAdd(reinterpret_cast<float4*>(o + j), make_float4(g4.x * sc, g4.y * sc, g4.z * sc, g4.w * sc));
              } else {
                for (int e = 0; e < 4; ++e)
                  if (cbase + j + e < p.Cin) {
                    const float gv = __uint_as_float(r[j + e]);
                    if (p.d_scale) dot += gv * __ldg(p.w_master + woff + j + e);
                    atomicAdd(o + j + e, gv * sc);
                  }
              }
            }
            if (p.d_scale) atomicAdd(p.d_scale + co, dot);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbarrier_arrive_leader(tempty);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

static int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box, const cuuint32_t* estr, const char* what,
                      CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  PFN_encodeTiled enc = get_encode();
  LDB_REQUIRE(enc, LDB_ENOSUP, "cuTensorMapEncodeTiled not available from the driver");
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LDB_REQUIRE(r == CUDA_SUCCESS, LDB_EINVAL, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
  return 0;
}

static int g_conv_sms = 0;

// optional programmatic dependent launch (LDB200_PDL=1): measured neutral on the LD step (299 vs 300 img/s: the
// 197 KB-smem CTAs cannot co-reside with their successors), so plain launches are the default
template <typename K, typename... Args>
static cudaError_t launch_pdl(K kernel, int grid, int block, size_t smem, cudaStream_t stream, const Args&... args) {
  static const bool use_pdl = getenv("LDB200_PDL") != nullptr;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <int BN, bool SHORTK, int EXT, int FEAT>
static int launch_conv_ext(const ConvMaps& maps, const CUtensorMap& tmB, const ConvParams& p, cudaStream_t stream) {
  using Cfg = ConvCfg<BN, SHORTK>;
  static bool attr = false;
  if (!attr) {
    LDB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, SHORTK, EXT, FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    attr = true;
  }
  const int total = p.seg_tile0[p.nseg];
  const int slots = g_conv_sms * Cfg::MIN_BLOCKS;
  ConvParams q = p;
  // enough pixel tiles to fill the GPU: keep a pixel tile's output-channel tiles on one CTA (A re-read from L2)
  // (opt-in, LDB200_NT_INNER=1: measured neutral-to-slower on the LD shapes, DRAM reads unchanged)
  static const bool nt_inner_env = getenv("LDB200_NT_INNER") != nullptr;
  q.nt_inner = (nt_inner_env && p.tiles_n > 1 && total / p.tiles_n >= slots) ? 1 : 0;
  const int grid = total < slots ? total : slots;
  LDB_CUDA(launch_pdl(conv_tc_kernel<BN, SHORTK, EXT, FEAT>, grid, Cfg::THREADS, Cfg::SMEM, stream, maps, tmB, q));
  return 0;
}

// ext: epilogue feature set of the generic epilogue (0 plain, 1 GroupNorm statistics, 3 statistics + fp32 extras of the
// split mode); feat: compile-time feature set of the lean epilogue (EF_* bits), -1 = generic
template <int BN, bool SHORTK>
static int launch_conv(const ConvMaps& maps, const CUtensorMap& tmB, const ConvParams& p, cudaStream_t stream, int ext,
                       int feat) {
  if constexpr (BN == 128) {
    if (ext == 0) {
      switch (feat) {
        case 0: return launch_conv_ext<BN, SHORTK, 0, 0>(maps, tmB, p, stream);
        case EF_BIAS: return launch_conv_ext<BN, SHORTK, 0, EF_BIAS>(maps, tmB, p, stream);
        case EF_AFFINE: return launch_conv_ext<BN, SHORTK, 0, EF_AFFINE>(maps, tmB, p, stream);
        case EF_AFFINE | EF_RELU: return launch_conv_ext<BN, SHORTK, 0, EF_AFFINE | EF_RELU>(maps, tmB, p, stream);
        case EF_AFFINE | EF_RELU | EF_ADD: return launch_conv_ext<BN, SHORTK, 0, EF_AFFINE | EF_RELU | EF_ADD>(maps, tmB, p, stream);
        case EF_MASK: return launch_conv_ext<BN, SHORTK, 0, EF_MASK>(maps, tmB, p, stream);
        case EF_MASK | EF_ADD: return launch_conv_ext<BN, SHORTK, 0, EF_MASK | EF_ADD>(maps, tmB, p, stream);
        default: break;
      }
    }
  }
  if (ext == 0) return launch_conv_ext<BN, SHORTK, 0, -1>(maps, tmB, p, stream);
  if (ext == 1) return launch_conv_ext<BN, SHORTK, 1, -1>(maps, tmB, p, stream);
  return launch_conv_ext<BN, SHORTK, 3, -1>(maps, tmB, p, stream);
}

// 2-CTA kernel (default for Cout % 256 == 0 and more than 4 k-blocks; LDB200_NO_CONV2=1 falls back to the 1-CTA kernel): one cluster of two CTAs per pair-tile, grid = resident clusters
template <int EXT, int FEAT>
static int launch_conv2_ext(const ConvMaps& maps, const CUtensorMap& tmB, const ConvParams& p, cudaStream_t stream) {
  static int max_clusters = 0;
  if (!max_clusters) {
    LDB_CUDA(cudaFuncSetAttribute(conv2_tc_kernel<EXT, FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Conv2Cfg::SMEM));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(g_conv_sms / 2 * 2);
    cfg.blockDim = dim3(Conv2Cfg::THREADS);
    cfg.dynamicSmemBytes = Conv2Cfg::SMEM;
    int n = 0;
    LDB_CUDA(cudaOccupancyMaxActiveClusters(&n, conv2_tc_kernel<EXT, FEAT>, &cfg));
    LDB_REQUIRE(n >= 1, LDB_ENOSUP, "conv2d_tc: no resident 2-CTA cluster");
    max_clusters = n < g_conv_sms / 2 ? n : g_conv_sms / 2;
  }
  const int mt_total = p.seg_tile0[p.nseg] / p.tiles_n;
  const int pairs = p.tiles_n * ((mt_total + 1) / 2);
  const int clusters = pairs < max_clusters ? pairs : max_clusters;
  conv2_tc_kernel<EXT, FEAT><<<2 * clusters, Conv2Cfg::THREADS, Conv2Cfg::SMEM, stream>>>(maps, tmB, p);
  LDB_CUDA(cudaGetLastError());
  return 0;
}
static int launch_conv2(const ConvMaps& maps, const CUtensorMap& tmB, const ConvParams& p, cudaStream_t stream, int ext,
                        int feat) {
  if (ext == 0) {
    switch (feat) {
      case 0: return launch_conv2_ext<0, 0>(maps, tmB, p, stream);
      case EF_BIAS: return launch_conv2_ext<0, EF_BIAS>(maps, tmB, p, stream);
      case EF_AFFINE: return launch_conv2_ext<0, EF_AFFINE>(maps, tmB, p, stream);
      case EF_AFFINE | EF_RELU: return launch_conv2_ext<0, EF_AFFINE | EF_RELU>(maps, tmB, p, stream);
      case EF_AFFINE | EF_RELU | EF_ADD: return launch_conv2_ext<0, EF_AFFINE | EF_RELU | EF_ADD>(maps, tmB, p, stream);
      case EF_MASK: return launch_conv2_ext<0, EF_MASK>(maps, tmB, p, stream);
      case EF_MASK | EF_ADD: return launch_conv2_ext<0, EF_MASK | EF_ADD>(maps, tmB, p, stream);
      default: break;
    }
    return launch_conv2_ext<0, -1>(maps, tmB, p, stream);
  }
  if (ext == 1) return launch_conv2_ext<1, -1>(maps, tmB, p, stream);
  return launch_conv2_ext<3, -1>(maps, tmB, p, stream);
}

// constants of fastdiv(n, mul, sh) = n / d for 0 <= n < 2^31 (round-up method: mul = floor((2^s - d) 2^32 / d) + 1,
// s = ceil(log2 d); q = (umulhi(n, mul) + n) >> s)
static void make_div(int d, uint32_t* mul, uint32_t* sh) {
  uint32_t s = 0;
  while ((1u << s) < (uint32_t)d) ++s;
  *sh = s;
  *mul = (uint32_t)(((((uint64_t)1 << s) - (uint64_t)d) << 32) / (uint64_t)d + 1);
}

// geometry + tensor maps of one segment
static int build_segment(const ldb_conv_desc* d, int BN, const ConvParams& common, ConvSeg* g, CUtensorMap* tmA,
                         CUtensorMap* tmR, CUtensorMap* tmM, CUtensorMap* tmO, CUtensorMap* tmRL, CUtensorMap* tmML,
                         CUtensorMap* tmAP) {
  LDB_REQUIRE(d->in && d->out, LDB_EINVAL, "conv2d_tc: null pointer");
  LDB_REQUIRE(d->in_ld >= d->Cin && d->in_ld % 8 == 0, LDB_EINVAL, "conv2d_tc: in_ld must be >= Cin and a multiple of 8");
  LDB_REQUIRE(d->N >= 1 && d->HO >= 1 && d->WO >= 1, LDB_EINVAL, "conv2d_tc: shape");
  // the input feeds TMA (16-byte aligned); the output only needs element alignment -- rows that are not 16-byte
  // aligned (e.g. fp32 rows of 81 channels, the LDv2 class logits) take the scalar epilogue path (vec_ok = 0)
  LDB_REQUIRE(((uintptr_t)d->in & 15) == 0 && ((uintptr_t)d->out & (d->out_f32 ? 3 : 1)) == 0, LDB_EINVAL,
              "conv2d_tc: input must be 16-byte aligned, output element-aligned");
  memset(g, 0, sizeof(*g));
  g->N = d->N; g->HO = d->HO; g->WO = d->WO;
  g->OH = d->OH ? d->OH : d->HO; g->OW = d->OW ? d->OW : d->WO;
  g->oys = d->oys ? d->oys : 1; g->oxs = d->oxs ? d->oxs : 1; g->oyo = d->oyo; g->oxo = d->oxo;
  g->out = d->out;
  g->add_src = d->add_src;
  g->add_H = d->add_H ? d->add_H : g->OH; g->add_W = d->add_W ? d->add_W : g->OW;
  g->mask_src = d->mask_src;
  g->gn_stats = d->gn_stats;
  g->pre_src = d->pre_src;
  LDB_REQUIRE(!d->pre_src || (d->out_f32 && ((uintptr_t)d->pre_src & 15) == 0), LDB_EINVAL,
              "conv2d_tc: pre_src needs an fp32 output and a 16-byte aligned pointer");
  LDB_REQUIRE(!d->gn_stats || (!d->flat && d->Cout % 32 == 0 && common.vec_ok), LDB_EINVAL,
              "conv2d_tc: gn_stats needs a non-flat conv, Cout %% 32 == 0 and 16-byte aligned rows");
  if (g->add_src) LDB_REQUIRE(((uintptr_t)g->add_src & 15) == 0, LDB_EINVAL, "conv2d_tc: add_src alignment");
  if (d->flat) {
    LDB_REQUIRE(d->ntaps == 1 && d->sx == 1 && d->sy == 1, LDB_EINVAL, "conv2d_tc: flat mode needs a 1x1 stride-1 conv");
    const long long M = (long long)d->N * d->HO * d->WO;
    LDB_REQUIRE(!g->add_src || (g->add_H == g->OH && g->add_W == g->OW), LDB_ENOSUP,
                "conv2d_tc: flat mode cannot resample add_src");
    g->N = 1; g->HO = 1; g->WO = (int)M; g->TH = 1; g->TW = TILE_M;
    g->OH = 1; g->OW = (int)M; g->oys = g->oxs = 1; g->oyo = g->oxo = 0;
    g->add_H = 1; g->add_W = (int)M;
    // split mode: the planes are the 4th coordinate (S.N = 1, so plane p is "image" p)
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)M, 1, (cuuint64_t)common.planes};
    cuuint64_t strides[3] = {(cuuint64_t)d->in_ld * 2, (cuuint64_t)d->in_ld * 2 * (cuuint64_t)M,
                             (cuuint64_t)d->in_ld * 2 * (cuuint64_t)M};
    cuuint32_t box[4] = {KCH, TILE_M, 1, 1}, es[4] = {1, 1, 1, 1};
    int e = encode_map(tmA, d->in, 4, dims, strides, box, es, "A(flat)");
    if (e) return e;
    if (tmAP) {
      cuuint32_t pbox[4] = {(cuuint32_t)(d->Cin < APF_CH ? d->Cin : APF_CH), TILE_M, 1, 1};
      e = encode_map(tmAP, d->in, 4, dims, strides, pbox, es, "A prefetch (flat)", CU_TENSOR_MAP_SWIZZLE_NONE);
      if (e) return e;
    }
  } else {
    int best_th = 8, best_tiles = 1 << 30;
    for (int th = 1; th <= TILE_M; th *= 2) {
      const int tw = TILE_M / th;
      if (tw * d->sx > 256 || th * d->sy > 256) continue;
      const int tiles = ((d->HO + th - 1) / th) * ((d->WO + tw - 1) / tw);
      if (tiles < best_tiles || (tiles == best_tiles && abs(th - 8) < abs(best_th - 8))) {
        best_tiles = tiles;
        best_th = th;
      }
    }
    g->TH = best_th; g->TW = TILE_M / best_th;
    // split mode: `in` holds `planes` stacked [N,H,W,in_ld] tensors = images [0, planes*N)
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N * common.planes};
    cuuint64_t strides[3] = {(cuuint64_t)d->in_ld * 2, (cuuint64_t)d->in_ld * 2 * d->W,
                             (cuuint64_t)d->in_ld * 2 * d->W * d->H};
    cuuint32_t box[4] = {KCH, (cuuint32_t)(g->TW * d->sx), (cuuint32_t)(g->TH * d->sy), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)d->sx, (cuuint32_t)d->sy, 1};
    int e = encode_map(tmA, d->in, 4, dims, strides, box, es, "A");
    if (e) return e;
  }
  g->tiles_y = (g->HO + g->TH - 1) / g->TH;
  g->tiles_x = (g->WO + g->TW - 1) / g->TW;
  make_div(g->tiles_x, &g->dx_mul, &g->dx_sh);
  make_div(g->tiles_y, &g->dy_mul, &g->dy_sh);
  g->tw_shift = 0;
  while ((1 << g->tw_shift) < g->TW) ++g->tw_shift;
  // optional L2 prefetch maps for the residual / mask tensors (identity pixel mapping, same extent as the output)
  *tmR = *tmA;
  *tmM = *tmA;
  const bool ident = g->oys == 1 && g->oxs == 1 && g->oyo == 0 && g->oxo == 0;
  auto aux_map = [&](CUtensorMap* m, const void* base, int ld, const char* what) -> int {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)g->OW, (cuuint64_t)g->OH, (cuuint64_t)g->N};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * g->OW, (cuuint64_t)ld * 2 * g->OW * g->OH};
    cuuint32_t box[4] = {(cuuint32_t)(d->Cout < BN ? (d->Cout + 7) / 8 * 8 : BN), (cuuint32_t)g->TW, (cuuint32_t)g->TH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    return encode_map(m, base, 4, dims, strides, box, es, what, CU_TENSOR_MAP_SWIZZLE_NONE);
  };
  static const bool l2_prefetch = getenv("LDB200_NO_L2_PREFETCH") == nullptr;
  if (l2_prefetch && g->add_src && !common.add_f32 && ident && g->add_H == g->OH && g->add_W == g->OW && common.add_ld % 8 == 0 && !d->out_f32) {
    if (aux_map(tmR, g->add_src, common.add_ld, "residual prefetch") == 0) g->pf_add = 1;
  }
  if (l2_prefetch && g->mask_src && !common.mask_f32 && ident && common.mask_ld % 8 == 0) {
    if (aux_map(tmM, g->mask_src, common.mask_ld, "mask prefetch") == 0) g->pf_mask = 1;
  }
  // TMA epilogue: one box = the 32 pixels of an epilogue warp (32 x 1 of a wide tile, TW x 32/TW of a narrow one)
  // x 64 channels, 128-byte swizzled in shared memory
  *tmO = *tmA; *tmRL = *tmA; *tmML = *tmA;
  if (common.tma_epi) {
    const cuuint32_t bw = g->TW < EPI_BOX_PIX ? g->TW : EPI_BOX_PIX, bh = EPI_BOX_PIX / bw;
    auto box_map = [&](CUtensorMap* m, const void* base, int ld, const char* what) -> int {
      cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)g->OW, (cuuint64_t)g->OH, (cuuint64_t)g->N};
      cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * g->OW, (cuuint64_t)ld * 2 * g->OW * g->OH};
      cuuint32_t box[4] = {64, bw, bh, 1};
      cuuint32_t es[4] = {1, 1, 1, 1};
      return encode_map(m, base, 4, dims, strides, box, es, what);
    };
    int e = box_map(tmO, g->out, common.ldc, "output store");
    if (e) return e;
    if (g->add_src && g->add_H == g->OH && g->add_W == g->OW) {
      e = box_map(tmRL, g->add_src, common.add_ld, "residual load");
      if (e) return e;
      g->tl_add = 1;
    }
    if (g->mask_src) {
      e = box_map(tmML, g->mask_src, common.mask_ld, "mask load");
      if (e) return e;
      g->tl_mask = 1;
    }
    // the staging-box loads are issued before the accumulator wait already; an additional producer-side L2 prefetch
    // of the same tile only duplicated the DRAM reads (ncu, layer-1 expand conv: 975 MB read vs 688 MB without)
    if (g->tl_add) g->pf_add = 0;
    if (g->tl_mask) g->pf_mask = 0;
  }
  return 0;
}

}  // namespace ldb

using namespace ldb;

// `count` convolutions that share weights, taps, stride and epilogue configuration (the five FPN levels of a
// level-shared GFL head conv, gfl_head.py:145-183) as ONE launch over the concatenated tile list.
extern "C" int ldb200_conv2d_tc_multi(const ldb_conv_desc* descs, int count, void* stream_) {
  LDB_REQUIRE(descs && count >= 1 && count <= CONV_MAX_SEG, LDB_EINVAL, "conv2d_tc: 1..%d segments", CONV_MAX_SEG);
  const ldb_conv_desc* d = &descs[0];
  LDB_REQUIRE(d->weight, LDB_EINVAL, "conv2d_tc: null weight");
  LDB_REQUIRE(d->Cin > 0 && d->Cout >= 1, LDB_EINVAL, "conv2d_tc: channels");
  LDB_REQUIRE(d->ntaps >= 1 && d->ntaps <= 9, LDB_EINVAL, "conv2d_tc: ntaps");
  LDB_REQUIRE(d->sx >= 1 && d->sy >= 1 && d->sx <= 8 && d->sy <= 8, LDB_EINVAL, "conv2d_tc: stride");
  LDB_REQUIRE(d->out_ld >= d->Cout, LDB_EINVAL, "conv2d_tc: out_ld < Cout");
  LDB_REQUIRE(((uintptr_t)d->weight & 15) == 0, LDB_EINVAL, "conv2d_tc: weight must be 16-byte aligned");
  LDB_REQUIRE(d->w_k % 8 == 0, LDB_EINVAL, "conv2d_tc: weight row length must be a multiple of 8");
  for (int i = 1; i < count; ++i) {
    const ldb_conv_desc* e = &descs[i];
    LDB_REQUIRE(e->weight == d->weight && e->Cin == d->Cin && e->Cout == d->Cout && e->w_k == d->w_k &&
                e->ntaps == d->ntaps && e->sx == d->sx && e->sy == d->sy && e->out_ld == d->out_ld &&
                e->out_f32 == d->out_f32 && e->scale == d->scale && e->bias == d->bias && e->relu == d->relu &&
                e->accumulate == d->accumulate && e->flat == d->flat && e->in_ld == d->in_ld &&
                (e->add_src != nullptr) == (d->add_src != nullptr) && e->add_ld == d->add_ld &&
                (e->mask_src != nullptr) == (d->mask_src != nullptr) && e->mask_ld == d->mask_ld &&
                (e->gn_stats != nullptr) == (d->gn_stats != nullptr) && (e->pre_src != nullptr) == (d->pre_src != nullptr) &&
                e->split_pass == d->split_pass && e->split == d->split && e->w_plane == d->w_plane && e->add_f32 == d->add_f32 && e->mask_f32 == d->mask_f32 &&
                memcmp(e->tap_dx, d->tap_dx, sizeof(d->tap_dx)) == 0 && memcmp(e->tap_dy, d->tap_dy, sizeof(d->tap_dy)) == 0 &&
                memcmp(e->tap_wk, d->tap_wk, sizeof(d->tap_wk)) == 0,
                LDB_EINVAL, "conv2d_tc_multi: segment %d does not share weights / taps / epilogue with segment 0", i);
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!g_conv_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_conv_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  ConvParams p;
  memset(&p, 0, sizeof(p));
  // channels beyond Cin inside the last 64-wide k-block are TMA out-of-bounds reads (zeros); the weight rows must
  // hold finite values there (tap_wk spacing >= round_up(Cin, 64))
  p.Cout = d->Cout; p.Cin = (d->Cin + KCH - 1) / KCH * KCH; p.ntaps = d->ntaps;
  for (int t = 0; t < d->ntaps; ++t) {
    p.tap_dx[t] = d->tap_dx[t]; p.tap_dy[t] = d->tap_dy[t]; p.tap_wk[t] = d->tap_wk[t];
  }
  p.sx = d->sx; p.sy = d->sy;
  p.ldc = d->out_ld; p.out_f32 = d->out_f32;
  p.scale = d->scale; p.bias = d->bias;
  p.add_ld = d->add_ld ? d->add_ld : d->Cout;
  p.mask_ld = d->mask_ld ? d->mask_ld : d->out_ld;
  p.relu = d->relu; p.accumulate = d->accumulate;
  p.add_f32 = d->add_f32; p.mask_f32 = d->mask_f32;
  // split-bf16 mode (fp32-accurate): x = sum of `split` bf16 planes, W likewise; keep every plane product whose
  // magnitude is >= 2^-24 of the leading one, smallest terms first
  LDB_REQUIRE(d->split == 0 || d->split == 2 || d->split == 3, LDB_EINVAL, "conv2d_tc: split must be 0, 2 or 3");
  p.planes = d->split ? d->split : 1;
  p.w_plane = d->split ? d->w_plane : 0;
  LDB_REQUIRE(!d->split || (d->w_plane % 8 == 0 && d->w_plane * p.planes <= d->w_k), LDB_EINVAL,
              "conv2d_tc: split weights need w_plane %% 8 == 0 and planes * w_plane <= w_k");
  // split_pass: 0 = every plane product in one accumulation chain; 1 = the leading product (plane 0 x plane 0) only;
  // 2 = the correction products only.  Two launches (1, then 2 with pre_src = the result of 1) keep the long chain of
  // the large terms (whose round-off dominates: the tensor core accumulates with truncation) 6x shorter.
  LDB_REQUIRE(d->split_pass >= 0 && d->split_pass <= 2 && (d->split || !d->split_pass), LDB_EINVAL, "conv2d_tc: split_pass");
  if (d->split == 3) {
    static const int ca[6] = {2, 0, 1, 1, 0, 0}, cw[6] = {0, 2, 1, 0, 1, 0};
    const int lo = d->split_pass == 1 ? 5 : 0, hi = d->split_pass == 2 ? 5 : 6;
    p.ncombo = hi - lo;
    for (int i = lo; i < hi; ++i) { p.comb_a[i - lo] = ca[i]; p.comb_w[i - lo] = cw[i]; }
  } else if (d->split == 2) {
    static const int ca[3] = {1, 0, 0}, cw[3] = {0, 1, 0};
    const int lo = d->split_pass == 1 ? 2 : 0, hi = d->split_pass == 2 ? 2 : 3;
    p.ncombo = hi - lo;
    for (int i = lo; i < hi; ++i) { p.comb_a[i - lo] = ca[i]; p.comb_w[i - lo] = cw[i]; }
  } else {
    p.ncombo = 1;
  }
  // the fast epilogue path moves 16 bytes per access: bf16 rows need ld % 8 == 0, fp32 rows ld % 4 == 0
  p.vec_ok = (d->out_ld % (d->out_f32 ? 4 : 8) == 0) && (!d->mask_src || p.mask_ld % (d->mask_f32 ? 4 : 8) == 0) &&
             (!d->add_src || p.add_ld % (d->add_f32 ? 4 : 8) == 0) && (!d->scale || ((uintptr_t)d->scale & 15) == 0) &&
             (!d->bias || ((uintptr_t)d->bias & 15) == 0);
  for (int i = 0; i < count; ++i)
    if ((uintptr_t)descs[i].out & 15) p.vec_ok = 0;
  const bool shortk = p.ncombo * d->ntaps * (p.Cin / KCH) <= 4 && getenv("LDB200_NO_SHORTK") == nullptr;
  const int BN = shortk ? (d->Cout > 64 ? 128 : 64) : (d->Cout > 128 ? 256 : (d->Cout > 64 ? 128 : 64));
  p.tiles_n = (d->Cout + BN - 1) / BN;
  p.nseg = count;
  // TMA-staged epilogue: bf16 output with the identity pixel mapping (everything but the stride-2 dgrad parity
  // classes, fp32 head outputs and accumulating launches); LDB200_NO_TMA_EPI=1 forces the direct stores
  static const bool tma_epi_env = getenv("LDB200_NO_TMA_EPI") == nullptr;
  p.tma_epi = tma_epi_env && BN >= 128 && !d->out_f32 && !d->accumulate && p.vec_ok && d->Cout % 64 == 0 &&
              !d->add_f32 && !d->mask_f32;
  for (int i = 0; i < count && p.tma_epi; ++i) {
    const ldb_conv_desc* e = &descs[i];
    const bool ident = e->oys <= 1 && e->oxs <= 1 && e->oyo == 0 && e->oxo == 0 && (e->OH == 0 || e->OH == e->HO) &&
                       (e->OW == 0 || e->OW == e->WO);
    if (!ident || ((uintptr_t)e->out & 15) || (e->mask_src && ((uintptr_t)e->mask_src & 15))) p.tma_epi = 0;
  }
  // 256-wide tiles push TWO 64-column boxes per warp through the one staging buffer; measured on the LD shapes this
  // only pays when a residual or mask box comes in through it too (plain 1x1 convs: 768 vs 814 TFLOP/s)
  // (r02 re-measurement, tools/hot_ops.py convs with LDB200_TMA_EPI_ALL=1: plain 256-wide launches are faster through
  //  the boxes too -- 1x1 1024->256 44.8 vs 52.4 us, 2048->512 34.3 vs 36.4, 3x3 256->256 67.9 vs 69.8 -- so only the
  //  GroupNorm-statistics launches of the head towers keep the old rule; LDB200_TMA_EPI_OLD=1 restores it everywhere)
  static const bool tma_epi_old = getenv("LDB200_TMA_EPI_OLD") != nullptr;
  static const bool tma_epi_all = getenv("LDB200_TMA_EPI_ALL") != nullptr;
  if (p.tma_epi && BN == 256 && !tma_epi_all && (tma_epi_old || d->gn_stats)) {
    const bool same_size_add = d->add_src && (d->add_H == 0 || d->add_H == (d->OH ? d->OH : d->HO)) &&
                               (d->add_W == 0 || d->add_W == (d->OW ? d->OW : d->WO));
    if (!same_size_add && !d->mask_src) p.tma_epi = 0;
  }
  // A-operand L2 prefetch: flat 1x1 convs (single segment) whose input rows are at least 512 bytes wide
  // (opt-in, LDB200_A_PREFETCH=1: measured 7-10 % SLOWER on the 1x1 convs of the LD step -- 1024->256 64.4 vs 58.8 us,
  //  512->128 72.8 vs 66.9 us -- so the wide-row DRAM pattern is not what holds those launches back)
  static const bool a_pf_env = getenv("LDB200_A_PREFETCH") != nullptr;
  p.a_pf = (a_pf_env && d->flat && count == 1 && !d->split && d->Cin >= 256 && d->in_ld % 8 == 0) ? 1 : 0;
  ConvMaps maps;
  int tile0 = 0;
  for (int i = 0; i < count; ++i) {
    int e = build_segment(&descs[i], BN, p, &p.seg[i], &maps.a[i], &maps.r[i], &maps.m[i], &maps.o[i], &maps.rl[i],
                          &maps.ml[i], p.a_pf ? &maps.ap[i] : nullptr);
    if (e) return e;
    if (!p.a_pf) maps.ap[i] = maps.a[i];
    p.seg_tile0[i] = tile0;
    tile0 += p.seg[i].N * p.seg[i].tiles_y * p.seg[i].tiles_x * p.tiles_n;
  }
  for (int i = count; i <= CONV_MAX_SEG; ++i) p.seg_tile0[i] = tile0;
  for (int i = 0; i <= CONV_MAX_SEG; ++i) p.seg_mt0[i] = p.seg_tile0[i] / p.tiles_n;
  make_div(p.tiles_n, &p.dn_mul, &p.dn_sh);
  make_div((p.seg_mt0[count] + 1) / 2, &p.dh_mul, &p.dh_sh);
  for (int i = count; i < CONV_MAX_SEG; ++i) {
    maps.a[i] = maps.r[i] = maps.m[i] = maps.o[i] = maps.rl[i] = maps.ml[i] = maps.ap[i] = maps.a[0];
  }
  static const bool conv2_env = getenv("LDB200_NO_CONV2") == nullptr;   // 2-CTA kernel is the default for 256-wide tiles
  const bool use2 = conv2_env && !shortk && BN == 256 && d->Cout % 256 == 0;
  CUtensorMap tmB;
  {
    cuuint64_t dims[2] = {(cuuint64_t)d->w_k, (cuuint64_t)d->Cout};
    cuuint64_t strides[1] = {(cuuint64_t)d->w_k * 2};
    cuuint32_t box[2] = {KCH, (cuuint32_t)(use2 ? BN / 2 : BN)}, es[2] = {1, 1};
    int e = encode_map(&tmB, d->weight, 2, dims, strides, box, es, "B");
    if (e) return e;
  }
  // epilogue feature set of this launch (compiled out of the plain bf16 path)
  int ext = 0;
  for (int i = 0; i < count; ++i) {
    if (descs[i].gn_stats) ext |= 1;
    if (descs[i].pre_src) ext |= 2;
  }
  if (d->add_f32 || d->mask_f32) ext |= 2;
  // lean epilogue: TMA-staged bf16 output of complete 64-column boxes, residual / mask through the staging boxes, scale
  // and bias both present or both absent (LDB200_NO_FAST_EPI=1: always the generic epilogue)
  static const bool fast_env = getenv("LDB200_NO_FAST_EPI") == nullptr;
  int feat = -1;
  if (fast_env && ext == 0 && p.tma_epi && d->Cout % BN == 0 && !p.accumulate && !p.out_f32 && p.vec_ok &&
      (p.bias != nullptr || p.scale == nullptr)) {
    bool ok = true;
    for (int i = 0; i < count; ++i)
      ok = ok && (p.seg[i].tl_add == (p.seg[i].add_src ? 1 : 0)) && (p.seg[i].tl_mask == (p.seg[i].mask_src ? 1 : 0));
    if (ok)
      feat = (p.scale ? EF_AFFINE : (p.bias ? EF_BIAS : 0)) | (p.relu ? EF_RELU : 0) | (d->add_src ? EF_ADD : 0) | (d->mask_src ? EF_MASK : 0);
  }
  if (use2) return launch_conv2(maps, tmB, p, stream, ext, feat);
  if (shortk) return BN == 128 ? launch_conv<128, true>(maps, tmB, p, stream, ext, feat) : launch_conv<64, true>(maps, tmB, p, stream, ext, feat);
  if (BN == 256) return launch_conv<256, false>(maps, tmB, p, stream, ext, feat);
  if (BN == 128) return launch_conv<128, false>(maps, tmB, p, stream, ext, feat);
  return launch_conv<64, false>(maps, tmB, p, stream, ext, feat);
}

extern "C" int ldb200_conv2d_tc(const ldb_conv_desc* d, void* stream_) {
  LDB_REQUIRE(d, LDB_EINVAL, "conv2d_tc: null descriptor");
  return ldb200_conv2d_tc_multi(d, 1, stream_);
}

namespace ldb {

template <int BN>
static int launch_wgrad(const CUtensorMap& tmA, const CUtensorMap& tmB, const WgradParams& p, cudaStream_t stream) {
  using Cfg = WgradCfg<BN>;
  static bool attr = false;
  if (!attr) {
    LDB_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    attr = true;
  }
  const int units = p.co_tiles * p.ci_tiles * p.ntaps * p.splits;
  const int grid = units < g_conv_sms ? units : g_conv_sms;
  LDB_CUDA(launch_pdl(wgrad_tc_kernel<BN>, grid, CONV_THREADS, Cfg::SMEM, stream, tmA, tmB, p));
  return 0;
}

}  // namespace ldb

extern "C" int ldb200_conv2d_wgrad_tc(const ldb_wgrad_desc* d, void* stream_) {
  LDB_REQUIRE(d && d->dy && d->x && d->dw, LDB_EINVAL, "wgrad_tc: null pointer");
  LDB_REQUIRE(d->Cin > 0 && d->Cin % KCH == 0, LDB_ENOSUP, "wgrad_tc: Cin=%d must be a multiple of 64", d->Cin);
  LDB_REQUIRE(d->Cout > 0, LDB_EINVAL, "wgrad_tc: Cout");
  LDB_REQUIRE(d->dy_ld >= d->Cout && d->dy_ld % 8 == 0 && d->x_ld >= d->Cin && d->x_ld % 8 == 0, LDB_EINVAL,
              "wgrad_tc: leading dimensions");
  LDB_REQUIRE(d->ntaps >= 1 && d->ntaps <= 9 && d->sx >= 1 && d->sy >= 1, LDB_EINVAL, "wgrad_tc: taps/stride");
  LDB_REQUIRE(d->w_k % 4 == 0 && ((uintptr_t)d->dw & 15) == 0, LDB_EINVAL, "wgrad_tc: dw alignment");
  LDB_REQUIRE(((uintptr_t)d->dy & 15) == 0 && ((uintptr_t)d->x & 15) == 0, LDB_EINVAL, "wgrad_tc: pointer alignment");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!g_conv_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_conv_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->N; p.Cout = d->Cout; p.Cin = d->Cin; p.ntaps = d->ntaps;
  for (int t = 0; t < d->ntaps; ++t) {
    p.tap_dx[t] = d->tap_dx[t]; p.tap_dy[t] = d->tap_dy[t]; p.tap_wk[t] = d->tap_wk[t];
  }
  p.sx = d->sx; p.sy = d->sy; p.dw = d->dw; p.w_k = d->w_k;
  p.scale = d->scale; p.w_master = d->w_master; p.d_scale = d->d_scale;
  LDB_REQUIRE(d->split == 0 || d->split == 2 || d->split == 3, LDB_EINVAL, "wgrad_tc: split must be 0, 2 or 3");
  p.planes = d->split ? d->split : 1;
  LDB_REQUIRE(d->split_pass >= 0 && d->split_pass <= 2 && (d->split || !d->split_pass), LDB_EINVAL, "wgrad_tc: split_pass");
  if (d->split == 3) {
    static const int ca[6] = {2, 0, 1, 1, 0, 0}, cb[6] = {0, 2, 1, 0, 1, 0};
    const int lo = d->split_pass == 1 ? 5 : 0, hi = d->split_pass == 2 ? 5 : 6;
    p.ncombo = hi - lo;
    for (int i = lo; i < hi; ++i) { p.comb_a[i - lo] = ca[i]; p.comb_b[i - lo] = cb[i]; }
  } else if (d->split == 2) {
    static const int ca[3] = {1, 0, 0}, cb[3] = {0, 1, 0};
    const int lo = d->split_pass == 1 ? 2 : 0, hi = d->split_pass == 2 ? 2 : 3;
    p.ncombo = hi - lo;
    for (int i = lo; i < hi; ++i) { p.comb_a[i - lo] = ca[i]; p.comb_b[i - lo] = cb[i]; }
  } else {
    p.ncombo = 1;
  }
  LDB_REQUIRE(!d->d_scale || (d->w_master && ((uintptr_t)d->w_master & 15) == 0), LDB_EINVAL, "wgrad_tc: d_scale needs aligned w_master");
  int best_th = 8, best_tiles = 1 << 30;
  for (int th = 1; th <= TILE_M; th *= 2) {
    const int tw = TILE_M / th;
    if (tw * d->sx > 256 || th * d->sy > 256) continue;
    const int tiles = ((d->HO + th - 1) / th) * ((d->WO + tw - 1) / tw);
    if (tiles < best_tiles || (tiles == best_tiles && abs(th - 8) < abs(best_th - 8))) {
      best_tiles = tiles;
      best_th = th;
    }
  }
  p.TH = best_th; p.TW = TILE_M / best_th;
  p.tiles_y = (d->HO + p.TH - 1) / p.TH;
  p.tiles_x = (d->WO + p.TW - 1) / p.TW;
  const int BN = d->Cin > 128 ? 256 : (d->Cin > 64 ? 128 : 64);
  p.co_tiles = (d->Cout + TILE_M - 1) / TILE_M;
  p.ci_tiles = (d->Cin + BN - 1) / BN;
  const int ptiles = p.N * p.tiles_y * p.tiles_x;
  const int base_units = p.co_tiles * p.ci_tiles * p.ntaps;
  // one wave of work units: every extra split costs a 128 x BN fp32 tile of L2 atomics and the mainloop per unit
  // shrinks, so fill the SMs once (round to nearest) instead of twice
  int splits = (g_conv_sms + base_units / 2) / base_units;
  if (splits > ptiles) splits = ptiles;
  if (splits < 1) splits = 1;
  p.splits = splits;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)d->WO, (cuuint64_t)d->HO, (cuuint64_t)d->N * p.planes};
    cuuint64_t strides[3] = {(cuuint64_t)d->dy_ld * 2, (cuuint64_t)d->dy_ld * 2 * d->WO,
                             (cuuint64_t)d->dy_ld * 2 * d->WO * d->HO};
    cuuint32_t box[4] = {KCH, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1}, es[4] = {1, 1, 1, 1};
    int e = encode_map(&tmA, d->dy, 4, dims, strides, box, es, "wgrad dY");
    if (e) return e;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N * p.planes};
    cuuint64_t strides[3] = {(cuuint64_t)d->x_ld * 2, (cuuint64_t)d->x_ld * 2 * d->W,
                             (cuuint64_t)d->x_ld * 2 * d->W * d->H};
    cuuint32_t box[4] = {KCH, (cuuint32_t)(p.TW * d->sx), (cuuint32_t)(p.TH * d->sy), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)d->sx, (cuuint32_t)d->sy, 1};
    int e = encode_map(&tmB, d->x, 4, dims, strides, box, es, "wgrad X");
    if (e) return e;
  }
  static const bool wgrad2_env = getenv("LDB200_NO_WGRAD2") == nullptr;
  if (wgrad2_env && BN == 256 && d->Cout % 256 == 0 && d->Cin % 256 == 0) {
    static int max_clusters = 0;
    if (!max_clusters) {
      LDB_CUDA(cudaFuncSetAttribute(wgrad2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Wgrad2Cfg::SMEM));
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(g_conv_sms / 2 * 2);
      cfg.blockDim = dim3(CONV_THREADS);
      cfg.dynamicSmemBytes = Wgrad2Cfg::SMEM;
      int n = 0;
      LDB_CUDA(cudaOccupancyMaxActiveClusters(&n, wgrad2_tc_kernel, &cfg));
      LDB_REQUIRE(n >= 1, LDB_ENOSUP, "wgrad_tc: no resident 2-CTA cluster");
      max_clusters = n < g_conv_sms / 2 ? n : g_conv_sms / 2;
    }
    p.co_tiles = d->Cout / 256;
    const int base2 = p.co_tiles * p.ci_tiles * p.ntaps;
    int sp = (max_clusters + base2 / 2) / base2;
    if (sp > ptiles) sp = ptiles;
    if (sp < 1) sp = 1;
    p.splits = sp;
    const int units = base2 * sp;
    const int clusters = units < max_clusters ? units : max_clusters;
    wgrad2_tc_kernel<<<2 * clusters, CONV_THREADS, Wgrad2Cfg::SMEM, stream>>>(tmA, tmB, p);
    LDB_CUDA(cudaGetLastError());
    return 0;
  }
  if (BN == 256) return launch_wgrad<256>(tmA, tmB, p, stream);
  if (BN == 128) return launch_wgrad<128>(tmA, tmB, p, stream);
  return launch_wgrad<64>(tmA, tmB, p, stream);
}
