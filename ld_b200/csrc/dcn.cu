// Deformable convolution v1, FORWARD (the frozen X101-DCN / R101-DCN teachers of the LD configs:
// configs/imv2/gflv2_x101_fpn_2x_coco.py:17-18, configs/gfl/gfl_r101_fpn_dconv_c3-c5_mstrain_2x_coco.py:11-12;
// mmdet/models/backbones/resnext.py:64-75 -> mmcv.ops.DeformConv2dPack, whose source is NOT in /root/reference:
// restated from the published DCNv1 algorithm -- bilinear-sampled im2col followed by a GEMM -- with the offset layout
// mmcv and torchvision.ops.deform_conv2d share: channel 2*(g*K + k) = dy, 2*(g*K + k) + 1 = dx of kernel point k of
// deformable group g).
//
//   col[n, yo, xo, k*C + c] = bilinear(x[n, :, :, c], yo*sy - pad + (k / kw)*dil + dy, xo*sx - pad + (k % kw)*dil + dx)
//
// samples outside (-1, H) x (-1, W) are zero; corner taps outside the image contribute zero.  The tensor-core GEMM
// that follows is the ordinary flat 1x1 path of ldb200_conv2d_tc on [N*HO*WO, K*C] with the conv's packed weights
// ([Cout, kh*kw*Cin], the same layout as a regular 3x3), so BN / ReLU fuse into its epilogue as everywhere else.
#include "common.cuh"

namespace ldb {

template <typename T> struct Vec8;
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float v[8]) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __bfloat1622float2(h[e]);
      v[2 * e] = f.x;
      v[2 * e + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float v[8]) {
    uint4 q;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    *reinterpret_cast<uint4*>(p) = q;
  }
};
template <> struct Vec8<float> {
  static __device__ __forceinline__ void ld(const float* p, float v[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void st(float* p, const float v[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};

// one thread = one (output pixel, kernel point, 8 channels): four 16/32-byte corner reads, one store.  Consecutive
// threads walk the channels of one kernel point, so a warp reads 4 x 512 (1024) contiguous bytes per corner row.
template <typename T>
__global__ void __launch_bounds__(256)
deform_im2col_kernel(const T* __restrict__ x, const float* __restrict__ offset, T* __restrict__ col, int N, int H, int W,
                     int C, int HO, int WO, int kh, int kw, int sy, int sx, int pad, int dil, int dgroups, int off_ld) {
  const int cg = C / 8, K = kh * kw;
  const int cpg = C / dgroups;    // channels per deformable group
  const long long total = (long long)N * HO * WO * K * cg;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int g8 = (int)(t % cg);
    long long r = t / cg;
    const int k = (int)(r % K);
    const long long pix = r / K;
    const int xo = (int)(pix % WO), yo = (int)((pix / WO) % HO), n = (int)(pix / ((long long)WO * HO));
    const int dg = (g8 * 8) / cpg;
    const float* op = offset + (size_t)pix * off_ld + 2 * (dg * K + k);
    const float py = (float)(yo * sy - pad + (k / kw) * dil) + __ldg(op);
    const float px = (float)(xo * sx - pad + (k % kw) * dil) + __ldg(op + 1);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    if (py > -1.f && px > -1.f && py < (float)H && px < (float)W) {
      const int y0 = (int)floorf(py), x0 = (int)floorf(px);
      const float ly = py - (float)y0, lx = px - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
      const T* base = x + (size_t)n * H * W * C + g8 * 8;
      float v[8];
      if (y0 >= 0 && x0 >= 0) {
        Vec8<T>::ld(base + ((size_t)y0 * W + x0) * C, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += hy * hx * v[e];
      }
      if (y0 >= 0 && x0 + 1 <= W - 1) {
        Vec8<T>::ld(base + ((size_t)y0 * W + x0 + 1) * C, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += hy * lx * v[e];
      }
      if (y0 + 1 <= H - 1 && x0 >= 0) {
        Vec8<T>::ld(base + ((size_t)(y0 + 1) * W + x0) * C, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += ly * hx * v[e];
      }
      if (y0 + 1 <= H - 1 && x0 + 1 <= W - 1) {
        Vec8<T>::ld(base + ((size_t)(y0 + 1) * W + x0 + 1) * C, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += ly * lx * v[e];
      }
    }
    Vec8<T>::st(col + ((size_t)pix * K + k) * C + g8 * 8, acc);
  }
}

}  // namespace ldb

using namespace ldb;

extern "C" int ldb200_deform_im2col(const void* x, const float* offset, void* col, int N, int H, int W, int C, int HO,
                                    int WO, int kh, int kw, int stride, int pad, int dilation, int deform_groups,
                                    int offset_ld, int dtype, void* stream_) {
  LDB_REQUIRE(x && offset && col, LDB_EINVAL, "deform_im2col: null pointer");
  LDB_REQUIRE(N > 0 && H > 0 && W > 0 && HO > 0 && WO > 0 && kh > 0 && kw > 0 && stride > 0 && dilation > 0, LDB_EINVAL,
              "deform_im2col: shape");
  LDB_REQUIRE(C % 8 == 0 && deform_groups >= 1 && C % deform_groups == 0 && (C / deform_groups) % 8 == 0, LDB_EINVAL,
              "deform_im2col: channels per deformable group must be a multiple of 8");
  LDB_REQUIRE(offset_ld >= 2 * deform_groups * kh * kw, LDB_EINVAL, "deform_im2col: offset_ld");
  const long long total = (long long)N * HO * WO * kh * kw * (C / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  cudaStream_t s = (cudaStream_t)stream_;
  if (dtype == LDB_F32)
    deform_im2col_kernel<float><<<(unsigned)blocks, 256, 0, s>>>((const float*)x, offset, (float*)col, N, H, W, C, HO, WO, kh,
                                                                 kw, stride, stride, pad, dilation, deform_groups, offset_ld);
  else
    deform_im2col_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, s>>>((const __nv_bfloat16*)x, offset, (__nv_bfloat16*)col,
                                                                         N, H, W, C, HO, WO, kh, kw, stride, stride, pad,
                                                                         dilation, deform_groups, offset_ld);
  LDB_CUDA(cudaGetLastError());
  return 0;
}
