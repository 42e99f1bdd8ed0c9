// Input side of the training path: the image half of the reference's train pipeline
//   Resize(img_scale, keep_ratio=True) -> RandomFlip -> Normalize(mean, std, to_rgb) -> Pad(size_divisor) ->
//   DefaultFormatBundle / collate          (mmdet/datasets/pipelines/transforms.py:26-300, 303-470, 524-600, 473-560;
//                                           configs/ld/ld_r18_gflv1_r101_fpn_coco_1x.py:66-77)
// for a whole batch in ONE launch, from the decoded uint8 HWC (BGR) images to the [N, 3, Hp, Wp] fp32 network input.
// mmcv.imrescale / imflip / imnormalize / impad are cv2 / numpy calls (mmcv is not vendored in the reference); their
// arithmetic is restated exactly, so the result is bit-identical to the CPU pipeline:
//   * bilinear resize = cv2.resize(INTER_LINEAR) on 8-bit data: 11-bit fixed-point coefficients
//     (cvRound(f * 2048), f in fp32 from the fp64 source coordinate), horizontal pass in int32, vertical pass
//     ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2; horizontal coefficients are zeroed at the borders,
//     vertical ones are not (only the row index is clamped) -- as OpenCV's resizeGeneric does;
//   * flip = index reversal along x; to_rgb = channel reversal;
//   * normalize = float(double(float(v) - mean) * (1 / double(std)))      (cv2.subtract, cv2.multiply by a double scalar);
//   * padding (to the divisor and to the largest image of the batch) = zeros.
#include "common.cuh"

namespace ldb {

struct PrepImage {
  long long src_off;     // byte offset of the image in the packed uint8 buffer
  int sh, sw;            // source height / width
  int dh, dw;            // resized height / width
  int flip;              // horizontal flip after the resize
};

__device__ __forceinline__ void lin_coef(int d, int dsize, int ssize, bool zero_at_border, int& i0, int& i1, int& a0, int& a1) {
  const double scale = (double)ssize / (double)dsize;
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (zero_at_border) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
  }
  a0 = max(-32768, min(32767, __float2int_rn((1.f - f) * 2048.f)));
  a1 = max(-32768, min(32767, __float2int_rn(f * 2048.f)));
  i0 = max(0, min(ssize - 1, s));
  i1 = max(0, min(ssize - 1, s + 1));
}

// one thread per output pixel (all three channels); grid.z = image
__global__ void __launch_bounds__(256)
preprocess_kernel(const unsigned char* __restrict__ src, const PrepImage* __restrict__ imgs, float* __restrict__ out,
                  int Hp, int Wp, float m0, float m1, float m2, double s0, double s1, double s2, int to_rgb) {
  const int n = blockIdx.z;
  const PrepImage im = imgs[n];
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= Wp) return;
  float* o = out + (size_t)n * 3 * Hp * Wp + (size_t)y * Wp + x;
  if (y >= im.dh || x >= im.dw) {       // Pad / collate region
    o[0] = 0.f; o[(size_t)Hp * Wp] = 0.f; o[2 * (size_t)Hp * Wp] = 0.f;
    return;
  }
  const int xr = im.flip ? im.dw - 1 - x : x;      // column of the resized (unflipped) image
  int x0, x1, ax0, ax1, y0, y1, ay0, ay1;
  lin_coef(xr, im.dw, im.sw, true, x0, x1, ax0, ax1);
  lin_coef(y, im.dh, im.sh, false, y0, y1, ay0, ay1);
  const unsigned char* p = src + im.src_off;
  const unsigned char* r0 = p + (size_t)y0 * im.sw * 3;
  const unsigned char* r1 = p + (size_t)y1 * im.sw * 3;
  const float mean[3] = {m0, m1, m2};
  const double sinv[3] = {s0, s1, s2};
#pragma unroll
  for (int c = 0; c < 3; ++c) {          // c = channel of the SOURCE (BGR) image
    const int h0 = (int)r0[x0 * 3 + c] * ax0 + (int)r0[x1 * 3 + c] * ax1;
    const int h1 = (int)r1[x0 * 3 + c] * ax0 + (int)r1[x1 * 3 + c] * ax1;
    int v = (((ay0 * (h0 >> 4)) >> 16) + ((ay1 * (h1 >> 4)) >> 16) + 2) >> 2;
    v = max(0, min(255, v));
    const int oc = to_rgb ? 2 - c : c;   // mean / std are given in the OUTPUT channel order
    const float d = (float)v - mean[oc];
    o[(size_t)oc * Hp * Wp] = (float)((double)d * sinv[oc]);
  }
}

}  // namespace ldb

using namespace ldb;

extern "C" int ldb200_preprocess_batch(const void* src_u8, const ldb_prep_image* imgs_dev, int N, float* out, int Hp, int Wp,
                                       const float* mean3_host, const float* std3_host, int to_rgb, void* stream_) {
  LDB_REQUIRE(src_u8 && imgs_dev && out && mean3_host && std3_host && N > 0 && Hp > 0 && Wp > 0, LDB_EINVAL,
              "preprocess_batch: bad argument");
  static_assert(sizeof(ldb_prep_image) == sizeof(PrepImage), "ldb_prep_image layout");
  dim3 grid((Wp + 255) / 256, Hp, N);
  preprocess_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(
      (const unsigned char*)src_u8, (const PrepImage*)imgs_dev, out, Hp, Wp, mean3_host[0], mean3_host[1], mean3_host[2],
      1.0 / (double)std3_host[0], 1.0 / (double)std3_host[1], 1.0 / (double)std3_host[2], to_rgb);
  LDB_CUDA(cudaGetLastError());
  return 0;
}
