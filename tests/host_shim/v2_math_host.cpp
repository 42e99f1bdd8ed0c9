// Host build of ld_b200/csrc/v2_math.cuh for the CPU test tests/test_v2_math_host.py (test infrastructure only: the
// product calls the same inline functions from the CUDA kernels in ld_b200/csrc/dgqp.cu).
#include "../../ld_b200/csrc/v2_math.cuh"

using namespace ldb;

extern "C" {

// rows x [4*bins] logits -> rows quality values
void host_dgqp_forward(const float* reg, int rows, int bins, int topk, int add_mean, int hid, const float* W1,
                       const float* b1, const float* W2, float b2, float* q) {
  DgqpShape s{bins, topk, add_mean, hid};
  for (int r = 0; r < rows; ++r) q[r] = dgqp_row_forward(reg + (size_t)r * 4 * bins, s, W1, b1, W2, b2);
}

// dq [rows] -> d_reg [rows, 4*bins] (overwritten), dW1 [hid*nstat], db1 [hid], dW2 [hid], db2 [1] (overwritten)
void host_dgqp_backward(const float* reg, int rows, int bins, int topk, int add_mean, int hid, const float* W1,
                        const float* b1, const float* W2, float b2, const float* dq, float* d_reg, float* dW1,
                        float* db1, float* dW2, float* db2) {
  DgqpShape s{bins, topk, add_mean, hid};
  const int ns = s.nstat();
  for (int i = 0; i < hid * ns; ++i) dW1[i] = 0.f;
  for (int j = 0; j < hid; ++j) db1[j] = dW2[j] = 0.f;
  *db2 = 0.f;
  for (int r = 0; r < rows; ++r) {
    float* dr = d_reg + (size_t)r * 4 * bins;
    for (int i = 0; i < 4 * bins; ++i) dr[i] = 0.f;
    float dhpre[DGQP_MAX_HID], stat[DGQP_MAX_STAT], h[DGQP_MAX_HID];
    const float d_o = dgqp_row_backward(reg + (size_t)r * 4 * bins, s, W1, b1, W2, b2, dq[r], dr, dhpre, stat, h);
    for (int j = 0; j < hid; ++j) {
      for (int i = 0; i < ns; ++i) dW1[j * ns + i] += dhpre[j] * stat[i];
      db1[j] += dhpre[j];
      dW2[j] += d_o * h[j];
    }
    *db2 += d_o;
  }
}

void host_qflp(const float* p, const float* y, const int* is_pos, int n, float beta, float* loss, float* grad) {
  for (int i = 0; i < n; ++i) {
    if (is_pos[i]) qflp_pos(p[i], y[i], beta, loss[i], grad[i]);
    else qflp_neg(p[i], beta, loss[i], grad[i]);
  }
}

}  // extern "C"
