// Host build of pecanpy_b200/csrc/b2w_pairwise.cuh (test infrastructure: checks the header's arithmetic
// against NumPy on a box without a GPU; the product only ever runs this code inside CUDA kernels).
#include <cstdint>
#include "../pecanpy_b200/csrc/b2w_pairwise.cuh"

namespace {
template <typename T> T clip0(T x) { return (x >= (T)0 || x != x) ? x : (T)0; }

struct Raw { const float* p; float next() { return *p++; } };
struct Sq { const float* p; float mean; float next() { float x = *p++ - mean; return x * x; } };
struct Make { const float* row; Raw raw() const { return Raw{row}; } Sq centered_sq(float m) const { return Sq{row, m}; } };

struct DStream {
  const double* row; const uint8_t* nz; uint32_t c; bool sq; double mean;
  double next() { while (!nz[c]) ++c; double x = row[c++]; if (sq) { double t = x - mean; x = t * t; } return x; }
};
struct DMake {
  const double* row; const uint8_t* nz;
  DStream raw() const { return DStream{row, nz, 0u, false, 0.0}; }
  DStream centered_sq(double m) const { return DStream{row, nz, 0u, true, m}; }
};
}  // namespace

extern "C" float h_sum_f32(const float* a, uint32_t n) { Raw r{a}; return b2w_pairwise_sum<float>(r, n); }

extern "C" void h_thr_csr(uint32_t n, const uint32_t* indptr, const float* data, double gamma, float* thr) {
  const float g = (float)gamma;
  for (uint32_t i = 0; i < n; ++i) {
    float mean, sd;
    b2w_mean_std<float>(Make{data + indptr[i]}, indptr[i + 1] - indptr[i], mean, sd);
    thr[i] = clip0(mean + g * sd);
  }
}

extern "C" void h_thr_dense(uint32_t n, const double* data, const uint8_t* nz, double gamma, float* thr) {
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t cnt = 0;
    for (uint32_t c = 0; c < n; ++c) cnt += nz[(uint64_t)i * n + c] != 0;
    double mean, sd;
    b2w_mean_std<double>(DMake{data + (uint64_t)i * n, nz + (uint64_t)i * n}, cnt, mean, sd);
    thr[i] = clip0((float)(mean + gamma * sd));
  }
}
