// oracle/fixtures.cpp — TEST INFRASTRUCTURE ONLY.  Seeded input generators.
//
// Set R restates the reference's generator (code/cuda_fa1/main.cu:43-61, same in
// code/cutlass_cuda_fa1/run/test_flash_attn.cu:86-104): a FRESH std::mt19937(42) per tensor feeding
// std::normal_distribution<float>(0, 0.02), so Q, K and V receive identical values.  The stream is
// libstdc++-specific (normal_distribution is implementation-defined); SURVEY.md section 4 records its
// first eight values as built with this toolchain, and tests/test_oracle.py checks them.
//
// Set S is the stronger gating distribution of SURVEY.md section 8d: independent seeds, Q,K ~ N(0,1),
// V ~ U(-0.5,0.5), every value rounded to bf16 and |x| < 2^-14 flushed to zero, so the same reals are
// exactly representable in both fp16 (the reference's dtype) and bf16.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>

static float round_to_bf16_flush(float x) {
  if (std::fabs(x) < 6.103515625e-05f) return 0.f;   // 2^-14: below fp16's normal range
  uint32_t u;
  std::memcpy(&u, &x, 4);
  u += 0x7fffu + ((u >> 16) & 1u);                   // round to nearest even on the dropped 16 bits
  u &= 0xffff0000u;
  std::memcpy(&x, &u, 4);
  return x;
}

extern "C" {

// Set R: out[i] = N(mean, stddev) from mt19937(42) (main.cu:45-51).
void fixture_reference_stream(float* out, size_t n, float mean, float stddev) {
  std::mt19937 gen(42);
  std::normal_distribution<float> dist(mean, stddev);
  for (size_t i = 0; i < n; ++i) out[i] = dist(gen);
}

// Set S normal part: out[i] = bf16-rounded N(0, stddev) from mt19937(seed).
void fixture_normal_bf16(float* out, size_t n, uint32_t seed, float stddev) {
  std::mt19937 gen(seed);
  std::normal_distribution<float> dist(0.f, stddev);
  for (size_t i = 0; i < n; ++i) out[i] = round_to_bf16_flush(dist(gen));
}

// Set S uniform part: out[i] = bf16-rounded U(lo, hi) from mt19937(seed).
void fixture_uniform_bf16(float* out, size_t n, uint32_t seed, float lo, float hi) {
  std::mt19937 gen(seed);
  std::uniform_real_distribution<float> dist(lo, hi);
  for (size_t i = 0; i < n; ++i) out[i] = round_to_bf16_flush(dist(gen));
}

}  // extern "C"
