// oracle/ref_fa1_wrapper.cu — TEST INFRASTRUCTURE ONLY.
// Builds the reference's own FA1 kernel (code/cuda_fa1/flashAttention.cu:7-152) UNMODIFIED, from the
// source where it lies under /root/reference, and exposes a C launcher that reproduces the reference's
// launch geometry (code/cuda_fa1/main.cu:377-385).  Output: oracle/_ref/libref_fa1.so (git-ignored).
#include "flashAttention.cu"   // resolved through -I$(REF)/code/cuda_fa1 ; not copied into this repo

extern "C" int ref_fa1_forward(const void* Q, const void* K, const void* V, void* O, float* l, float* m,
                               int B, int H, int N, int d, int M, void* stream) {
  // launch math of main.cu:377-385
  int Bc = (int)ceilf((float)M / (4.0f * (float)d));
  int Br = (Bc < d) ? Bc : d;
  int Tr = (N + Br - 1) / Br;
  dim3 grid(Tr, B * H);
  dim3 block(Br);
  size_t shmem = (size_t)(Br * d + Bc * d + Bc * d) * sizeof(__half) + (size_t)(Br * d) * sizeof(float);
  if (shmem > 48 * 1024) return 2;       // the reference never opts in to > 48 KB
  if (d > 128 || Bc > 128) return 3;     // fixed float[128] arrays, flashAttention.cu:86,94,107
  flash_attention_forward<<<grid, block, shmem, (cudaStream_t)stream>>>(
      (const __half*)Q, (const __half*)K, (const __half*)V, (__half*)O, l, m, B, H, N, d, M);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
