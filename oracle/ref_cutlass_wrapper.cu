// oracle/ref_cutlass_wrapper.cu — TEST INFRASTRUCTURE ONLY.
// Builds the reference's WMMA kernel + dispatcher (code/cutlass_cuda_fa1/run/flash_attn_cutlass.cu,
// entry :519-544) UNMODIFIED from where it lies, against the CUTLASS tree vendored inside the
// reference, and exposes it through a C symbol.  Output: oracle/_ref/libref_cutlass.so (git-ignored).
#include "flash_attn_cutlass.cu"   // resolved through -I$(REF)/code/cutlass_cuda_fa1/run ; not copied

extern "C" int ref_cutlass_forward(const void* Q, const void* K, const void* V, void* O, int B, int H, int N,
                                   int d, void* stream) {
  flash_attention_cutlass_dispatch((const cutlass::half_t*)Q, (const cutlass::half_t*)K,
                                   (const cutlass::half_t*)V, (cutlass::half_t*)O, B, H, N, d,
                                   (cudaStream_t)stream);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
