// fa_ref_shims.cu — C++-linkage functions with the reference's exact names and argument lists, so
// that callers written against the reference link against libfa_b200.so unchanged.
//
// The reference has no header for these: consumers re-declare them
// (code/cutlass_cuda_fa1/run/test_flash_attn.cu:22-58, perf_flash_attn_cutlass.cu:24-34).
//   flash_attention_cutlass_dispatch      flash_attn_cutlass.cu:519-544
//   flash_attention_forward_dispatch      flash_attn_unified.cu:545-571
//   flash_attention_small_tile_dispatch   flash_attn_unified.cu:573-599
//   attention_reference_dispatch          flash_attn_unified.cu:604-617
// Like the reference they return void and report problems on stderr; unlike it they do not
// silently skip the launch without a message.
#include <cuda_runtime.h>

#include <cstdio>

#include "../../include/fa_b200.h"

namespace cutlass {
struct half_t;  // only the name matters for linkage; bit-identical to __half (2-byte IEEE fp16)
}

namespace {
void run(const char* who, const void* Q, const void* K, const void* V, void* O, int B, int H, int N, int d,
         cudaStream_t stream) {
  int rc = fa_b200_forward_fp16(Q, K, V, O, B, H, N, d, stream);
  if (rc == FA_B200_ERR_HEAD_DIM)
    fprintf(stderr, "Unsupported head_dim=%d for %s (supported: multiples of 8 up to 128)\n", d, who);
  else if (rc != FA_B200_OK)
    fprintf(stderr, "%s failed: %s (%s)\n", who, fa_b200_status_string(rc), fa_b200_last_error());
}
}  // namespace

void flash_attention_cutlass_dispatch(const cutlass::half_t* Q, const cutlass::half_t* K, const cutlass::half_t* V,
                                      cutlass::half_t* O, int batch_size, int num_heads, int seq_len,
                                      int head_dim, cudaStream_t stream) {
  run("flash_attention_cutlass_dispatch", Q, K, V, O, batch_size, num_heads, seq_len, head_dim, stream);
}
void flash_attention_forward_dispatch(const cutlass::half_t* Q, const cutlass::half_t* K, const cutlass::half_t* V,
                                      cutlass::half_t* O, int batch_size, int num_heads, int seq_len,
                                      int head_dim, cudaStream_t stream) {
  run("flash_attention_forward_dispatch", Q, K, V, O, batch_size, num_heads, seq_len, head_dim, stream);
}
void flash_attention_small_tile_dispatch(const cutlass::half_t* Q, const cutlass::half_t* K, const cutlass::half_t* V,
                                         cutlass::half_t* O, int batch_size, int num_heads, int seq_len,
                                         int head_dim, cudaStream_t stream) {
  run("flash_attention_small_tile_dispatch", Q, K, V, O, batch_size, num_heads, seq_len, head_dim, stream);
}
void attention_reference_dispatch(const cutlass::half_t* Q, const cutlass::half_t* K, const cutlass::half_t* V,
                                  cutlass::half_t* O, int batch_size, int num_heads, int seq_len, int head_dim,
                                  cudaStream_t stream) {
  run("attention_reference_dispatch", Q, K, V, O, batch_size, num_heads, seq_len, head_dim, stream);
}
