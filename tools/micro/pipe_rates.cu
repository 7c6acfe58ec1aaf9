// tools/micro/pipe_rates.cu — per-SMSP issue-rate microbenchmark for the instructions of the softmax inner loop
// (MUFU.EX2, F2FP bf16x2 pack, FFMA2, FADD2, FMNMX3) alone and mixed, 1 or 2 warps per SMSP.
// Used to decide which pipe bounds the softmax warpgroups.  Not part of the product.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>

template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = -0.001f * (threadIdx.x + i);
  unsigned acc = 0;
  float2 s2 = make_float2(0.f, 0.f);
  float mx = -1e30f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      if (MODE == 0 || MODE == 2 || MODE == 5) {          // MUFU
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i + 1]));
      }
      if (MODE == 1 || MODE == 2 || MODE == 5) {          // F2FP
        __nv_bfloat162 v = __floats2bfloat162_rn(x[i], x[i + 1]);
        unsigned u = *reinterpret_cast<unsigned*>(&v);
        asm volatile("" : "+r"(u));
        acc ^= u;
      }
      if (MODE == 3 || MODE == 5) {                        // FFMA2 + FADD2
        float2 a = make_float2(x[i], x[i + 1]);
        a = __ffma2_rn(a, make_float2(1.0001f, 1.0001f), make_float2(-1e-6f, -1e-6f));
        s2 = __fadd2_rn(s2, a);
        if (MODE == 3) { x[i] = a.x; x[i + 1] = a.y; }
      }
      if (MODE == 4 || MODE == 5) {                        // FMNMX3
        mx = fmaxf(mx, fmaxf(x[i], x[i + 1]));
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  float r = mx + s2.x + s2.y;
  for (int i = 0; i < 16; ++i) r += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r + acc;
}

template <int MODE>
void run(const char* name, int warps_per_smsp, int ops_per_pair) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 2000, threads = 128 * warps_per_smsp;
  k<MODE><<<1, threads>>>(out, 10, cyc);
  k<MODE><<<1, threads>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double pairs = (double)iters * 8 * warps_per_smsp;   // element pairs per SMSP
  printf("%-28s warps/SMSP=%d : %.2f clk per element pair per SMSP (%.2f clk per warp-instruction of the mix, %d instr/pair)\n",
         name, warps_per_smsp, h / pairs, h / pairs / ops_per_pair, ops_per_pair);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w = 1; w <= 2; ++w) {
    run<0>("MUFU.EX2 x2", w, 2);
    run<1>("F2FP.BF16 pack x1", w, 1);
    run<2>("MUFU x2 + F2FP", w, 3);
    run<3>("FFMA2 + FADD2", w, 2);
    run<4>("FMNMX3", w, 1);
    run<5>("full mix (2 MUFU,F2FP,FFMA2,FADD2,FMNMX3)", w, 6);
  }
  return 0;
}
