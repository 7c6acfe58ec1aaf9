// tools/micro/softmax_loop.cu — how fast can ONE warp per SMSP run the forward kernel's exponentiation loop?
// Replicates exp_group of fa_fwd_sm100.cuh on register data (128 scores per thread, compiler-scheduled, no volatile
// ordering between the arithmetic instructions) and times it with clock64, with parts of the mix removed, for 1 and 2
// warps per SMSP.  Answers whether MUFU.EX2 and the FMA-pipe work of the same warp overlap.
// Not part of the product.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ unsigned pack(float2 v) {
  __nv_bfloat162 b = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<unsigned*>(&b);
}

// bit 0: FFMA2 (scale/subtract), bit 1: MUFU, bit 2: FADD2 (row sum), bit 3: F2FP (pack)
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}

// SPIN: threads >= 128 (one or more extra warps per SMSP) spin on an mbarrier, as the kernel's waiting warps do, until
// the four measured warps are done.
template <int MIX, bool SPIN>
__global__ void __launch_bounds__(256, 1) k(float* out, const float* in, int iters, long long* cyc) {
  __shared__ __align__(8) unsigned long long bar;
  if (SPIN) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bar)), "r"(128));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x >= 128) {
      while (!try_wait(smem_addr(&bar), 0)) {}
      return;
    }
  }
  float s[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) s[i] = in[(threadIdx.x * 128 + i) & 1023];
  float c = in[3], nm = in[5];
  float2 ls[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
  unsigned acc = 0;
  if (!SPIN) __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const float2 c2 = make_float2(c, c), nm2 = make_float2(nm, nm);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      unsigned pk[16];
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        float2 x = make_float2(s[q * 32 + 2 * kk], s[q * 32 + 2 * kk + 1]);
        if (MIX & 1) x = __ffma2_rn(x, c2, nm2);
        float2 p = x;
        if (MIX & 2) { p.x = ex2(x.x); p.y = ex2(x.y); }
        if (MIX & 4) ls[kk & 3] = __fadd2_rn(ls[kk & 3], p);
        if (MIX & 8) pk[kk] = pack(p); else pk[kk] = __float_as_uint(p.x) ^ __float_as_uint(p.y);
      }
      // the kernel stores the 16 packed registers to TMEM here (one STTM); here they are folded into a checksum with
      // one 3-input LOP3 per two registers (0.5 ALU instruction per pair more than the kernel issues)
#pragma unroll
      for (int kk = 0; kk < 16; kk += 2) acc ^= pk[kk] ^ pk[kk + 1];
    }
    // a loop-carried dependence of no cost, so iterations cannot be merged
    asm volatile("" : "+f"(c), "+f"(nm));
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) *cyc = t1 - t0;
  if (SPIN) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&bar)) : "memory");
  float r = 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u) r += ls[u].x + ls[u].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r + __uint_as_float(acc);
}

template <int MIX, bool SPIN = false>
void run(const char* name, int warps_per_smsp, float* in) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 400, threads = 128 * warps_per_smsp;
  k<MIX, SPIN><<<1, threads>>>(out, in, 10, cyc);
  k<MIX, SPIN><<<1, threads>>>(out, in, iters, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_tile = (double)h / iters;
  if (SPIN) warps_per_smsp = 1;   // one measured warp per SMSP, the others spin
  printf("%-40s warps/SMSP=%d : %7.1f clk per 64-pair row  = %.2f clk per pair per warp, %.2f per pair per SMSP\n", name,
         warps_per_smsp, per_tile, per_tile / 64, per_tile / 64 / warps_per_smsp);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  float* in; cudaMalloc(&in, 4096);
  float h[1024];
  for (int i = 0; i < 1024; ++i) h[i] = -0.01f * (i % 37);
  h[3] = 1.0001f; h[5] = -0.25f;
  cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
  for (int w = 1; w <= 2; ++w) {
    run<2>("MUFU only", w, in);
    run<2 | 8>("MUFU + F2FP", w, in);
    run<2 | 1>("MUFU + FFMA2", w, in);
    run<2 | 4>("MUFU + FADD2", w, in);
    run<1 | 4 | 8>("FFMA2 + FADD2 + F2FP (no MUFU)", w, in);
    run<1 | 2 | 8>("FFMA2 + MUFU + F2FP (no row sum)", w, in);
    run<15>("full mix (FFMA2, 2 MUFU, FADD2, F2FP)", w, in);
  }
  // one measured warp per SMSP + one warp per SMSP spinning on an mbarrier (255 registers per thread cap the block at 256)
  run<15, true>("full mix + 1 spinning warp/SMSP", 2, in);
  return 0;
}
