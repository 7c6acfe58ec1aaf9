// tools/micro/softmax_tile.cu — the forward kernel's per-tile softmax loop in isolation (one CTA, no MMA, no TMA):
// tcgen05.ld of 128 score columns, row max, exponentials (FFMA2, MUFU.EX2, FADD2, F2FP), tcgen05.st of P, wait::st, fence,
// mbarrier arrive — timed with clock64 with pieces switched off, with 1 or 2 warps per SMSP.  Tells whether the softmax
// warp's own instruction stream takes the time the kernel's timeline shows or whether the rest of
// the CTA (tensor pipe traffic to TMEM, the other warps) slows it down.  Not part of the product.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>

#include "../../flash_attention_impls_b200/csrc/sm100_ptx.cuh"

using namespace fa;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ unsigned pack(float2 v) {
  __nv_bfloat162 b = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<unsigned*>(&b);
}

// bits: 1 tcgen05.ld of S   2 row max   4 tcgen05.st of P   8 wait::st + fence   16 mbarrier arrive (twice per tile)
template <int OPT>
__global__ void __launch_bounds__(256, 1) k(float* out, float c, int iters, long long* cyc) {
  __shared__ __align__(8) unsigned long long bar[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar[0]), blockDim.x);
    mbar_init(smem_u32(&bar[1]), blockDim.x);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc<512>(smem_u32(&tmem_slot));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t lane_addr = uint32_t((warp & 3) * 32) << 16;
  const uint32_t tS = tmem_base + lane_addr + (warp >> 2) * 128;

  uint32_t sr[4][32];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int i = 0; i < 32; ++i) sr[q][i] = __float_as_uint(-0.01f * ((threadIdx.x + q * 32 + i) % 37));
  // put finite scores into TMEM once, so that tcgen05.ld returns something sensible
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    tmem_st16(tS + q * 32, &sr[q][0]);
    tmem_st16(tS + q * 32 + 16, &sr[q][16]);
  }
  tmem_wait_st();
  float m_ref = 0.f, l = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (OPT & 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) tmem_ld32(tS + q * 32, sr[q]);
      tmem_wait_ld();
    }
    if (OPT & 2) {
      float mxp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int kk = 0; kk < 32; kk += 8)
#pragma unroll
          for (int u = 0; u < 4; ++u)
            mxp[u] = fmaxf(mxp[u], fmaxf(__uint_as_float(sr[q][kk + 2 * u]), __uint_as_float(sr[q][kk + 2 * u + 1])));
      m_ref = fmaxf(m_ref, fmaxf(fmaxf(mxp[0], mxp[1]), fmaxf(mxp[2], mxp[3])) * c);
    }
    const float2 c2 = make_float2(c, c), nm2 = make_float2(-m_ref, -m_ref);
    float2 ls[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t pk[16];
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        const float2 x = __ffma2_rn(make_float2(__uint_as_float(sr[q][2 * kk]), __uint_as_float(sr[q][2 * kk + 1])), c2, nm2);
        float2 p;
        p.x = ex2(x.x);
        p.y = ex2(x.y);
        ls[kk & 3] = __fadd2_rn(ls[kk & 3], p);
        pk[kk] = pack(p);
      }
      if (OPT & 4) {
        tmem_st16(tS + q * 16, pk);
      } else {
        unsigned acc = 0;
#pragma unroll
        for (int kk = 0; kk < 16; kk += 2) acc ^= pk[kk] ^ pk[kk + 1];
        l += __uint_as_float(acc & 1u);
      }
      if (q == 1 || q == 3) {
        if (OPT & 8) {
          tmem_wait_st();
          tc_fence_before();
        }
        if (OPT & 16) mbar_arrive(smem_u32(&bar[q >> 1]));
      }
    }
    const float2 r = __fadd2_rn(__fadd2_rn(ls[0], ls[1]), __fadd2_rn(ls[2], ls[3]));
    l += r.x + r.y;
    if (!(OPT & 1)) asm volatile("" : "+f"(c));   // keep iterations apart when nothing is reloaded
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) *cyc = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = l + m_ref;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int OPT>
void run(const char* name, int warps_per_smsp) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 400, threads = 128 * warps_per_smsp;
  k<OPT><<<1, threads>>>(out, 1.0001f, 10, cyc);
  k<OPT><<<1, threads>>>(out, 1.0001f, iters, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-58s warps/SMSP=%d : %7.1f clk per tile row (%s)\n", name, warps_per_smsp, (double)h / iters, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w = 1; w <= 2; ++w) {
    run<0>("exponentials only (registers)", w);
    run<4>("+ tcgen05.st of P", w);
    run<4 | 8>("+ st, wait::st, fence", w);
    run<4 | 8 | 16>("+ st, wait::st, fence, arrive", w);
    run<1 | 4 | 8 | 16>("+ tcgen05.ld of S (no max)", w);
    run<1 | 2>("ld + max + exponentials, no stores", w);
    run<1 | 2 | 4 | 8 | 16>("full softmax tile: ld, max, exp, st, wait, arrive", w);
  }
  return 0;
}
