// fa_bwd_sm100.cuh — attention backward for sm_100a (B200): dQ, dK, dV from dO and the forward's logsumexp.
//
// the only other algorithm of the reference is the Triton backward
//   code/triton_fa2/FA2-triton.py:98-170 (_bwd_kernel: recompute P from the saved row statistics, dV += P^T dO,
//   dP = dO V^T, dS, dQ += dS K, dK += dS^T Q with fp16 atomic adds), :207-237 (_FlashAttnFn.backward).
// Here it is two atomics-free tcgen05 kernels built from the operand modes the forward kernel already proved on
// hardware (SS K-major for the score-like products, TS with an MN-major shared-memory B operand for the products that
// consume the 16-bit probabilities from TMEM), plus an HBM-bound pre-pass delta_i = dO_i . O_i:
//
//   kDQ = true  ("dQ kernel"):   one CTA owns a 128-row Q tile (Q_i, dO_i resident) and streams the K/V tiles:
//        S = Q_i K_j^T,  dP = dO_i V_j^T,  P = exp(S*scale - lse_i),  dS = scale * P o (dP - delta_i),  dQ_i += dS K_j
//   kDQ = false ("dK/dV kernel"): one CTA owns a 128-row K/V tile (K_j, V_j resident) and streams the Q/dO tiles; the
//        score tile is computed TRANSPOSED so that keys are the TMEM lanes and P^T, dS^T are TMEM A operands as they are:
//        S^T = K_j Q_i^T,  dP^T = V_j dO_i^T,  P^T, dS^T as above with lse_i / delta_i indexed by COLUMN,
//        dV_j += P^T dO_i,  dK_j += dS^T Q_i
//   The dQ kernel recomputes S and dP (7 tile products in total instead of 5) and in exchange nothing is accumulated
//   through global memory: every output tile has exactly one writer, results are deterministic.
//
// Both are the same code: "fixed" tiles F1, F2 (A operands of the two score-like products), "streamed" tiles T1, T2
// (B operands, K-major for the score-like products and MN-major for the accumulating ones) through a 2-stage TMA ring.
//   warps 0-3 : compute warpgroup, one thread per TMEM lane (score row): P / dS, then the epilogue (TMA store)
//   warp  4   : TMA producer          warp 5 : tcgen05.mma issuer          warp 6 : TMEM allocator
// TMEM (512 columns): S | dP | acc1 (D columns: dS x T1 = dQ or dK) | acc2 (D columns: P x T2 = dV).  The 16-bit P
// and dS overwrite the first 64 columns of S and dP.  The dK/dV kernel runs one tile at a time (the tensor pipe idles
// while the compute warpgroup works).  The dQ kernel needs no P in TMEM and no acc2, so it keeps a second dP buffer in
// columns 384..511 and issues the score-like products of step n+1 while the compute warpgroup works on step n.
// This is the correctness baseline for the backward, not a tuned kernel.
#pragma once
#include "fa_fwd_sm100.cuh"

namespace fa {

struct BwdArgs {
  const float* lse;     // [BH, N] natural-log logsumexp of the scaled scores (forward output)
  const float* delta;   // [BH, N] dO_i . O_i (bwd_delta_kernel)
  int N, H, num_tiles;  // num_tiles = ceil(N / 128)
  float scale;          // softmax_scale
  float scale_log2;     // softmax_scale * log2(e)
  unsigned int perm;    // axis order of the (dense, identically shaped) tensor maps
  unsigned long long desc_k;    // K-major descriptor bits (score-like products)
  unsigned long long desc_mn;   // MN-major descriptor bits (accumulating products)
  unsigned int idesc_ss, idesc_ts;
};

constexpr int kBwdThreads = 256;

template <int D>
struct BwdTraits {
  using F = FwdTraits<D>;
  static constexpr int kTileBytes = F::kTileBytes;
  static constexpr int kAuxBytes = 4096;   // barriers (256 B) + column statistics 2 stages x 2 x 512 B
  static constexpr int kSmemBytes = 6 * kTileBytes + kAuxBytes;   // F1 F2 | 2 stages x (T1 T2) | aux
};

// delta[row] = sum_t dO[row,t] * O[row,t]  (fp32).  One thread per 8 elements, d/8 lanes per row.
template <bool kBF16>
__global__ void __launch_bounds__(256)
bwd_delta_kernel(const uint4* __restrict__ O, const uint4* __restrict__ dO, float* __restrict__ delta, long long rows, int d) {
  const int lanes = d / 8;   // 4, 8 or 16: a power of two, so a row never straddles a warp
  const long long total = rows * lanes;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // every thread of a warp runs the same number of iterations (total is rounded up to whole warps by the shuffles' mask)
  for (long long base = (long long)blockIdx.x * blockDim.x; base < total; base += stride) {
    const long long idx = base + threadIdx.x;
    float s = 0.f;
    if (idx < total) {
      const uint4 a = O[idx], b = dO[idx];
      const float2 a0 = unpack2<kBF16>(a.x), a1 = unpack2<kBF16>(a.y), a2 = unpack2<kBF16>(a.z), a3 = unpack2<kBF16>(a.w);
      const float2 b0 = unpack2<kBF16>(b.x), b1 = unpack2<kBF16>(b.y), b2 = unpack2<kBF16>(b.z), b3 = unpack2<kBF16>(b.w);
      s = a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
    }
    for (int off = lanes / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (idx < total && (idx % lanes) == 0) delta[idx / lanes] = s;
  }
}

template <int D, bool kBF16, bool kCausal, bool kDQ>
__global__ void __launch_bounds__(kBwdThreads, 1)
fa_bwd_sm100_kernel(const __grid_constant__ CUtensorMap tmF1, const __grid_constant__ CUtensorMap tmF2,
                    const __grid_constant__ CUtensorMap tmT1, const __grid_constant__ CUtensorMap tmT2,
                    const __grid_constant__ CUtensorMap tmOut1, const __grid_constant__ CUtensorMap tmOut2,
                    const BwdArgs a) {
  using T = FwdTraits<D>;
  constexpr uint32_t kTileBytes = T::kTileBytes;
  constexpr uint32_t kBoxBytes = T::kBoxBytes;
  constexpr int kNumBoxes = T::kNumBoxes;
  constexpr int kBoxCols = T::kBoxCols;
  constexpr uint32_t kRowBytes = T::kRowBytes;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  const uint32_t sF1 = smem_base, sF2 = smem_base + kTileBytes;
  const uint32_t sT = smem_base + 2 * kTileBytes;             // stage s: T1 at sT + 2 s tile, T2 one tile further
  const uint32_t bars = smem_base + 6 * kTileBytes;
  const uint32_t bar_f_full = bars;            //      TMA -> MMA  (fixed tiles)
  const uint32_t bar_t_full = bars + 8;        // [2]  TMA -> MMA  (streamed tiles)
  const uint32_t bar_t_empty = bars + 24;      // [2]  MMA -> TMA
  const uint32_t bar_s_full = bars + 40;       //      MMA -> compute (S and dP are in TMEM)
  const uint32_t bar_p_full = bars + 48;       //      compute -> MMA (P and dS are in TMEM; 128 arrivals)
  const uint32_t bar_acc_full = bars + 56;     //      MMA -> compute (all accumulating products have landed)
  const uint32_t bar_s_free = bars + 72;       //      compute -> MMA (dQ kernel: S is in registers; 128 arrivals)
  const uint32_t tmem_slot = bars + 64;
  const uint32_t s_stats = bars + 256;         // [2 stages][lse2 | delta][128] fp32 (dK/dV kernel only)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int f = int(blockIdx.x % uint32_t(a.num_tiles));     // fixed tile
  const int bh = int(blockIdx.x / uint32_t(a.num_tiles));
  const int b_idx = bh / a.H, h_idx = bh - b_idx * a.H;
  // streamed tiles this CTA visits
  const int t_begin = (kCausal && !kDQ) ? f : 0;
  const int t_end = (kCausal && kDQ) ? f + 1 : a.num_tiles;
  const int steps = t_end - t_begin;

  if (warp == 4 && lane == 0) {
    prefetch_tensormap(&tmF1); prefetch_tensormap(&tmF2); prefetch_tensormap(&tmT1); prefetch_tensormap(&tmT2);
    prefetch_tensormap(&tmOut1);
    if (!kDQ) prefetch_tensormap(&tmOut2);
  }
  if (warp == 5 && lane == 0) {
    mbar_init(bar_f_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_t_full + 8 * s, 1);
      mbar_init(bar_t_empty + 8 * s, 1);
    }
    mbar_init(bar_s_full, 1);
    mbar_init(bar_p_full, 128);
    mbar_init(bar_acc_full, 1);
    mbar_init(bar_s_free, 128);
    fence_mbar_init();
  }
  if (warp == 6) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  auto load_tile = [&](const CUtensorMap* tm, uint32_t dst, uint32_t bar, int row0) {
#pragma unroll
    for (int h = 0; h < kNumBoxes; ++h)
      tma_load_tile(dst + h * kBoxBytes, tm, bar, h * kBoxCols, row0, h_idx, b_idx, a.perm);
  };

  if (warp == 4) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_f_full, 2 * kTileBytes);
      load_tile(&tmF1, sF1, bar_f_full, f * kBlockM);
      load_tile(&tmF2, sF2, bar_f_full, f * kBlockM);
      for (int n = 0; n < steps; ++n) {
        const int s = n & 1;
        mbar_wait(bar_t_empty + 8 * s, (uint32_t(n >> 1) & 1u) ^ 1u, 500 + s);
        mbar_arrive_expect_tx(bar_t_full + 8 * s, 2 * kTileBytes);
        load_tile(&tmT1, sT + (2 * s) * kTileBytes, bar_t_full + 8 * s, (t_begin + n) * kBlockN);
        load_tile(&tmT2, sT + (2 * s + 1) * kTileBytes, bar_t_full + 8 * s, (t_begin + n) * kBlockN);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // =========================== MMA issuer ===========================
    const uint32_t hi_k = uint32_t(a.desc_k >> 32), lo_k = uint32_t(a.desc_k);
    const uint32_t hi_mn = uint32_t(a.desc_mn >> 32), lo_mn = uint32_t(a.desc_mn);
    const uint32_t tS = tmem_base, tdP = tmem_base + kBlockN, tA1 = tmem_base + 2 * kBlockN;
    [[maybe_unused]] const uint32_t tA2 = tA1 + D;
    // D_tmem = A (K-major smem tile) x B^T (K-major smem tile): D/16 k-steps
    auto issue_ss = [&](uint32_t d_tmem, uint32_t sa, uint32_t sb) {
      const uint32_t a_lo = lo_k | (sa >> 4), b_lo = lo_k | (sb >> 4);
#pragma unroll
      for (int k = 0; k < D / 16; ++k) {
        const uint32_t off = ((k / 4) * kBoxBytes + (k % 4) * 32) >> 4;
        umma_ss(d_tmem, a_lo + off, hi_k, b_lo + off, hi_k, a.idesc_ss, k > 0 ? 1u : 0u);
      }
    };
    // D_tmem += A (16-bit, 64 TMEM columns) x B (MN-major smem tile of 128 rows): 8 k-steps of 16 rows
    auto issue_ts = [&](uint32_t d_tmem, uint32_t a_tmem, uint32_t sb, bool acc) {
      const uint32_t b_lo = lo_mn | (sb >> 4);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_ts(d_tmem, a_tmem + k * 8, b_lo + k * (16 * kRowBytes / 16), hi_mn, a.idesc_ts, (acc || k > 0) ? 1u : 0u);
    };
    mbar_wait(bar_f_full, 0, 600);
    if constexpr (kDQ) {
      // The dQ kernel never needs P in TMEM (dS is all the accumulating product consumes), so S is free as soon as the
      // compute warpgroup has it in registers, and 384 + D <= 512 columns leave room for a second dP buffer: the
      // score-like products of step n+1 are issued while the compute warpgroup works on step n.
      auto tdP_of = [&](int n) { return tmem_base + ((n & 1) ? 3 * kBlockN : kBlockN); };
      mbar_wait(bar_t_full, 0, 610);
      tc_fence_after();
      if (elect_one_sync()) {
        issue_ss(tS, sF1, sT);                     // S(0)
        issue_ss(tdP_of(0), sF2, sT + kTileBytes); // dP(0)
        umma_commit(bar_s_full);
      }
      __syncwarp();
      for (int n = 0; n < steps; ++n) {
        const int s = n & 1;
        if (n + 1 < steps) {
          const int s1 = (n + 1) & 1;
          const uint32_t sN1 = sT + (2 * s1) * kTileBytes;
          mbar_wait(bar_t_full + 8 * s1, uint32_t((n + 1) >> 1) & 1u, 610 + s1);
          mbar_wait(bar_s_free, uint32_t(n) & 1u, 630);     // S(n) is in the compute warpgroup's registers
          tc_fence_after();
          if (elect_one_sync()) {
            issue_ss(tS, sF1, sN1);                          // S(n+1)
            issue_ss(tdP_of(n + 1), sF2, sN1 + kTileBytes);  // dP(n+1) into the other dP buffer
            umma_commit(bar_s_full);
          }
          __syncwarp();
        }
        mbar_wait(bar_p_full, uint32_t(n) & 1u, 620);
        tc_fence_after();
        if (elect_one_sync()) {
          issue_ts(tA1, tdP_of(n), sT + (2 * s) * kTileBytes, n > 0);   // dQ += dS(n) x K(n)
          umma_commit(bar_t_empty + 8 * s);
          if (n == steps - 1) umma_commit(bar_acc_full);
        }
        __syncwarp();
      }
    } else {
      for (int n = 0; n < steps; ++n) {
        const int s = n & 1;
        const uint32_t sT1 = sT + (2 * s) * kTileBytes, sT2 = sT1 + kTileBytes;
        mbar_wait(bar_t_full + 8 * s, uint32_t(n >> 1) & 1u, 610 + s);
        tc_fence_after();
        if (elect_one_sync()) {
          issue_ss(tS, sF1, sT1);     // S^T  = K Q^T
          issue_ss(tdP, sF2, sT2);    // dP^T = V dO^T
          umma_commit(bar_s_full);
        }
        __syncwarp();
        mbar_wait(bar_p_full, uint32_t(n) & 1u, 620);
        tc_fence_after();
        if (elect_one_sync()) {
          issue_ts(tA1, tdP, sT1, n > 0);     // dK += dS^T x Q
          issue_ts(tA2, tS, sT2, n > 0);      // dV += P^T  x dO
          umma_commit(bar_t_empty + 8 * s);
          if (n == steps - 1) umma_commit(bar_acc_full);
        }
        __syncwarp();
      }
    }
  } else if (warp < 4) {
    // =========================== compute warpgroup ===========================
    const int row_in_tile = warp * 32 + lane;
    const uint32_t lane_addr = uint32_t(warp * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr, tdP = tS + kBlockN, tA1 = tS + 2 * kBlockN, tA2 = tA1 + D;
    const int fixed_idx = f * kBlockM + row_in_tile;     // query (dQ kernel) or key (dK/dV kernel) of this thread
    const long long stat_base = (long long)bh * a.N;
    const float log2e = 1.4426950408889634f;
    float lse2_r = INFINITY, delta_r = 0.f;              // dQ kernel: this row's statistics
    if (kDQ && fixed_idx < a.N) {
      lse2_r = a.lse[stat_base + fixed_idx] * log2e;
      delta_r = a.delta[stat_base + fixed_idx];
    }
    for (int n = 0; n < steps; ++n) {
      const int t = t_begin + n;
      const uint32_t st = s_stats + uint32_t(n & 1) * 1024u;
      if (!kDQ) {   // column statistics of the streamed Q tile: thread r publishes those of query t*128 + r
        const int qi = t * kBlockN + row_in_tile;
        const float l2 = qi < a.N ? a.lse[stat_base + qi] * log2e : INFINITY;
        const float dl = qi < a.N ? a.delta[stat_base + qi] : 0.f;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(st + uint32_t(row_in_tile) * 4u), "f"(l2) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(st + 512u + uint32_t(row_in_tile) * 4u), "f"(dl) : "memory");
        named_bar_sync(1, 128);
      }
      // visible columns of this row: [c_lo, c_hi]
      int c_lo = 0, c_hi = kBlockN - 1;
      if (kDQ) {   // columns are keys t*128 + c: key < N and (causal) key <= query
        int last = a.N - 1;
        if (kCausal) last = min(last, fixed_idx);
        c_hi = min(c_hi, last - t * kBlockN);
        if (fixed_idx >= a.N) c_hi = -1;
      } else {     // columns are queries t*128 + c: query < N (lse2 = inf takes care of it) and (causal) query >= key
        if (kCausal) c_lo = max(0, fixed_idx - t * kBlockN);
        if (fixed_idx >= a.N) c_lo = kBlockN;
      }
      mbar_wait(bar_s_full, uint32_t(n) & 1u, 700);
      tc_fence_after();
      if constexpr (kDQ) {
        // all of S into registers first, then hand the S columns back to the MMA warp (it issues step n+1 meanwhile);
        // dP(n) lives in buffer n & 1 and is consumed in 32-column groups, dS written back over its first 64 columns
        const uint32_t tdPn = tS + ((n & 1) ? 3 * kBlockN : kBlockN);
        uint32_t sr[4][32];
#pragma unroll
        for (int q = 0; q < 4; ++q) tmem_ld32(tS + q * 32, sr[q]);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(bar_s_free);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t dr[32];
          tmem_ld32(tdPn + q * 32, dr);
          tmem_wait_ld();
          uint32_t dk[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            float dv[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = q * 32 + 2 * k + e;
              const bool vis = (c >= c_lo) && (c <= c_hi);
              const float p = vis ? ex2_approx(fmaf(__uint_as_float(sr[q][2 * k + e]), a.scale_log2, -lse2_r)) : 0.f;
              dv[e] = p * (__uint_as_float(dr[2 * k + e]) - delta_r) * a.scale;
            }
            dk[k] = pack2<kBF16>(dv[0], dv[1]);
          }
          tmem_st16(tdPn + q * 16, dk);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t sr[32], dr[32];
          tmem_ld32(tS + q * 32, sr);
          tmem_ld32(tdP + q * 32, dr);
          tmem_wait_ld();
          uint32_t pk[16], dk[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            float pv[2], dv[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = q * 32 + 2 * k + e;
              float l2, dl;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(l2) : "r"(st + uint32_t(c) * 4u));
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(dl) : "r"(st + 512u + uint32_t(c) * 4u));
              const bool vis = (c >= c_lo) && (c <= c_hi);
              const float p = vis ? ex2_approx(fmaf(__uint_as_float(sr[2 * k + e]), a.scale_log2, -l2)) : 0.f;
              pv[e] = p;
              dv[e] = p * (__uint_as_float(dr[2 * k + e]) - dl) * a.scale;
            }
            pk[k] = pack2<kBF16>(pv[0], pv[1]);
            dk[k] = pack2<kBF16>(dv[0], dv[1]);
          }
          // 16-bit P over S columns [16q, 16q+16), dS over dP columns [16q, 16q+16): both inside column groups that
          // are already in registers (groups <= q)
          tmem_st16(tS + q * 16, pk);
          tmem_st16(tdP + q * 16, dk);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(bar_p_full);
    }

    // ---- epilogue: accumulators -> 16-bit -> swizzled staging tiles (the fixed-tile buffers, dead by now) -> TMA store
    mbar_wait(bar_acc_full, 0, 710);
    tc_fence_after();
    constexpr uint32_t kChunksPerRow = kRowBytes / 16;
    const uint32_t swz = (kRowBytes == 128) ? uint32_t(row_in_tile & 7) : uint32_t((row_in_tile >> 1) & 3);
#pragma unroll
    for (int o = 0; o < (kDQ ? 1 : 2); ++o) {
      const uint32_t tA = o ? tA2 : tA1;
      const uint32_t sO = o ? sF2 : sF1;
#pragma unroll
      for (int q = 0; q < D / 32; ++q) {
        uint32_t orow[32];
        tmem_ld32(tA + q * 32, orow);
        tmem_wait_ld();
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint32_t wv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            wv[e] = pack2<kBF16>(__uint_as_float(orow[v * 8 + 2 * e]), __uint_as_float(orow[v * 8 + 2 * e + 1]));
          const uint32_t c16 = uint32_t(q * 4 + v);
          const uint32_t addr = sO + (c16 / kChunksPerRow) * kBoxBytes + row_in_tile * kRowBytes +
                                (((c16 % kChunksPerRow) ^ swz) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(wv[0]), "r"(wv[1]), "r"(wv[2]),
                       "r"(wv[3]) : "memory");
        }
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (row_in_tile == 0) {
#pragma unroll
      for (int h = 0; h < kNumBoxes; ++h) {
        tma_store_tile(&tmOut1, sF1 + h * kBoxBytes, h * kBoxCols, f * kBlockM, h_idx, b_idx, a.perm);
        if (!kDQ) tma_store_tile(&tmOut2, sF2 + h * kBoxBytes, h * kBoxCols, f * kBlockM, h_idx, b_idx, a.perm);
      }
      tma_store_commit();
      tma_store_wait<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 6) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace fa
