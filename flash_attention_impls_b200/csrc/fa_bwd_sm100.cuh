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
#include <type_traits>

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
  unsigned int idesc_ss_half;   // score-like product with N = 64 (pipelined dK/dV kernel)
};

constexpr int kBwdThreads = 256;

template <int D>
struct BwdTraits {
  using F = FwdTraits<D>;
  static constexpr int kTileBytes = F::kTileBytes;
  static constexpr int kAuxBytes = 4096;   // barriers (256 B) + column statistics 2 stages x 2 x 512 B
  static constexpr int kSmemBytes = 6 * kTileBytes + kAuxBytes;   // F1 F2 | 2 stages x (T1 T2) | aux   (round-1 kernel)
  // Round-2 kernels: the streamed tiles go through a ring of single-tile SLOTS (tile t of the stream, two per step, sits
  // in slot t % kSlots with its own full / empty barrier), as many as fit beside the two fixed tiles: 5 at d = 128
  // (7 x 32 KB + 3 KB = the 227 KB opt-in maximum), 8 below.  With 2 pair-stages the load of step n+2 could only start
  // when step n had completely finished and was needed at once: its whole latency sat on every step (ncu r02: the MMA
  // warp spins on the "tile landed" barrier, the tensor pipe is 48 % active whatever the compute warps do).  With 5
  // slots the slot of the first tile of step n+2 is already free during step n, and the second one as soon as the
  // first product that reads it has completed (dQ kernel: V_j is only read by dP = dO V^T).
  static constexpr int kAux2Bytes = 3072;  // barriers (256 B) + column statistics (2 KB)
  static constexpr int kSlotsMax = (kSmemLimit - kAux2Bytes - 2 * kTileBytes) / kTileBytes;
  static constexpr int kSlots = kSlotsMax > 8 ? 8 : kSlotsMax;
  static constexpr int kSmem2Bytes = (2 + kSlots) * kTileBytes + kAux2Bytes;
  static_assert(kSlots >= 5, "slot ring too shallow");
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
      // Tiles no row of this warp needs a mask for (everything except causal-diagonal and ragged tiles) take a path
      // without the per-element select.  The softmax scale of dS is applied once per OUTPUT element in the epilogue
      // (dQ = scale * sum dS' K, dK = scale * sum dS'^T Q with dS' = P o (dP - delta)), not per score element.
      const bool masked = !__all_sync(0xffffffffu, c_lo <= 0 && c_hi >= kBlockN - 1);
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
        auto run = [&](auto masked_c) {
          constexpr bool kMasked = decltype(masked_c)::value;
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
                float p = ex2_approx(fmaf(__uint_as_float(sr[q][2 * k + e]), a.scale_log2, -lse2_r));
                if (kMasked) p = ((c >= c_lo) && (c <= c_hi)) ? p : 0.f;
                dv[e] = p * (__uint_as_float(dr[2 * k + e]) - delta_r);
              }
              dk[k] = pack2<kBF16>(dv[0], dv[1]);
            }
            tmem_st16(tdPn + q * 16, dk);
          }
        };
        if (masked) run(std::true_type{}); else run(std::false_type{});
      } else {
        auto run = [&](auto masked_c) {
          constexpr bool kMasked = decltype(masked_c)::value;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t sr[32], dr[32];
            tmem_ld32(tS + q * 32, sr);
            tmem_ld32(tdP + q * 32, dr);
            tmem_wait_ld();
            uint32_t pk[16], dk[16];
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
              // column statistics of 4 queries with two 16-byte broadcast loads
              float l2[4], dl[4];
              const uint32_t c0 = uint32_t(q * 32 + 4 * k4);
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(l2[0]), "=f"(l2[1]), "=f"(l2[2]), "=f"(l2[3]) : "r"(st + c0 * 4u));
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(dl[0]), "=f"(dl[1]), "=f"(dl[2]), "=f"(dl[3]) : "r"(st + 512u + c0 * 4u));
              float pv[4], dv[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int c = int(c0) + e;
                float p = ex2_approx(fmaf(__uint_as_float(sr[4 * k4 + e]), a.scale_log2, -l2[e]));
                if (kMasked) p = ((c >= c_lo) && (c <= c_hi)) ? p : 0.f;
                pv[e] = p;
                dv[e] = p * (__uint_as_float(dr[4 * k4 + e]) - dl[e]);
              }
              pk[2 * k4] = pack2<kBF16>(pv[0], pv[1]);
              pk[2 * k4 + 1] = pack2<kBF16>(pv[2], pv[3]);
              dk[2 * k4] = pack2<kBF16>(dv[0], dv[1]);
              dk[2 * k4 + 1] = pack2<kBF16>(dv[2], dv[3]);
            }
            // 16-bit P over S columns [16q, 16q+16), dS over dP columns [16q, 16q+16): both inside column groups that
            // are already in registers (groups <= q)
            tmem_st16(tS + q * 16, pk);
            tmem_st16(tdP + q * 16, dk);
          }
        };
        if (masked) run(std::true_type{}); else run(std::false_type{});
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
      const float osc = o ? 1.f : a.scale;     // acc1 (dQ or dK) carries dS' = dS / scale, see the compute loop
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
            wv[e] = pack2<kBF16>(__uint_as_float(orow[v * 8 + 2 * e]) * osc, __uint_as_float(orow[v * 8 + 2 * e + 1]) * osc);
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

// ---------------------------------------------------------------------------------------------------------------------
// dK/dV kernel, pipelined version (round 2).  Same mathematics and operand modes as fa_bwd_sm100_kernel<kDQ = false>,
// but the tensor pipe no longer idles while the compute threads work: every streamed 128-query tile is processed as
// two 64-query HALVES with their own TMEM buffers and their own compute warpgroup,
//     TMEM:  S_a | dP_a | S_b | dP_b  (64 columns each)  | dK (D) | dV (D)        = 256 + 2 D <= 512 columns
//     warps 0-3: warpgroup A (queries 0..63 of every tile)     warps 4-7: warpgroup B (queries 64..127)
//     warp 8: TMA producer     warp 9: tcgen05.mma issuer     warp 10: TMEM allocator
// and the MMA warp issues, in this order,   acc_a(n)  scores_a(n+1)  acc_b(n)  scores_b(n+1)   where
//     scores_h(n): S_h^T = K_j Q_i[h]^T, dP_h^T = V_j dO_i[h]^T      (SS, M = 128 keys, N = 64 queries)
//     acc_h(n):    dK_j += dS_h^T Q_i[h],  dV_j += P_h^T dO_i[h]     (TS, A = 16-bit P_h^T / dS_h^T in TMEM, K = 64)
// so warpgroup A exponentiates half a of tile n+1 while the pipe runs acc_b(n) + scores_b(n+1), and vice versa - the
// forward kernel's two-tile ping-pong, applied to the two halves of one tile.  The in-order tensor pipe makes the
// overwrite of S_h / dP_h by scores_h(n+1) safe after acc_h(n) has consumed P_h / dS_h.  Two warps per SMSP also hide
// each other's MUFU / FMA latencies.  The N = 64 score products are shared-memory-bound in SS mode (48 instead of 32
// clk per instruction), so a tile costs 2560 tensor-pipe clk instead of 2048 - against 4820 clk per tile for the
// one-tile-at-a-time kernel (MMA 2048 + compute 2770, serialised).
constexpr int kBwdDkdvThreads = 384;

template <int D, bool kBF16, bool kCausal>
__global__ void __launch_bounds__(kBwdDkdvThreads, 1)
fa_bwd_dkdv_sm100_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                         const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                         const __grid_constant__ CUtensorMap tmdK, const __grid_constant__ CUtensorMap tmdV,
                         const BwdArgs a) {
  using T = FwdTraits<D>;
  constexpr uint32_t kTileBytes = T::kTileBytes;
  constexpr uint32_t kBoxBytes = T::kBoxBytes;
  constexpr int kNumBoxes = T::kNumBoxes;
  constexpr int kBoxCols = T::kBoxCols;
  constexpr uint32_t kRowBytes = T::kRowBytes;
  constexpr int kHalf = kBlockN / 2;     // 64 queries per half

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  const uint32_t sK = smem_base, sV = smem_base + kTileBytes;
  constexpr int kSlots = BwdTraits<D>::kSlots;
  const uint32_t sT = smem_base + 2 * kTileBytes;             // slot ring: stream tile t (Q_i: t = 2 n, dO_i: t = 2 n + 1)
  const uint32_t bars = smem_base + (2 + kSlots) * kTileBytes;
  const uint32_t bar_f_full = bars;            //      TMA -> MMA  (K_j, V_j)
  const uint32_t bar_t_full = bars + 8;        // [kSlots]  TMA -> MMA  (a streamed tile has landed)
  const uint32_t bar_t_empty = bars + 72;      // [kSlots]  MMA -> TMA  (every product reading the slot has completed)
  const uint32_t bar_s_full = bars + 136;      // [2 halves]  MMA -> compute (S_h and dP_h are in TMEM)
  const uint32_t bar_p_full = bars + 152;      // [2 halves]  compute -> MMA (P_h and dS_h are in TMEM; 128 arrivals)
  const uint32_t bar_acc_full = bars + 168;    //      MMA -> compute (all accumulating products have landed)
  const uint32_t tmem_slot = bars + 176;
  auto slot_of = [&](int t) { return uint32_t(t % kSlots); };
  auto phase_of = [&](int t) { return uint32_t((t / kSlots) & 1); };
  const uint32_t s_stats = bars + 256;         // [2 halves][2 stages][lse2(64) | delta(64)] fp32 = 2 KB

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int f = int(blockIdx.x % uint32_t(a.num_tiles));     // this CTA's K/V tile
  const int bh = int(blockIdx.x / uint32_t(a.num_tiles));
  const int b_idx = bh / a.H, h_idx = bh - b_idx * a.H;
  const int t_begin = kCausal ? f : 0;                       // causal: only query tiles at or below the diagonal
  const int steps = a.num_tiles - t_begin;

  if (warp == 8 && lane == 0) {
    prefetch_tensormap(&tmK); prefetch_tensormap(&tmV); prefetch_tensormap(&tmQ); prefetch_tensormap(&tmdO);
    prefetch_tensormap(&tmdK); prefetch_tensormap(&tmdV);
  }
  if (warp == 9 && lane == 0) {
    mbar_init(bar_f_full, 1);
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(bar_t_full + 8 * s, 1);
      mbar_init(bar_t_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_s_full + 8 * s, 1);
      mbar_init(bar_p_full + 8 * s, 128);
    }
    mbar_init(bar_acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 10) {
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

  if (warp == 8) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_f_full, 2 * kTileBytes);
      load_tile(&tmK, sK, bar_f_full, f * kBlockM);
      load_tile(&tmV, sV, bar_f_full, f * kBlockM);
      for (int t = 0; t < 2 * steps; ++t) {
        const uint32_t s = slot_of(t);
        mbar_wait(bar_t_empty + 8 * s, phase_of(t) ^ 1u, 500);
        mbar_arrive_expect_tx(bar_t_full + 8 * s, kTileBytes);
        load_tile((t & 1) ? &tmdO : &tmQ, sT + s * kTileBytes, bar_t_full + 8 * s, (t_begin + (t >> 1)) * kBlockN);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // =========================== MMA issuer ===========================
    const uint32_t hi_k = uint32_t(a.desc_k >> 32), lo_k = uint32_t(a.desc_k);
    const uint32_t hi_mn = uint32_t(a.desc_mn >> 32), lo_mn = uint32_t(a.desc_mn);
    const uint32_t tdK = tmem_base + 2 * kBlockN, tdV = tdK + D;
    const uint32_t idesc_half = a.idesc_ss_half;
    // S_h^T = K_j x Q_i[h]^T and dP_h^T = V_j x dO_i[h]^T: B operand = rows [64 h, 64 h + 64) of the K-major streamed tile
    auto issue_scores = [&](int h, uint32_t sQ, uint32_t sdO) {
      const uint32_t tS = tmem_base + h * kBlockN, tdP = tS + kHalf;
      const uint32_t half_off = (uint32_t(h) * kHalf * kRowBytes) >> 4;
      const uint32_t aK = lo_k | (sK >> 4), aV = lo_k | (sV >> 4);
      const uint32_t bQ = (lo_k | (sQ >> 4)) + half_off, bdO = (lo_k | (sdO >> 4)) + half_off;
#pragma unroll
      for (int k = 0; k < D / 16; ++k) {
        const uint32_t off = ((k / 4) * kBoxBytes + (k % 4) * 32) >> 4;
        umma_ss(tS, aK + off, hi_k, bQ + off, hi_k, idesc_half, k > 0 ? 1u : 0u);
      }
#pragma unroll
      for (int k = 0; k < D / 16; ++k) {
        const uint32_t off = ((k / 4) * kBoxBytes + (k % 4) * 32) >> 4;
        umma_ss(tdP, aV + off, hi_k, bdO + off, hi_k, idesc_half, k > 0 ? 1u : 0u);
      }
    };
    // dK += dS_h^T x Q_i[h], dV += P_h^T x dO_i[h]: A = 32 TMEM columns (64 16-bit queries), 4 k-steps of 16 queries;
    // B = rows [64 h + 16 k, +16) of the MN-major streamed tile
    auto issue_acc = [&](int h, uint32_t sQ, uint32_t sdO, bool acc) {
      const uint32_t tP = tmem_base + h * kBlockN, tdS = tP + kHalf;
      const uint32_t bQ = lo_mn | (sQ >> 4), bdO = lo_mn | (sdO >> 4);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_ts(tdK, tdS + k * 8, bQ + (4 * h + k) * (16 * kRowBytes / 16), hi_mn, a.idesc_ts, (acc || k > 0) ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_ts(tdV, tP + k * 8, bdO + (4 * h + k) * (16 * kRowBytes / 16), hi_mn, a.idesc_ts, (acc || k > 0) ? 1u : 0u);
    };
    mbar_wait(bar_f_full, 0, 600);
    mbar_wait(bar_t_full + 8 * slot_of(0), phase_of(0), 610);
    mbar_wait(bar_t_full + 8 * slot_of(1), phase_of(1), 611);
    tc_fence_after();
    if (elect_one_sync()) {
      issue_scores(0, sT + slot_of(0) * kTileBytes, sT + slot_of(1) * kTileBytes);
      umma_commit(bar_s_full);
      issue_scores(1, sT + slot_of(0) * kTileBytes, sT + slot_of(1) * kTileBytes);
      umma_commit(bar_s_full + 8);
    }
    __syncwarp();
    for (int n = 0; n < steps; ++n) {
      const uint32_t sQ = sT + slot_of(2 * n) * kTileBytes, sdO = sT + slot_of(2 * n + 1) * kTileBytes;
      const uint32_t sQ1 = sT + slot_of(2 * n + 2) * kTileBytes, sdO1 = sT + slot_of(2 * n + 3) * kTileBytes;
      const bool more = n + 1 < steps;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        mbar_wait(bar_p_full + 8 * h, uint32_t(n) & 1u, 620 + h);
        if (h == 0 && more) {
          mbar_wait(bar_t_full + 8 * slot_of(2 * n + 2), phase_of(2 * n + 2), 612);
          mbar_wait(bar_t_full + 8 * slot_of(2 * n + 3), phase_of(2 * n + 3), 613);
        }
        tc_fence_after();
        if (elect_one_sync()) {
          issue_acc(h, sQ, sdO, n > 0 || h > 0);
          if (h == 1) {                                          // every read of Q_i and dO_i has been issued
            umma_commit(bar_t_empty + 8 * slot_of(2 * n));
            umma_commit(bar_t_empty + 8 * slot_of(2 * n + 1));
          }
          if (more) {
            issue_scores(h, sQ1, sdO1);
            umma_commit(bar_s_full + 8 * h);
          } else if (h == 1) {
            umma_commit(bar_acc_full);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp < 8) {
    // =========================== compute warpgroups (A: half 0, B: half 1) ===========================
    const int h = warp >> 2;
    const int wl = warp & 3;
    const int row_in_tile = wl * 32 + lane;
    const uint32_t lane_addr = uint32_t(wl * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + h * kBlockN, tdP = tS + kHalf;
    const int key_idx = f * kBlockM + row_in_tile;
    const long long stat_base = (long long)bh * a.N;
    const float log2e = 1.4426950408889634f;
    // column statistics of this half's 64 queries of tile t: threads 0..63 fetch lse * log2e, threads 64..127 delta.
    // The global load for step n+1 is issued during step n, so its latency is not on the step's critical chain.
    auto fetch_stat = [&](int t) -> float {
      const int qi = t * kBlockN + h * kHalf + (row_in_tile & (kHalf - 1));
      if (row_in_tile < kHalf) return qi < a.N ? a.lse[stat_base + qi] * log2e : INFINITY;
      return qi < a.N ? a.delta[stat_base + qi] : 0.f;
    };
    float stat_next = fetch_stat(t_begin);
    for (int n = 0; n < steps; ++n) {
      const int t = t_begin + n;
      const uint32_t st = s_stats + uint32_t(h) * 1024u + uint32_t(n & 1) * 512u;
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(st + uint32_t(row_in_tile) * 4u), "f"(stat_next) : "memory");
      named_bar_sync(1 + h, 128);
      if (n + 1 < steps) stat_next = fetch_stat(t + 1);
      // visible query columns of this key row inside the half: c >= c_lo (causal: query >= key); keys beyond N see nothing
      int c_lo = 0;
      if (kCausal) c_lo = max(0, key_idx - t * kBlockN - h * kHalf);
      if (key_idx >= a.N) c_lo = kHalf;
      const bool masked = __any_sync(0xffffffffu, c_lo > 0);
      mbar_wait(bar_s_full + 8 * h, uint32_t(n) & 1u, 700 + h);
      tc_fence_after();
      auto run = [&](auto masked_c) {
        constexpr bool kMasked = decltype(masked_c)::value;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t sr[32], dr[32];
          tmem_ld32(tS + g * 32, sr);
          tmem_ld32(tdP + g * 32, dr);
          tmem_wait_ld();
          uint32_t pk[16], dk[16];
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            float l2[4], dl[4];
            const uint32_t c0 = uint32_t(g * 32 + 4 * k4);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(l2[0]), "=f"(l2[1]), "=f"(l2[2]), "=f"(l2[3]) : "r"(st + c0 * 4u));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(dl[0]), "=f"(dl[1]), "=f"(dl[2]), "=f"(dl[3]) : "r"(st + 256u + c0 * 4u));
            float pv[4], dv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float p = ex2_approx(fmaf(__uint_as_float(sr[4 * k4 + e]), a.scale_log2, -l2[e]));
              if (kMasked) p = (int(c0) + e >= c_lo) ? p : 0.f;
              pv[e] = p;
              dv[e] = p * (__uint_as_float(dr[4 * k4 + e]) - dl[e]);     // dS / scale; the scale is applied in the epilogue
            }
            pk[2 * k4] = pack2<kBF16>(pv[0], pv[1]);
            pk[2 * k4 + 1] = pack2<kBF16>(pv[2], pv[3]);
            dk[2 * k4] = pack2<kBF16>(dv[0], dv[1]);
            dk[2 * k4 + 1] = pack2<kBF16>(dv[2], dv[3]);
          }
          // 16-bit P_h^T over S_h columns [16 g, 16 g + 16), dS_h^T over dP_h columns [16 g, 16 g + 16): inside the
          // column groups that are already in registers
          tmem_st16(tS + g * 16, pk);
          tmem_st16(tdP + g * 16, dk);
        }
      };
      if (masked) run(std::true_type{}); else run(std::false_type{});
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(bar_p_full + 8 * h);
    }

    // ---- epilogue: warpgroup A stores dK (scaled), warpgroup B stores dV, through the K_j / V_j buffers (dead by now)
    mbar_wait(bar_acc_full, 0, 710);
    tc_fence_after();
    constexpr uint32_t kChunksPerRow = kRowBytes / 16;
    const uint32_t swz = (kRowBytes == 128) ? uint32_t(row_in_tile & 7) : uint32_t((row_in_tile >> 1) & 3);
    const uint32_t tA = tmem_base + lane_addr + 2 * kBlockN + h * D;
    const uint32_t sO = h ? sV : sK;
    const float osc = h ? 1.f : a.scale;
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
          wv[e] = pack2<kBF16>(__uint_as_float(orow[v * 8 + 2 * e]) * osc, __uint_as_float(orow[v * 8 + 2 * e + 1]) * osc);
        const uint32_t c16 = uint32_t(q * 4 + v);
        const uint32_t addr = sO + (c16 / kChunksPerRow) * kBoxBytes + row_in_tile * kRowBytes +
                              (((c16 % kChunksPerRow) ^ swz) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(wv[0]), "r"(wv[1]), "r"(wv[2]),
                     "r"(wv[3]) : "memory");
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(1 + h, 128);
    if (row_in_tile == 0) {
#pragma unroll
      for (int bx = 0; bx < kNumBoxes; ++bx)
        tma_store_tile(h ? &tmdV : &tmdK, sO + bx * kBoxBytes, bx * kBoxCols, f * kBlockM, h_idx, b_idx, a.perm);
      tma_store_commit();
      tma_store_wait<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// dQ kernel with two compute warpgroups (round 2).  Same pipeline as fa_bwd_sm100_kernel<kDQ = true> (S released as soon
// as it is in registers, two dP buffers, the score-like products of step n+1 issued while step n is being
// exponentiated), but the 128 key columns of every score row are split between two warpgroups:
//     warps 0-3: keys 0..63 of the tile        warps 4-7: keys 64..127        (one thread per query row in both)
// The kernel was bound by its single compute warpgroup (3390 clk per step against 1536 clk of MMAs, tensor pipe 47 %
// active, ncu r02); two warps per SMSP halve the per-thread work and hide each other's MUFU latency.  No exchange is
// needed - the row statistics are inputs.  dS (16-bit) of keys 0..63 overwrites dP columns 0..31 and that of keys
// 64..127 dP columns 64..95: each warpgroup only overwrites fp32 columns it has itself already loaded, and the TS
// product takes its A operand per k-step (8 columns), so the two halves need not be adjacent.
template <int D, bool kBF16, bool kCausal>
__global__ void __launch_bounds__(kBwdDkdvThreads, 1)
fa_bwd_dq_sm100_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                       const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                       const __grid_constant__ CUtensorMap tmdQ, const BwdArgs a) {
  using T = FwdTraits<D>;
  constexpr uint32_t kTileBytes = T::kTileBytes;
  constexpr uint32_t kBoxBytes = T::kBoxBytes;
  constexpr int kNumBoxes = T::kNumBoxes;
  constexpr int kBoxCols = T::kBoxCols;
  constexpr uint32_t kRowBytes = T::kRowBytes;
  constexpr int kHalf = kBlockN / 2;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  const uint32_t sQ = smem_base, sdO = smem_base + kTileBytes;
  constexpr int kSlots = BwdTraits<D>::kSlots;
  const uint32_t sT = smem_base + 2 * kTileBytes;             // slot ring: stream tile t (V_j: t = 2 n, K_j: t = 2 n + 1)
  const uint32_t bars = smem_base + (2 + kSlots) * kTileBytes;
  const uint32_t bar_f_full = bars;            //      TMA -> MMA  (Q_i, dO_i)
  const uint32_t bar_t_full = bars + 8;        // [kSlots]  TMA -> MMA  (a streamed tile has landed)
  const uint32_t bar_t_empty = bars + 72;      // [kSlots]  MMA -> TMA  (every product reading the slot has completed)
  const uint32_t bar_s_full = bars + 136;      //      MMA -> compute (S and dP are in TMEM)
  const uint32_t bar_p_full = bars + 144;      //      compute -> MMA (dS is in TMEM; 256 arrivals)
  const uint32_t bar_acc_full = bars + 152;    //      MMA -> compute (dQ complete)
  const uint32_t bar_s_free = bars + 160;      //      compute -> MMA (S is in registers; 256 arrivals)
  const uint32_t tmem_slot = bars + 176;
  // V_j goes first in the stream: its slot is released as soon as dP = dO V^T has completed (early in the step), and
  // with 5 slots that is the slot K_{j+2} lands in, while V_{j+2} takes the slot K_{j-1} gave up a step earlier
  auto slot_of = [&](int t) { return uint32_t(t % kSlots); };
  auto phase_of = [&](int t) { return uint32_t((t / kSlots) & 1); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int f = int(blockIdx.x % uint32_t(a.num_tiles));     // this CTA's query tile
  const int bh = int(blockIdx.x / uint32_t(a.num_tiles));
  const int b_idx = bh / a.H, h_idx = bh - b_idx * a.H;
  const int steps = kCausal ? f + 1 : a.num_tiles;           // causal: key tiles up to the diagonal

  if (warp == 8 && lane == 0) {
    prefetch_tensormap(&tmQ); prefetch_tensormap(&tmdO); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV);
    prefetch_tensormap(&tmdQ);
  }
  if (warp == 9 && lane == 0) {
    mbar_init(bar_f_full, 1);
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(bar_t_full + 8 * s, 1);
      mbar_init(bar_t_empty + 8 * s, 1);
    }
    mbar_init(bar_s_full, 1);
    mbar_init(bar_p_full, 256);
    mbar_init(bar_acc_full, 1);
    mbar_init(bar_s_free, 256);
    fence_mbar_init();
  }
  if (warp == 10) {
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
  // dP(n) lives in TMEM columns 128.. (even n) or 384.. (odd n); S in columns 0..127; dQ in 256..256+D
  auto dp_col = [&](int n) { return uint32_t((n & 1) ? 3 * kBlockN : kBlockN); };

  if (warp == 8) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_f_full, 2 * kTileBytes);
      load_tile(&tmQ, sQ, bar_f_full, f * kBlockM);
      load_tile(&tmdO, sdO, bar_f_full, f * kBlockM);
      for (int t = 0; t < 2 * steps; ++t) {
        const uint32_t s = slot_of(t);
        mbar_wait(bar_t_empty + 8 * s, phase_of(t) ^ 1u, 500);
        mbar_arrive_expect_tx(bar_t_full + 8 * s, kTileBytes);
        load_tile((t & 1) ? &tmK : &tmV, sT + s * kTileBytes, bar_t_full + 8 * s, (t >> 1) * kBlockN);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // =========================== MMA issuer ===========================
    const uint32_t hi_k = uint32_t(a.desc_k >> 32), lo_k = uint32_t(a.desc_k);
    const uint32_t hi_mn = uint32_t(a.desc_mn >> 32), lo_mn = uint32_t(a.desc_mn);
    const uint32_t tS = tmem_base, tdQ = tmem_base + 2 * kBlockN;
    auto issue_ss = [&](uint32_t d_tmem, uint32_t sa, uint32_t sb) {
      const uint32_t a_lo = lo_k | (sa >> 4), b_lo = lo_k | (sb >> 4);
#pragma unroll
      for (int k = 0; k < D / 16; ++k) {
        const uint32_t off = ((k / 4) * kBoxBytes + (k % 4) * 32) >> 4;
        umma_ss(d_tmem, a_lo + off, hi_k, b_lo + off, hi_k, a.idesc_ss, k > 0 ? 1u : 0u);
      }
    };
    // S(n) = Q_i K_n^T, dP(n) = dO_i V_n^T; the V slot is released by dP's own completion
    auto issue_scores = [&](int n) {
      const uint32_t sV = sT + slot_of(2 * n) * kTileBytes, sK = sT + slot_of(2 * n + 1) * kTileBytes;
      issue_ss(tmem_base + dp_col(n), sdO, sV);
      umma_commit(bar_t_empty + 8 * slot_of(2 * n));
      issue_ss(tS, sQ, sK);
      umma_commit(bar_s_full);
    };
    mbar_wait(bar_f_full, 0, 600);
    mbar_wait(bar_t_full + 8 * slot_of(0), phase_of(0), 610);
    mbar_wait(bar_t_full + 8 * slot_of(1), phase_of(1), 611);
    tc_fence_after();
    if (elect_one_sync()) issue_scores(0);
    __syncwarp();
    for (int n = 0; n < steps; ++n) {
      if (n + 1 < steps) {
        mbar_wait(bar_t_full + 8 * slot_of(2 * n + 2), phase_of(2 * n + 2), 612);
        mbar_wait(bar_t_full + 8 * slot_of(2 * n + 3), phase_of(2 * n + 3), 613);
        mbar_wait(bar_s_free, uint32_t(n) & 1u, 630);          // S(n) is in the compute warpgroups' registers
        tc_fence_after();
        if (elect_one_sync()) issue_scores(n + 1);             // dP(n+1) into the other dP buffer, S(n+1) over S(n)
        __syncwarp();
      }
      mbar_wait(bar_p_full, uint32_t(n) & 1u, 620);
      tc_fence_after();
      if (elect_one_sync()) {
        // dQ += dS(n) x K(n): k-steps 0..3 take keys 0..63 from dP columns 0..31, k-steps 4..7 keys 64..127 from 64..95
        const uint32_t b_lo = lo_mn | ((sT + slot_of(2 * n + 1) * kTileBytes) >> 4);
        const uint32_t tdS = tmem_base + dp_col(n);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ts(tdQ, tdS + (k / 4) * kHalf + (k % 4) * 8, b_lo + k * (16 * kRowBytes / 16), hi_mn, a.idesc_ts,
                  (n > 0 || k > 0) ? 1u : 0u);
        umma_commit(bar_t_empty + 8 * slot_of(2 * n + 1));
        if (n == steps - 1) umma_commit(bar_acc_full);
      }
      __syncwarp();
    }
  } else if (warp < 8) {
    // =========================== compute warpgroups (keys 0..63 / 64..127 of every tile) ===========================
    const int h = warp >> 2;
    const int wl = warp & 3;
    const int row_in_tile = wl * 32 + lane;
    const uint32_t lane_addr = uint32_t(wl * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + h * kHalf;
    const int q_idx = f * kBlockM + row_in_tile;
    const long long stat_base = (long long)bh * a.N;
    float lse2_r = INFINITY, delta_r = 0.f;
    if (q_idx < a.N) {
      lse2_r = a.lse[stat_base + q_idx] * 1.4426950408889634f;
      delta_r = a.delta[stat_base + q_idx];
    }
    for (int n = 0; n < steps; ++n) {
      // visible key columns of this row inside the half: c <= c_hi (key < N, causal: key <= query)
      int last = a.N - 1;
      if (kCausal) last = min(last, q_idx);
      int c_hi = min(kHalf - 1, last - n * kBlockN - h * kHalf);
      if (q_idx >= a.N) c_hi = -1;
      const bool masked = __any_sync(0xffffffffu, c_hi < kHalf - 1);
      const uint32_t tdP = tmem_base + lane_addr + dp_col(n) + h * kHalf;
      mbar_wait(bar_s_full, uint32_t(n) & 1u, 700 + h);
      tc_fence_after();
      uint32_t sr[2][32];
      tmem_ld32(tS, sr[0]);
      tmem_ld32(tS + 32, sr[1]);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(bar_s_free);
      auto run = [&](auto masked_c) {
        constexpr bool kMasked = decltype(masked_c)::value;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t dr[32];
          tmem_ld32(tdP + g * 32, dr);
          tmem_wait_ld();
          uint32_t dk[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            float dv[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              float p = ex2_approx(fmaf(__uint_as_float(sr[g][2 * k + e]), a.scale_log2, -lse2_r));
              if (kMasked) p = (g * 32 + 2 * k + e <= c_hi) ? p : 0.f;
              dv[e] = p * (__uint_as_float(dr[2 * k + e]) - delta_r);     // dS / scale; the scale is applied in the epilogue
            }
            dk[k] = pack2<kBF16>(dv[0], dv[1]);
          }
          tmem_st16(tdP + g * 16, dk);     // over fp32 columns this thread has already loaded
        }
      };
      if (masked) run(std::true_type{}); else run(std::false_type{});
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(bar_p_full);
    }

    // ---- epilogue: dQ * scale -> 16-bit -> swizzled staging (the Q_i buffer, dead by now) -> TMA store;
    //      warpgroup A converts columns [0, D/2), warpgroup B [D/2, D)
    mbar_wait(bar_acc_full, 0, 710);
    tc_fence_after();
    constexpr uint32_t kChunksPerRow = kRowBytes / 16;
    const uint32_t swz = (kRowBytes == 128) ? uint32_t(row_in_tile & 7) : uint32_t((row_in_tile >> 1) & 3);
    const uint32_t tA = tmem_base + lane_addr + 2 * kBlockN;
    constexpr int kGroups = D / 32;                       // 32-column groups of the accumulator: 1, 2 or 4
    constexpr int kPerWg = kGroups > 1 ? kGroups / 2 : 1;
    if (kGroups > 1 || h == 0) {
#pragma unroll
      for (int qq = 0; qq < kPerWg; ++qq) {
        const int q = (kGroups > 1 ? h * kPerWg : 0) + qq;
        uint32_t orow[32];
        tmem_ld32(tA + q * 32, orow);
        tmem_wait_ld();
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint32_t wv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            wv[e] = pack2<kBF16>(__uint_as_float(orow[v * 8 + 2 * e]) * a.scale, __uint_as_float(orow[v * 8 + 2 * e + 1]) * a.scale);
          const uint32_t c16 = uint32_t(q * 4 + v);
          const uint32_t addr = sQ + (c16 / kChunksPerRow) * kBoxBytes + row_in_tile * kRowBytes +
                                (((c16 % kChunksPerRow) ^ swz) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(wv[0]), "r"(wv[1]), "r"(wv[2]),
                       "r"(wv[3]) : "memory");
        }
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 256);
    if (warp == 0 && lane == 0) {
#pragma unroll
      for (int bx = 0; bx < kNumBoxes; ++bx)
        tma_store_tile(&tmdQ, sQ + bx * kBoxBytes, bx * kBoxCols, f * kBlockM, h_idx, b_idx, a.perm);
      tma_store_commit();
      tma_store_wait<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace fa
