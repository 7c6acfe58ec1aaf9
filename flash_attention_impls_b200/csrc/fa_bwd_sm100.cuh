// fa_bwd_sm100.cuh — attention backward for sm_100a (B200): dQ, dK, dV from dO and the forward's logsumexp.
//
// the only other algorithm of the reference is the Triton backward
//   code/triton_fa2/FA2-triton.py:98-170 (_bwd_kernel: recompute P from the saved row statistics, dV += P^T dO,
//   dP = dO V^T, dS, dQ += dS K, dK += dS^T Q with fp16 atomic adds), :207-237 (_FlashAttnFn.backward, which passes
//   every tensor's strides).
// Here it is two atomics-free tcgen05 kernels built from the operand modes the forward kernel proved on hardware
// (SS K-major for the score-like products, TS with an MN-major shared-memory B operand for the products that consume
// the 16-bit probabilities from TMEM), plus an HBM-bound pre-pass delta_i = dO_i . O_i:
//
//   dQ kernel    (fa_bwd_dq_sm100_kernel):   a CTA owns a 128-row Q tile (Q_i, dO_i resident) and streams K_j, V_j:
//        S = Q_i K_j^T,  dP = dO_i V_j^T,  P = exp(S*scale - lse_i),  dS' = P o (dP - delta_i),  dQ_i += dS' K_j
//   dK/dV kernel (fa_bwd_dkdv_sm100_kernel): a CTA owns a 128-row K/V tile (K_j, V_j resident) and streams Q_i, dO_i;
//        the score tile is computed TRANSPOSED so that keys are the TMEM lanes and P^T, dS'^T are TMEM A operands as
//        they stand:  S^T = K_j Q_i^T,  dP^T = V_j dO_i^T,  dV_j += P^T dO_i,  dK_j += dS'^T Q_i
//   The softmax scale of dS is applied once per output element in the epilogues.  The dQ kernel recomputes S and dP
//   (7 tile products in total instead of 5) and in exchange nothing is accumulated through global memory: every output
//   tile has exactly one writer, results are deterministic.
// Both kernels: 384 threads - two compute warpgroups (one thread per TMEM lane), a TMA producer warp, a tcgen05.mma
// issuer warp, a TMEM allocator warp - and a ring of single-tile shared-memory slots for the streamed tiles.
// Tensors are [B,H,N,d] / [B,H,N_kv,d] views with free batch / head / row strides (4-D tensor maps, as in the forward);
// the causal mask is the forward's: key col is masked for query row when col > row + (N_kv - N).
#pragma once
#include <type_traits>

#include "fa_fwd_sm100.cuh"

namespace fa {

struct BwdArgs {
  const float* lse;     // [BH, Nq] natural-log logsumexp of the scaled scores (forward output; -inf: row saw no key)
  const float* delta;   // [BH, Nq] dO_i . O_i (bwd_delta_kernel)
  int Nq, Nkv, H;
  int causal_off;       // Nkv - Nq
  int q_tiles, kv_tiles;   // ceil(Nq / 128), ceil(Nkv / 128)
  float scale;          // softmax_scale
  float scale_log2;     // softmax_scale * log2(e)
  // axis order of the tensor maps (see make_tmap): Q, dO, K|V, dQ, dK|dV
  unsigned int perm_q, perm_do, perm_kv, perm_dq, perm_dkv;
  unsigned long long desc_k;    // K-major descriptor bits (score-like products)
  unsigned long long desc_mn;   // MN-major descriptor bits (accumulating products)
  unsigned int idesc_ss, idesc_ts;
  unsigned int idesc_ss_half;   // score-like product with N = 64 (dK/dV kernel)
};

template <int D>
struct BwdTraits {
  using F = FwdTraits<D>;
  static constexpr int kTileBytes = F::kTileBytes;
  // The streamed tiles go through a ring of single-tile SLOTS (tile t of the stream, two per step, sits
  // in slot t % kSlots with its own full / empty barrier), as many as fit beside the two fixed tiles: 5 at d = 128
  // (7 x 32 KB + 3 KB = the 227 KB opt-in maximum), 8 below.  With the 2 pair-stages of round 1 the load of step n+2
  // could only start when step n had completely finished and was needed at once: its whole latency sat on every step
  // (ncu r02: the MMA warp spins on the "tile landed" barrier, tensor pipe 48 % active whatever the compute warps do).  With 5
  // slots the slot of the first tile of step n+2 is already free during step n, and the second one as soon as the
  // first product that reads it has completed (dQ kernel: V_j is only read by dP = dO V^T).
  static constexpr int kAux2Bytes = 3072;  // barriers (256 B) + column statistics (2 KB)
  static constexpr int kSlotsMax = (kSmemLimit - kAux2Bytes - 2 * kTileBytes) / kTileBytes;
  static constexpr int kSlots = kSlotsMax > 8 ? 8 : kSlotsMax;
  static constexpr int kSmem2Bytes = (2 + kSlots) * kTileBytes + kAux2Bytes;
  static_assert(kSlots >= 5, "slot ring too shallow");
};

// delta[row] = sum_t dO[row,t] * O[row,t]  (fp32), row = (b*H + h)*N + n.  One thread per 8 elements, d/8 lanes per
// row; O and dO are [B,H,N,d] views with element strides (sb, sh, sn); kDense: both are dense, so the row / head /
// batch decomposition (64-bit divisions) is skipped and the tensors are read as plain 16-byte streams.
template <bool kBF16, bool kDense>
__global__ void __launch_bounds__(256)
bwd_delta_kernel(const char* __restrict__ O, const char* __restrict__ dO, float* __restrict__ delta, long long rows, int d,
                 int H, int N, long long o_sb, long long o_sh, long long o_sn, long long do_sb, long long do_sh,
                 long long do_sn) {
  const int lanes = d / 8;   // 4, 8 or 16: a power of two, so a row never straddles a warp
  const long long total = rows * lanes;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // every thread of a warp runs the same number of iterations (total is rounded up to whole warps by the shuffles' mask)
  for (long long base = (long long)blockIdx.x * blockDim.x; base < total; base += stride) {
    const long long idx = base + threadIdx.x;
    float s = 0.f;
    if (idx < total) {
      uint4 a, g;
      if constexpr (kDense) {
        a = reinterpret_cast<const uint4*>(O)[idx];
        g = reinterpret_cast<const uint4*>(dO)[idx];
      } else {
        const long long row = idx / lanes;
        const int v = int(idx - row * lanes);
        const long long bh = row / N, n = row - bh * N;
        const long long b = bh / H, h = bh - b * H;
        a = *reinterpret_cast<const uint4*>(O + (b * o_sb + h * o_sh + n * o_sn + v * 8) * 2);
        g = *reinterpret_cast<const uint4*>(dO + (b * do_sb + h * do_sh + n * do_sn + v * 8) * 2);
      }
      const float2 a0 = unpack2<kBF16>(a.x), a1 = unpack2<kBF16>(a.y), a2 = unpack2<kBF16>(a.z), a3 = unpack2<kBF16>(a.w);
      const float2 b0 = unpack2<kBF16>(g.x), b1 = unpack2<kBF16>(g.y), b2 = unpack2<kBF16>(g.z), b3 = unpack2<kBF16>(g.w);
      s = a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
    }
    for (int off = lanes / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (idx < total && (idx % lanes) == 0) delta[idx / lanes] = s;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// dK/dV kernel.  Round 1 ran one streamed tile at a time (score products -> compute -> accumulating products, tensor
// pipe idle while the compute threads worked); here every streamed 128-query tile is processed as
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
constexpr int kBwdThreads = 384;

template <int D, bool kBF16, bool kCausal>
__global__ void __launch_bounds__(kBwdThreads, 1)
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

  const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0);   // lane-0 shuffle: the compiler then knows it is warp-uniform
  const int lane = threadIdx.x & 31;
  const int f = int(blockIdx.x % uint32_t(a.kv_tiles));      // this CTA's K/V tile
  const int bh = int(blockIdx.x / uint32_t(a.kv_tiles));
  const int b_idx = bh / a.H, h_idx = bh - b_idx * a.H;
  // causal: key j is seen by queries i >= j - off only, so the first query tile worth visiting is the one holding
  // query f*128 - off; a K/V tile that no query sees (N_kv > N + ...) has steps <= 0 and gets zero gradients
  int t_begin = 0;
  if (kCausal) t_begin = max(0, f * kBlockN - a.causal_off) / kBlockM;
  const int steps = a.q_tiles - t_begin;
  const bool empty = steps <= 0;

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

  auto load_tile = [&](const CUtensorMap* tm, unsigned int perm, uint32_t dst, uint32_t bar, int row0) {
#pragma unroll
    for (int h = 0; h < kNumBoxes; ++h)
      tma_load_tile(dst + h * kBoxBytes, tm, bar, h * kBoxCols, row0, h_idx, b_idx, perm);
  };

  if (warp == 8) {
    // =========================== TMA producer ===========================
    if (lane == 0 && !empty) {
      mbar_arrive_expect_tx(bar_f_full, 2 * kTileBytes);
      load_tile(&tmK, a.perm_kv, sK, bar_f_full, f * kBlockM);
      load_tile(&tmV, a.perm_kv, sV, bar_f_full, f * kBlockM);
      for (int t = 0; t < 2 * steps; ++t) {
        const uint32_t s = slot_of(t);
        mbar_wait(bar_t_empty + 8 * s, phase_of(t) ^ 1u, 500);
        mbar_arrive_expect_tx(bar_t_full + 8 * s, kTileBytes);
        load_tile((t & 1) ? &tmdO : &tmQ, (t & 1) ? a.perm_do : a.perm_q, sT + s * kTileBytes, bar_t_full + 8 * s,
                  (t_begin + (t >> 1)) * kBlockN);
      }
    }
    __syncwarp();
  } else if (warp == 9 && !empty) {
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
    const long long stat_base = (long long)bh * a.Nq;
    const float log2e = 1.4426950408889634f;
    // column statistics of this half's 64 queries of tile t: threads 0..63 fetch lse * log2e, threads 64..127 delta.
    // Queries beyond N and queries that saw no key (lse = -inf) get +inf, which makes their P exactly 0.
    // The global load for step n+1 is issued during step n, so its latency is not on the step's critical chain.
    auto fetch_stat = [&](int t) -> float {
      const int qi = t * kBlockN + h * kHalf + (row_in_tile & (kHalf - 1));
      if (row_in_tile < kHalf) {
        const float l = qi < a.Nq ? a.lse[stat_base + qi] : -INFINITY;
        return l == -INFINITY ? INFINITY : l * log2e;
      }
      return qi < a.Nq ? a.delta[stat_base + qi] : 0.f;
    };
    float stat_next = empty ? 0.f : fetch_stat(t_begin);
    for (int n = 0; n < steps; ++n) {
      const int t = t_begin + n;
      const uint32_t st = s_stats + uint32_t(h) * 1024u + uint32_t(n & 1) * 512u;
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(st + uint32_t(row_in_tile) * 4u), "f"(stat_next) : "memory");
      named_bar_sync(1 + h, 128);
      if (n + 1 < steps) stat_next = fetch_stat(t + 1);
      // visible query columns of this key row inside the half: c >= c_lo (causal: query >= key); keys beyond N see nothing
      int c_lo = 0;
      if (kCausal) c_lo = max(0, key_idx - a.causal_off - t * kBlockN - h * kHalf);
      if (key_idx >= a.Nkv) c_lo = kHalf;
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
    if (!empty) {
      mbar_wait(bar_acc_full, 0, 710);
      tc_fence_after();
    }
    constexpr uint32_t kChunksPerRow = kRowBytes / 16;
    const uint32_t swz = (kRowBytes == 128) ? uint32_t(row_in_tile & 7) : uint32_t((row_in_tile >> 1) & 3);
    const uint32_t tA = tmem_base + lane_addr + 2 * kBlockN + h * D;
    const uint32_t sO = h ? sV : sK;
    const float osc = h ? 1.f : a.scale;
#pragma unroll
    for (int q = 0; q < D / 32; ++q) {
      uint32_t orow[32];
      if (!empty) {
        tmem_ld32(tA + q * 32, orow);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) orow[e] = 0u;
      }
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
        tma_store_tile(h ? &tmdV : &tmdK, sO + bx * kBoxBytes, bx * kBoxCols, f * kBlockM, h_idx, b_idx, a.perm_dkv);
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
// dQ kernel.  S is released as soon
// as it is in registers, two dP buffers, the score-like products of step n+1 issued while step n is being
// exponentiated), but the 128 key columns of every score row are split between two warpgroups:
//     warps 0-3: keys 0..63 of the tile        warps 4-7: keys 64..127        (one thread per query row in both)
// The kernel was bound by its single compute warpgroup (3390 clk per step against 1536 clk of MMAs, tensor pipe 47 %
// active, ncu r02); two warps per SMSP halve the per-thread work and hide each other's MUFU latency.  No exchange is
// needed - the row statistics are inputs.  dS (16-bit) of keys 0..63 overwrites dP columns 0..31 and that of keys
// 64..127 dP columns 64..95: each warpgroup only overwrites fp32 columns it has itself already loaded, and the TS
// product takes its A operand per k-step (8 columns), so the two halves need not be adjacent.
template <int D, bool kBF16, bool kCausal>
__global__ void __launch_bounds__(kBwdThreads, 1)
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

  const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0);   // lane-0 shuffle: the compiler then knows it is warp-uniform
  const int lane = threadIdx.x & 31;
  const int f = int(blockIdx.x % uint32_t(a.q_tiles));       // this CTA's query tile
  const int bh = int(blockIdx.x / uint32_t(a.q_tiles));
  const int b_idx = bh / a.H, h_idx = bh - b_idx * a.H;
  // causal: key tiles up to the one holding the last key the tile's last query sees (none at all: dQ = 0)
  int steps = a.kv_tiles;
  if (kCausal) {
    const int last_row = min(f * kBlockM + kBlockM - 1, a.Nq - 1);
    const int last_col = last_row + a.causal_off;
    steps = last_col < 0 ? 0 : min(a.kv_tiles, last_col / kBlockN + 1);
  }
  const bool empty = steps <= 0;

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

  auto load_tile = [&](const CUtensorMap* tm, unsigned int perm, uint32_t dst, uint32_t bar, int row0) {
#pragma unroll
    for (int h = 0; h < kNumBoxes; ++h)
      tma_load_tile(dst + h * kBoxBytes, tm, bar, h * kBoxCols, row0, h_idx, b_idx, perm);
  };
  // dP(n) lives in TMEM columns 128.. (even n) or 384.. (odd n); S in columns 0..127; dQ in 256..256+D
  auto dp_col = [&](int n) { return uint32_t((n & 1) ? 3 * kBlockN : kBlockN); };

  if (warp == 8) {
    // =========================== TMA producer ===========================
    if (lane == 0 && !empty) {
      mbar_arrive_expect_tx(bar_f_full, 2 * kTileBytes);
      load_tile(&tmQ, a.perm_q, sQ, bar_f_full, f * kBlockM);
      load_tile(&tmdO, a.perm_do, sdO, bar_f_full, f * kBlockM);
      for (int t = 0; t < 2 * steps; ++t) {
        const uint32_t s = slot_of(t);
        mbar_wait(bar_t_empty + 8 * s, phase_of(t) ^ 1u, 500);
        mbar_arrive_expect_tx(bar_t_full + 8 * s, kTileBytes);
        load_tile((t & 1) ? &tmK : &tmV, a.perm_kv, sT + s * kTileBytes, bar_t_full + 8 * s, (t >> 1) * kBlockN);
      }
    }
    __syncwarp();
  } else if (warp == 9 && !empty) {
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
    const long long stat_base = (long long)bh * a.Nq;
    float lse2_r = INFINITY, delta_r = 0.f;     // +inf makes P exactly 0: rows beyond N, rows that saw no key (lse = -inf)
    if (q_idx < a.Nq) {
      const float l = a.lse[stat_base + q_idx];
      if (l != -INFINITY) lse2_r = l * 1.4426950408889634f;
      delta_r = a.delta[stat_base + q_idx];
    }
    for (int n = 0; n < steps; ++n) {
      // visible key columns of this row inside the half: c <= c_hi (key < N, causal: key <= query)
      int last = a.Nkv - 1;
      if (kCausal) last = min(last, q_idx + a.causal_off);
      int c_hi = min(kHalf - 1, last - n * kBlockN - h * kHalf);
      if (q_idx >= a.Nq) c_hi = -1;
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
    if (!empty) {
      mbar_wait(bar_acc_full, 0, 710);
      tc_fence_after();
    }
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
        if (!empty) {
          tmem_ld32(tA + q * 32, orow);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) orow[e] = 0u;
        }
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
        tma_store_tile(&tmdQ, sQ + bx * kBoxBytes, bx * kBoxCols, f * kBlockM, h_idx, b_idx, a.perm_dq);
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
