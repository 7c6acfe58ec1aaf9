// fa_fwd_sm100.cuh — warp-specialised FlashAttention forward for sm_100a (B200).
//
// Replaces the reference's device kernels on the hot path
//   code/cuda_fa1/flashAttention.cu:7-152              (flash_attention_forward, FA1, one thread/row)
//   code/cutlass_cuda_fa1/run/flash_attn_cutlass.cu:346-453 (WMMA FA1 kernel)
//   code/triton_fa2/FA2-triton.py:25-93                (Triton FA2 fwd; causal rule :70-73)
// with a from-scratch design:
//
//   CTA = 384 threads, one CTA per SM, two 128-row Q tiles per CTA (ping-pong).
//     warps 0-3  : softmax warpgroup for Q tile 0   (one thread owns one score row, no shuffles)
//     warps 4-7  : softmax warpgroup for Q tile 1
//     warp  8    : TMA producer  (Q once, then K_j / V_j through a multi-stage mbarrier ring)
//     warp  9    : tcgen05.mma issuer (one thread): S_i = Q_i K_j^T -> TMEM, O_i += P_i V_j
//     warp  10   : TMEM allocator / deallocator
//   TMEM (512 columns): S0 | S1 | O0 | O1, fp32.  P (bf16/fp16) is written back over the first
//   columns of its S tile and consumed by the PV MMA directly from TMEM (A-operand in TMEM).
//   Online softmax runs in the log2 domain with a lazily updated reference max: O and l are only
//   rescaled when the true row max has moved more than 2^8 above the reference max.
//   Epilogue: O/l -> 16-bit -> swizzled smem (re-using the dead Q tile) -> TMA store; the
//   logsumexp (and the reference's l, m) are written with coalesced fp32 stores.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>

#include "sm100_ptx.cuh"

namespace fa {

struct FwdArgs {
  float* lse;   // [BH, Nq] optional
  float* l;     // optional (reference semantics)
  float* m;     // optional (reference semantics)
  int Nq, Nkv;
  int causal_off;   // Nkv - Nq (bottom-right aligned causal; 0 for the square case)
  int num_q_blocks; // ceil(Nq / 256)
  float scale_log2; // softmax_scale * log2(e)
  long long stat_stride_bh;
  unsigned long long desc_hi_qk;  // K-major 128B-swizzle descriptor bits (Q, K)
  unsigned long long desc_hi_v;   // MN-major 128B-swizzle descriptor bits (V)
  unsigned int idesc_qk, idesc_pv;
  long long* trace;   // bring-up only (-DFA_TRACE): clock64 timeline of CTA 0, see tools/trace_report.py
};

#ifdef FA_TRACE
#define FA_TRACE_EV(j, ev)                                                          \
  do {                                                                              \
    if (blockIdx.x == 0 && a.trace != nullptr && (j) < 32) a.trace[(j) * 16 + (ev)] = clock64(); \
  } while (0)
#else
#define FA_TRACE_EV(j, ev) do {} while (0)
#endif

constexpr int kBlockM = 128;        // rows per Q tile
constexpr int kBlockN = 128;        // keys per K/V tile
constexpr int kNumThreads = 384;
constexpr int kSmemLimit = 232448;  // 227 KB opt-in maximum on sm_100
constexpr float kRescaleThreshold = 8.0f;  // log2 units

template <int D>
struct FwdTraits {
  static constexpr int kTileBytes = kBlockN * D * 2;       // one Q/K/V/O tile, 16-bit elements
  static constexpr int kBoxBytes = 128 * 64 * 2;           // one 64-column TMA box (128B swizzle)
  static constexpr int kNumBoxes = D / 64;
  static constexpr int kBarrierBytes = 1024;
  static constexpr int kStagesMax = (kSmemLimit - 1024 /*align slack*/ - kBarrierBytes - 2 * kTileBytes) / kTileBytes;
  static constexpr int kStages = kStagesMax > 8 ? 8 : kStagesMax;
  static constexpr int kSmemBytes = 1024 + (2 + kStages) * kTileBytes + kBarrierBytes;
  static_assert(D == 64 || D == 128, "head_dim must be 64 or 128");
  static_assert(kStages >= 2, "need at least a double-buffered K/V ring");
};

template <bool kBF16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  if constexpr (kBF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool kBF16>
__device__ __forceinline__ uint32_t pack2(float2 v) { return pack2<kBF16>(v.x, v.y); }

// 2^x for a pair of inputs on the FMA/ALU pipes (no MUFU): Cody-Waite split x = floor(x) + f,
// degree-4 minimax polynomial for 2^f on [0,1) (max relative error 3e-6 with p(0) = 1 exactly; fitted by
// linear programming on the relative error), floor(x) added straight into the exponent
// field.  Used for a fixed fraction of every score row to take load off the 16/clk/SM MUFU.EX2 unit
// (at d=128 the MMAs of one K/V tile take exactly as many cycles as its 16384 MUFU.EX2).
__device__ __forceinline__ float2 ex2_emulated(float2 x) {
  x.x = fmaxf(x.x, -127.f);
  x.y = fmaxf(x.y, -127.f);
  const float2 r = __fadd2_rd(x, make_float2(12582912.f, 12582912.f));     // 1.5*2^23 + floor(x)
  const float2 fl = __fadd2_rn(r, make_float2(-12582912.f, -12582912.f));  // floor(x), exact
  const float2 f = __ffma2_rn(fl, make_float2(-1.f, -1.f), x);             // x - floor(x) in [0,1)
  float2 p = __ffma2_rn(f, make_float2(0.013425154611468315f, 0.013425154611468315f),
                        make_float2(0.0522453747689724f, 0.0522453747689724f));
  p = __ffma2_rn(p, f, make_float2(0.24127855896949768f, 0.24127855896949768f));
  p = __ffma2_rn(p, f, make_float2(0.6930451393127441f, 0.6930451393127441f));
  p = __ffma2_rn(p, f, make_float2(1.f, 1.f));
  p.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(r.x) << 23));
  p.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(r.y) << 23));
  return p;
}

#ifndef FA_EMU_PAIRS_OF_4
#define FA_EMU_PAIRS_OF_4 1   // of every 4 score pairs, this many take the polynomial path
#endif

template <int D, bool kBF16, bool kCausal>
__global__ void __launch_bounds__(kNumThreads, 1)
fa_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                    const FwdArgs a) {
  using T = FwdTraits<D>;
  constexpr int kStages = T::kStages;
  constexpr uint32_t kTileBytes = T::kTileBytes;
  constexpr uint32_t kBoxBytes = T::kBoxBytes;
  constexpr int kNumBoxes = T::kNumBoxes;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // 128B swizzle atoms need 1 KB alignment
  const uint32_t sQ = smem_base;
  const uint32_t sKV = smem_base + 2 * kTileBytes;
  const uint32_t bars = smem_base + (2 + kStages) * kTileBytes;
  // barrier slots (8 bytes each)
  const uint32_t bar_q_full = bars;                        // [2]
  const uint32_t bar_s_full = bars + 16;                   // [2]  MMA -> softmax
  const uint32_t bar_o_full = bars + 32;                   // [2]  MMA -> softmax (PV done)
  const uint32_t bar_p_full = bars + 48;                   // [2 tiles][2 halves] softmax -> MMA (128 arrivals)
  const uint32_t bar_kv_full = bars + 80;                  // [kStages]
  const uint32_t bar_kv_empty = bars + 80 + 8 * kStages;   // [kStages]
  const uint32_t tmem_slot = bars + 80 + 16 * kStages;     // u32 written by tcgen05.alloc

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- work assignment: one CTA = 256 query rows of one (b,h)
  const int bh = blockIdx.x / a.num_q_blocks;
  const int qi = blockIdx.x - bh * a.num_q_blocks;
  const int qb = kCausal ? (a.num_q_blocks - 1 - qi) : qi;  // causal: longest blocks of a head first
  const int q0 = qb * 2 * kBlockM;

  // number of K/V tiles each Q tile visits (masked tiles above the diagonal are skipped)
  const int n_kv_tiles = (a.Nkv + kBlockN - 1) / kBlockN;
  auto tiles_for = [&](int r0) {
    int n = (r0 < a.Nq) ? n_kv_tiles : 0;
    if (kCausal && n > 0) {
      const int last_row = min(r0 + kBlockM - 1, a.Nq - 1);
      const int last_col = last_row + a.causal_off;   // largest visible key index
      n = last_col < 0 ? 0 : min(n, last_col / kBlockN + 1);
    }
    return n;
  };
  const int n_t0 = tiles_for(q0), n_t1 = tiles_for(q0 + kBlockM);
  const int n_max = max(n_t0, n_t1);

  // ---- one-time setup
  if (warp == 8 && lane == 0) {
    prefetch_tensormap(&tmQ);
    prefetch_tensormap(&tmK);
    prefetch_tensormap(&tmV);
    prefetch_tensormap(&tmO);
  }
  if (warp == 9 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_q_full + 8 * i, 1);
      mbar_init(bar_s_full + 8 * i, 1);
      mbar_init(bar_p_full + 16 * i, 128);
      mbar_init(bar_p_full + 16 * i + 8, 128);
      mbar_init(bar_o_full + 8 * i, 1);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_kv_full + 8 * s, 1);
      mbar_init(bar_kv_empty + 8 * s, 1);
    }
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

  if (warp >= 8) {
    setmaxnreg_dec<56>();
    if (warp == 8) {
      // =========================== TMA producer ===========================
      if (lane == 0) {
        auto load_tile = [&](const CUtensorMap* tm, uint32_t dst, uint32_t bar, int row0) {
          mbar_arrive_expect_tx(bar, kTileBytes);
#pragma unroll
          for (int h = 0; h < kNumBoxes; ++h) tma_load_3d(dst + h * kBoxBytes, tm, bar, h * 64, row0, bh);
        };
        if (n_t0 > 0) load_tile(&tmQ, sQ, bar_q_full, q0);
        int it = 0;
        for (int j = 0; j < n_max; ++j) {
#pragma unroll
          for (int kv = 0; kv < 2; ++kv) {
            const int stage = it % kStages;
            const uint32_t ph = (it / kStages) & 1;
            mbar_wait(bar_kv_empty + 8 * stage, ph ^ 1, 100 + kv);
            load_tile(kv == 0 ? &tmK : &tmV, sKV + stage * kTileBytes, bar_kv_full + 8 * stage,
                      j * kBlockN);
            ++it;
            if (j == 0 && kv == 0 && n_t1 > 0)
              load_tile(&tmQ, sQ + kTileBytes, bar_q_full + 8, q0 + kBlockM);
          }
        }
      }
      __syncwarp();
    } else if (warp == 9) {
      // =========================== MMA issuer ===========================
      // The whole warp runs the (uniform) control flow and the barrier waits; one elected lane issues
      // the tcgen05.mma / tcgen05.commit instructions.  Keeping the flow warp-uniform lets the compiler
      // hold descriptors in uniform registers: the issue cost per MMA must stay well under the 64
      // cycles one 128x128x16 MMA takes on the tensor pipe.
      const uint32_t hi_qk = uint32_t(a.desc_hi_qk >> 32), hi_v = uint32_t(a.desc_hi_v >> 32);
      const uint32_t lo_qk = uint32_t(a.desc_hi_qk), lo_v = uint32_t(a.desc_hi_v);
      const uint32_t idesc_qk = a.idesc_qk, idesc_pv = a.idesc_pv;
      int trace_j = 0;
      (void)trace_j;
      // S_i = Q_i K^T, then signal `bar_done` (and optionally release the K stage) when it has landed
      auto issue_qk = [&](int i, int stage, uint32_t bar_done, uint32_t bar_release) {
        const uint32_t a_lo = lo_qk | ((sQ + i * kTileBytes) >> 4);
        const uint32_t b_lo = lo_qk | ((sKV + stage * kTileBytes) >> 4);
        const uint32_t d_tmem = tmem_base + i * kBlockN;
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const uint32_t off = ((k / 4) * kBoxBytes + (k % 4) * 32) >> 4;
            umma_ss(d_tmem, a_lo + off, hi_qk, b_lo + off, hi_qk, idesc_qk, k > 0 ? 1u : 0u);
          }
          umma_commit(bar_done);
          if (bar_release) umma_commit(bar_release);
        }
        __syncwarp();
      };
      // O_i += P_i V_j, issued in two halves of 64 keys: the first half starts as soon as the softmax
      // warpgroup has published the first 64 columns of P, while it is still exponentiating the rest.
      auto issue_pv = [&](int i, int stage, bool acc, uint32_t parity, uint32_t bar_done, uint32_t bar_release) {
        const uint32_t b_lo = lo_v | ((sKV + stage * kTileBytes) >> 4);
        const uint32_t p_tmem = tmem_base + i * kBlockN;           // P aliases S_i
        const uint32_t d_tmem = tmem_base + 2 * kBlockN + i * D;   // O_i
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(bar_p_full + 16 * i + 8 * h, parity, 212 + 2 * i + h);
          tc_fence_after();
          if (lane == 0) FA_TRACE_EV(trace_j, 8 + 3 * i + h);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 4 * h; k < 4 * h + 4; ++k)
              umma_ts(d_tmem, p_tmem + k * 8, b_lo + k * (16 * 128 / 16), hi_v, idesc_pv, (acc || k > 0) ? 1u : 0u);
            if (h == 1) {
              umma_commit(bar_done);
              if (bar_release) umma_commit(bar_release);
            }
          }
          __syncwarp();
        }
      };
      auto stage_of = [&](int it) { return it % kStages; };
      auto phase_of = [&](int it) { return uint32_t((it / kStages) & 1); };

      if (n_max > 0) {
        // S_i(0) = Q_i K_0^T
        mbar_wait(bar_kv_full + 8 * stage_of(0), phase_of(0), 200);
        const uint32_t rel = bar_kv_empty + 8 * stage_of(0);
        if (n_t0 > 0) {
          mbar_wait(bar_q_full, 0, 201);
          tc_fence_after();
          issue_qk(0, stage_of(0), bar_s_full, n_t1 > 0 ? 0u : rel);
        }
        if (n_t1 > 0) {
          mbar_wait(bar_q_full + 8, 0, 202);
          tc_fence_after();
          issue_qk(1, stage_of(0), bar_s_full + 8, rel);
        }
      }
      for (int j = 0; j < n_max; ++j) {
        const int it_v = 2 * j + 1, it_k = 2 * j + 2;
        const bool has_next = (j + 1 < n_max);
        mbar_wait(bar_kv_full + 8 * stage_of(it_v), phase_of(it_v), 210);
        if (has_next) mbar_wait(bar_kv_full + 8 * stage_of(it_k), phase_of(it_k), 211);
        trace_j = j;
        const uint32_t rel_v = bar_kv_empty + 8 * stage_of(it_v);
        const uint32_t rel_k = bar_kv_empty + 8 * stage_of(it_k);
        // n_t1 >= n_t0 whenever both tiles exist (the second tile sits lower in the causal triangle);
        // tile 1 is absent (n_t1 == 0) only for a ragged last block.
        if (j < n_t0) issue_pv(0, stage_of(it_v), j > 0, j & 1, bar_o_full, j < n_t1 ? 0u : rel_v);
        if (j + 1 < n_t0) {
          if (lane == 0) FA_TRACE_EV(j, 10);
          issue_qk(0, stage_of(it_k), bar_s_full, j + 1 < n_t1 ? 0u : rel_k);
        }
        if (j < n_t1) issue_pv(1, stage_of(it_v), j > 0, j & 1, bar_o_full + 8, rel_v);
        if (j + 1 < n_t1) {
          if (lane == 0) FA_TRACE_EV(j, 13);
          issue_qk(1, stage_of(it_k), bar_s_full + 8, rel_k);
        }
      }
    }
  } else {
    // =========================== softmax warpgroups ===========================
    setmaxnreg_inc<224>();
    const int i = warp >> 2;           // which Q tile
    const int wl = warp & 3;           // warp within the warpgroup == TMEM lane quarter
    const int row_in_tile = wl * 32 + lane;
    const int row = q0 + i * kBlockM + row_in_tile;
    const int n_i = i ? n_t1 : n_t0;
    const uint32_t lane_addr = uint32_t(wl * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + i * kBlockN;
    const uint32_t tO = tmem_base + lane_addr + 2 * kBlockN + i * D;
    const uint32_t b_s_full = bar_s_full + 8 * i;
    const uint32_t b_p_full = bar_p_full + 16 * i;
    const uint32_t b_o_full = bar_o_full + 8 * i;

    const float c = a.scale_log2;
    // last visible key index for this row
    int limit = a.Nkv - 1;
    if (kCausal) limit = min(limit, row + a.causal_off);

    float m_ref = -INFINITY;   // max used in the exponents (lags the true max by <= threshold)
    float m_true = -INFINITY;  // true running row max (scaled, log2 units)
    float l = 0.f;             // sum of 2^(t - m_ref)

    for (int j = 0; j < n_i; ++j) {
      mbar_wait(b_s_full, j & 1, 300 + i);
      tc_fence_after();
      if (row_in_tile == 0) FA_TRACE_EV(j, 4 * i + 0);
      uint32_t sr[4][32];
#pragma unroll
      for (int q = 0; q < 4; ++q) tmem_ld32(tS + q * 32, sr[q]);
      tmem_wait_ld();
      if (row_in_tile == 0) FA_TRACE_EV(j, 4 * i + 1);

      // masking: key index > limit -> -inf (diagonal tiles of causal runs, ragged last tile)
      const int lim_local = limit - j * kBlockN;
      if (__any_sync(0xffffffffu, lim_local < kBlockN - 1)) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (q * 32 + k > lim_local) sr[q][k] = 0xff800000u;  // -inf
      }

      // row max with four independent chains (a single chain of 64 dependent FMNMX3 costs ~4 clk each)
      float mxp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < 32; k += 8)
#pragma unroll
          for (int u = 0; u < 4; ++u)
            mxp[u] = fmaxf(mxp[u], fmaxf(__uint_as_float(sr[q][k + 2 * u]), __uint_as_float(sr[q][k + 2 * u + 1])));
      const float mx = fmaxf(fmaxf(mxp[0], mxp[1]), fmaxf(mxp[2], mxp[3]));
      const float m_new = fmaxf(m_true, mx * c);
      m_true = m_new;

      if (j == 0) {
        m_ref = m_new;
      } else {
        const bool moved = (m_new - m_ref) > kRescaleThreshold;
        if (__any_sync(0xffffffffu, moved)) {
          // rescale this warp's 32 rows of O (and l) to the new reference max
          const float alpha = (m_new == -INFINITY) ? 1.f : ex2_approx(m_ref - m_new);
          m_ref = m_new;
          l *= alpha;
          mbar_wait(b_o_full, (j - 1) & 1, 310 + i);  // PV(j-1) has landed in TMEM
          tc_fence_after();
#pragma unroll
          for (int q = 0; q < D / 32; ++q) {
            uint32_t orow[32];
            tmem_ld32(tO + q * 32, orow);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 32; ++k) orow[k] = __float_as_uint(__uint_as_float(orow[k]) * alpha);
            tmem_st32(tO + q * 32, orow);
          }
        }
      }
      const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;

      // p = 2^(s*c - m_ref) with packed (2-wide) fp32 math; FA_EMU_PAIRS_OF_4 of every 4 pairs take the
      // polynomial path, the rest MUFU.EX2.  P is published to the MMA warp in two halves of 64 keys.
      const float2 c2 = make_float2(c, c), neg_m2 = make_float2(neg_m, neg_m);
      float2 lsum2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t pk[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(sr[q][2 * k]), __uint_as_float(sr[q][2 * k + 1])),
                                      c2, neg_m2);
          float2 pv;
          if ((k & 3) < FA_EMU_PAIRS_OF_4) {
            pv = ex2_emulated(x);
          } else {
            pv.x = ex2_approx(x.x);
            pv.y = ex2_approx(x.y);
          }
          lsum2[k & 3] = __fadd2_rn(lsum2[k & 3], pv);
          pk[k] = pack2<kBF16>(pv);
        }
        tmem_st16(tS + q * 16, pk);   // P(16-bit) over the first 64 columns of S
        if (q & 1) {
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(b_p_full + 8 * (q >> 1));
          if (row_in_tile == 0) FA_TRACE_EV(j, 4 * i + 2 + (q >> 1));
        }
      }
      const float2 lsum = __fadd2_rn(__fadd2_rn(lsum2[0], lsum2[1]), __fadd2_rn(lsum2[2], lsum2[3]));
      l += lsum.x + lsum.y;
    }

    // ---- epilogue: O_i / l -> 16-bit -> swizzled smem (dead Q tile) -> TMA store
    if (q0 + i * kBlockM < a.Nq) {
      if (n_i > 0) {
        mbar_wait(b_o_full, (n_i - 1) & 1, 320 + i);
        tc_fence_after();
      }
      if (limit < 0) l = 0.f;   // row sees no key at all: the polynomial exp2 returns 2^-127, not 0
      const float inv_l = (l > 0.f) ? (1.f / l) : 0.f;
      const uint32_t sO = sQ + i * kTileBytes;
#pragma unroll
      for (int q = 0; q < D / 32; ++q) {
        uint32_t orow[32];
        if (n_i > 0) {
          tmem_ld32(tO + q * 32, orow);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) orow[k] = 0u;
        }
        const uint32_t box = sO + (q / 2) * kBoxBytes + row_in_tile * 128;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            w[e] = pack2<kBF16>(__uint_as_float(orow[v * 8 + 2 * e]) * inv_l,
                                __uint_as_float(orow[v * 8 + 2 * e + 1]) * inv_l);
          const uint32_t chunk = uint32_t((q & 1) * 4 + v);
          const uint32_t addr = box + ((chunk ^ uint32_t(row_in_tile & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]),
                       "r"(w[2]), "r"(w[3]) : "memory");
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + i, 128);
      if (row_in_tile == 0) {
#pragma unroll
        for (int h = 0; h < kNumBoxes; ++h)
          tma_store_3d(&tmO, sO + h * kBoxBytes, h * 64, q0 + i * kBlockM, bh);
        tma_store_commit();
      }
      if (row < a.Nq) {
        const long long off = (long long)bh * a.stat_stride_bh + row;
        const float ln2 = 0.6931471805599453f;
        const bool any = l > 0.f;
        if (a.lse) a.lse[off] = any ? fmaf(m_ref, ln2, logf(l)) : -INFINITY;
        if (a.m) a.m[off] = any ? m_true * ln2 : -INFINITY;
        if (a.l) a.l[off] = any ? l * ex2_approx(m_ref - m_true) : 0.f;
      }
      if (row_in_tile == 0) tma_store_wait_read<0>();
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace fa
