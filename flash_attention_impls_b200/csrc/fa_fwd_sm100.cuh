// fa_fwd_sm100.cuh — warp-specialised FlashAttention forward for sm_100a (B200).
//
// Replaces the reference's device kernels on the hot path
//   code/cuda_fa1/flashAttention.cu:7-152              (flash_attention_forward, FA1, one thread/row)
//   code/cutlass_cuda_fa1/run/flash_attn_cutlass.cu:346-453 (WMMA FA1 kernel)
//   code/triton_fa2/FA2-triton.py:25-93                (Triton FA2 fwd; causal rule :70-73)
// with a from-scratch design:
//
//   Persistent kernel with a hardware tile scheduler: the grid has one CTA per work item, but a running CTA
//   (512 threads, one per SM) keeps its TMEM, barriers and pipelines and steals the next not-yet-launched
//   blockIdx through cluster launch control (clusterlaunchcontrol.try_cancel), so scheduling is dynamic
//   (longest-first for causal) without any global counter.  A work item is 256 query rows of one (b,h): two
//   128-row Q tiles that ping-pong on the tensor pipe; the item list is head-major (causal: longest
//   q-blocks of a head first) so resident CTAs share K/V in L2.
//     warps 0-3   : softmax warpgroup for Q tile 0   (one thread owns one score row, no shuffles)
//     warps 4-7   : softmax warpgroup for Q tile 1
//     warps 8-11  : correction/epilogue warpgroup: O_i * (1/l) -> 16-bit -> swizzled smem -> TMA store,
//                   overlapped with the next work item's MMAs and softmax
//     warp  12    : TMA producer  (Q per item, K_j / V_j through a multi-stage mbarrier ring that runs
//                   across items, so the next item's first tiles are prefetched)
//     warp  13    : tcgen05.mma issuer (one elected lane): S_i = Q_i K_j^T -> TMEM, O_i += P_i V_j
//     warp  14    : TMEM allocator / deallocator
//   TMEM (512 columns), d = 128: S0 | S1 | O0 | O1, fp32; P (bf16/fp16) is written back over the first
//   columns of its S tile and consumed by the PV MMA directly from TMEM (A-operand in TMEM).
//   d <= 64: S | P0 | P1 | O0 | O1 - one score buffer shared by the two Q tiles (a softmax warpgroup copies
//   its S into registers at once and hands the buffer back), P in columns of its own, so Q K^T of step j+1
//   is issued before P V of step j (see FA_SHARED_S below).
//   Online softmax runs in the log2 domain with a lazily updated reference max: O and l are only
//   rescaled when the true row max has moved more than 2^8 above the reference max; at d = 128, tiles
//   that provably need no new reference max skip the row-max pass altogether (fast softmax path: the row sum
//   of each published part of P bounds its entries; predicated barrier arrivals, redo on the exact path).
//   The logsumexp (and the reference's l, m) are written with coalesced fp32 stores.
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
  int num_items;    // B*H*num_q_blocks work items (= grid size; later ones are stolen by resident CTAs)
  int num_bh;       // B*H
  int group_heads;  // causal item ordering: heads per L2-sized group (see get_item)
  // split-KV: the (b,h,q-block) items from list position `split_begin` on are each cut into `nsplit` items that visit
  // `tiles_per_split` consecutive K/V tiles and write a partial (O, lse, m) into the caller's workspace, laid out by item:
  // [nsplit][num_ws_items][256 rows][d] / [nsplit][num_ws_items][256]; fa::item_combine_kernel merges them.
  //   split_begin == 0: launches with far fewer items than SMs (every item is split);
  //   split_begin  > 0: "tail split" - a launch whose last, partly filled wave of equal items would cost a whole round
  //                     (512 items on 148 SMs = 3.46 waves) gets only that tail cut, so the tail takes 1/nsplit of a round.
  // nsplit == 1: off.
  int nsplit, tiles_per_split, split_begin, num_ws_items;
  float* ws_lse;    // [nsplit][num_ws_items][256]
  float* ws_m;
  float scale_log2; // softmax_scale * log2(e)
  int need_stats;   // the caller wants l / m (reference semantics): every tile takes the exact-row-max path
  unsigned int zero;   // always 0; opaque to the compiler (pins the position of an mbarrier arrival in the instruction stream)
  long long stat_stride_b, stat_stride_h;
  int H;            // heads per batch (a work item's bh is split into (b, h) for the 4-D tensor maps)
  // Order of the three outer tensor-map axes for Q, K/V and O: axis k of the map carries the row (0), head (1) or
  // batch (2) coordinate, 2 bits each (the host sorts the axes by stride, see make_tmap).
  unsigned int perm_q, perm_kv, perm_o, perm_w;
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
constexpr int kNumThreads = 512;
constexpr int kSmemLimit = 232448;  // 227 KB opt-in maximum on sm_100
constexpr float kRescaleThreshold = 8.0f;  // log2 units
// Fast softmax path (no row-max pass; head dim 128 only): a tile keeps the reference max it inherited as long as the sum
// of the exponentials of each published part stays below this bound (so every p < 2^15: finite in fp16, harmless in bf16
// and in the fp32 sums); otherwise the tile is redone on the exact path.  0 compiles the fast path out (A/B switch).
// Measured on zero inputs (SM clock at its maximum, i.e. in cycles): c3 1572 -> 1617 TFLOP/s, on Set S c3 best-of-20
// 1391 -> 1413, c4 1284 -> 1296; at d = 64 the kernel is MUFU-bound and the path costs 3 %, so it is not compiled in there
// (profiles/r02_fast_softmax_ab.log).
constexpr float kFastSumLimit = 32768.0f;
// Timing-only ablations (results are WRONG; zero-input runs only): which instruction class the period is sensitive to.
//   1: no row-sum FADD2   2: no FFMA2 (scale / subtract)   4: MUFU.EX2 on every other pair only   8: no row-max pass
//   16: no tcgen05.wait::st before P is published   32: no F2FP pack   64: QK^T with half the k-steps   128: PV likewise
#ifndef FA_ABLATE
#define FA_ABLATE 0
#endif
// ---- compile-time switches (defaults = the shipped kernel; the others exist for interleaved A/B builds, `make variant`)
//
// FA_SHARED_S / FA_SHARED_S_MAX_D - TMEM protocol for head dims <= FA_SHARED_S_MAX_D (default 64):
//   ONE score buffer shared by the two Q tiles, P in columns of its own:  S (128) | P0 (64) | P1 (64) | O0 (D) | O1 (D).
//   A softmax warpgroup copies its S into registers at once (66 clk) and hands the buffer back (s_free), so QK^T of step
//   j+1 is issued while the softmax of step j is still running instead of behind P V of step j: the score product leaves
//   the softmax -> PV -> QK^T -> softmax chain.  Larger head dims (and the precise mode at any d, which needs 128 P
//   columns per tile) keep the round-1 protocol S0 | S1 | O0 | O1 with P written over the first 64 columns of its own S
//   tile; at d = 128 the shared buffer serialises the two tiles' QK^T through the copy-out hand-offs and measured slower
//   (1617 -> 1534 TFLOP/s on zero inputs; +1.4 % after the uniformity hints, within noise on real data).
#ifndef FA_SHARED_S
#define FA_SHARED_S 1
#endif
#ifndef FA_SHARED_S_MAX_D
#define FA_SHARED_S_MAX_D 64
#endif
// FA_FAST_SOFTMAX / FA_FAST_MIN_D - the fast softmax path (see kFastSumLimit) for head dims >= FA_FAST_MIN_D (default 128).
#ifndef FA_FAST_SOFTMAX
#define FA_FAST_SOFTMAX 1
#endif
#ifndef FA_FAST_MIN_D
#define FA_FAST_MIN_D 128
#endif
// FA_UNIFORM_HINT - warp index and work-item ids pass through __shfl_sync(.., 0), which tells the compiler they are
//   warp-uniform: item state, TMEM / barrier addresses and UMMA descriptor arithmetic then live in uniform registers
//   instead of being recomputed from threadIdx after every barrier wait (c3 on zero inputs 1616 -> 1719 TFLOP/s, no spill
//   left in the 40-register issuer warp).
#ifndef FA_UNIFORM_HINT
#define FA_UNIFORM_HINT 1
#endif
// FA_EP_NAMED_BAR - the epilogue warpgroup's item-long wait for 1/l is a hardware named barrier (no polling) instead of
//   an mbarrier (+0.3 %).
#ifndef FA_EP_NAMED_BAR
#define FA_EP_NAMED_BAR 1
#endif
// FA_MMA_SPIN - the MMA warp polls its barriers with mbarrier.test_wait in a tight loop instead of the suspending
//   try_wait.  Measured 5 % slower (the poll steals issue slots from the softmax warps of its SMSP); default 0.
#ifndef FA_MMA_SPIN
#define FA_MMA_SPIN 0
#endif
// FA_TWO_ISSUERS - TWO MMA issuer warps, one per Q tile (warps 13 and 15), each walking only its own tile's chain (S_i ->
//   softmax -> P_i V -> Q_i K^T) so that a wait for the other tile's P never holds up this tile's MMAs and nothing is
//   serialised at a work-item boundary; K/V stages are released by two arrivals (an issuer that does not use a stage
//   arrives for it after seeing it filled); d <= 64 TMEM map S0 | S1 | P0 | P1 | O0 | O1.  Correct (parity list green) but
//   MEASURED MUCH SLOWER at d = 128: c3 on zero inputs 1615 -> 1168 TFLOP/s (the same on random inputs), c2 523 -> 485,
//   d = 64 long sequences unchanged (873 -> 869): MMAs of two issuing threads interleave on the tensor pipe, and every
//   switch between accumulator tiles costs what a fixed order pays only four times per step.  Default 0: one issuer
//   (warp 13) for both tiles, fixed interleaved order (profiles/r02_fast_softmax_ab.log).
#ifndef FA_TWO_ISSUERS
#define FA_TWO_ISSUERS 0
#endif

template <int D>
struct FwdTraits {
  static constexpr int kTileBytes = kBlockN * D * 2;       // one Q/K/V/O tile, 16-bit elements
  // A tile is stored as TMA boxes of 128 rows x kBoxCols columns: 64 columns (128-byte rows, 128B swizzle) for
  // d >= 64, one 32-column box (64-byte rows, 64B swizzle) for d = 32.
  static constexpr int kBoxCols = D >= 64 ? 64 : 32;
  static constexpr int kRowBytes = kBoxCols * 2;
  static constexpr int kBoxBytes = 128 * kRowBytes;
  static constexpr int kNumBoxes = D / kBoxCols;
  static constexpr int kAuxBytes = 3072;                   // mbarriers (512 B) + 1/l hand-off (1 KB) + slack
  // shared memory: Q0 | Q1 | O staging (one tile) | K/V ring | aux
  static constexpr int kStagesMax = (kSmemLimit - kAuxBytes - 3 * kTileBytes) / kTileBytes;
  static constexpr int kStages = kStagesMax > 8 ? 8 : kStagesMax;
  static constexpr int kSmemBytes = (3 + kStages) * kTileBytes + kAuxBytes;
  static_assert(D == 32 || D == 64 || D == 128, "head_dim must be 32, 64 or 128");
  static_assert(kStages >= 3, "K/V ring too shallow");
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

template <bool kBF16>
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  if constexpr (kBF16) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  } else {
    return __half22float2(*reinterpret_cast<__half2*>(&u));
  }
}

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

// Of every 4 score pairs, this many take the polynomial path instead of MUFU.EX2.  Measured on the final
// (persistent, 208-register softmax) kernel: 0 -> c4 1276 / causal-16k 1332 TFLOP/s, 1 -> 1250 / 1307,
// 2 -> 1184 / 1237 (profiles/r01_sweep_v7_exp2_fraction.log): the extra FMA-pipe instructions and registers
// cost more than the MUFU cycles they free, so the default is 0; the switch stays for d = 64 experiments.
#ifndef FA_EMU_PAIRS_OF_4
#define FA_EMU_PAIRS_OF_4 0
#endif
// ... and, of every 8 pairs, for head dims <= 64 (instantiations 32 / 64), where the MUFU.EX2 unit, not the tensor pipe, is
// the bound (2 exponentials per MMA clock).  Round 2, after the uniform-register hints had taken the address arithmetic off the
// softmax warps: 2 of 8 -> 4,32,8192,64 900 -> 971 TFLOP/s on zero inputs, 895 -> 963 on Set S (cuDNN: 978), causal 847 -> 875,
// c2 543 -> 555; 4 of 8 -> 879 (profiles/r02_fast_softmax_ab.log).  At d = 128 25 % is +0.7 % in cycles and -7 % on real data
// (the extra FMA-pipe work costs power), so FA_EMU_PAIRS_OF_4 stays 0.
#ifndef FA_EMU_PAIRS_OF_8_D64
#define FA_EMU_PAIRS_OF_8_D64 2
#endif

// P is published to the MMA warp in two parts: the first FA_P_FIRST_Q groups of 32 keys, then the rest.  The part
// that is published last sits on the softmax -> PV -> QK^T -> softmax chain: with 2 (halves) four PV MMAs (256 clk at
// d = 128) follow the last arrival, with 3 only two (128 clk), at the price of a later start of the first burst.
#ifndef FA_P_FIRST_Q
#define FA_P_FIRST_Q 2
#endif

// Tensor maps are 4-D {d, a1, a2, a3}; `perm` says which of (row, head, batch) each outer axis carries.
__device__ __forceinline__ void tma_load_tile(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int col, int row,
                                              int h, int b, unsigned int perm) {
  auto pick = [&](unsigned int r) { return r == 0u ? row : (r == 1u ? h : b); };
  tma_load_4d(dst, tm, bar, col, pick(perm & 3u), pick((perm >> 2) & 3u), pick((perm >> 4) & 3u));
}
__device__ __forceinline__ void tma_store_tile(const CUtensorMap* tm, uint32_t src, int col, int row, int h, int b,
                                               unsigned int perm) {
  auto pick = [&](unsigned int r) { return r == 0u ? row : (r == 1u ? h : b); };
  tma_store_4d(tm, src, col, pick(perm & 3u), pick((perm >> 2) & 3u), pick((perm >> 4) & 3u));
}

// One work item = 256 query rows of one (b,h).  Every role walks the same deterministic item list.
struct WorkItem {
  int bh, q0, n_t0, n_t1, n_max;
  int kv_begin;          // first K/V tile this item visits (non-zero only with split-KV)
  int out_b;             // batch index of the outputs
  int ws_item, split;    // split items: position in the workspace, split index
  bool split_out;        // outputs go to the workspace (a partial), not to the caller's O / lse / l / m
  bool valid0, valid1;   // tile has at least one real query row
};

// (b*H+h, q-block) of the un-split list position w.
template <bool kCausal>
__host__ __device__ __forceinline__ void item_coords(const FwdArgs& a, int w, int& bh, int& qb) {
  if (kCausal) {
    // Causal items differ in length (q-block qb visits qb+1.. K/V tiles), so the list is ordered
    // longest-first, but only within groups of `group_heads` heads whose K/V fit the L2 together:
    // group-major, then q-block descending, then head.  This keeps the K/V of the running CTAs shared in
    // L2 and leaves only the shortest items of the last group for the tail of the launch.
    const int per_group = a.group_heads * a.num_q_blocks;
    const int g = w / per_group;
    const int r = w - g * per_group;
    const int heads = (a.group_heads < a.num_bh - g * a.group_heads) ? a.group_heads : (a.num_bh - g * a.group_heads);
    const int qi = r / heads;
    bh = g * a.group_heads + (r - qi * heads);
    qb = a.num_q_blocks - 1 - qi;
  } else {
    bh = w / a.num_q_blocks;
    qb = w - bh * a.num_q_blocks;
  }
}

template <bool kCausal>
__host__ __device__ __forceinline__ WorkItem get_item(const FwdArgs& a, int w) {
  WorkItem it;
  int qb;
  it.split = 0;
  it.ws_item = 0;
  it.split_out = false;
  if (a.nsplit > 1 && w >= a.split_begin) {   // split index is the fastest-varying part of the item id
    const int u = w - a.split_begin;
    it.split = u % a.nsplit;
    it.ws_item = u / a.nsplit;
    it.split_out = true;
    w = a.split_begin + it.ws_item;
  }
  item_coords<kCausal>(a, w, it.bh, qb);
  it.q0 = qb * 2 * kBlockM;
  const int n_kv_tiles = (a.Nkv + kBlockN - 1) / kBlockN;
  // number of K/V tiles each Q tile visits (masked tiles above the diagonal are skipped)
  auto tiles_for = [&](int r0) {
    int n = (r0 < a.Nq) ? n_kv_tiles : 0;
    if (kCausal && n > 0) {
      const int last_row = (r0 + kBlockM - 1 < a.Nq - 1) ? (r0 + kBlockM - 1) : (a.Nq - 1);
      const int last_col = last_row + a.causal_off;   // largest visible key index
      const int n_vis = last_col / kBlockN + 1;
      n = last_col < 0 ? 0 : (n < n_vis ? n : n_vis);
    }
    return n;
  };
  it.n_t0 = tiles_for(it.q0);
  it.n_t1 = tiles_for(it.q0 + kBlockM);
  it.kv_begin = 0;
  it.out_b = it.bh / a.H;
  if (it.split_out) {       // restrict both tiles to this split's K/V tile range
    it.kv_begin = it.split * a.tiles_per_split;
    auto clampn = [&](int n) {
      n -= it.kv_begin;
      return n < 0 ? 0 : (n < a.tiles_per_split ? n : a.tiles_per_split);
    };
    it.n_t0 = clampn(it.n_t0);
    it.n_t1 = clampn(it.n_t1);
  }
  it.n_max = it.n_t0 > it.n_t1 ? it.n_t0 : it.n_t1;
  it.valid0 = it.q0 < a.Nq;
  it.valid1 = it.q0 + kBlockM < a.Nq;
  return it;
}

// kPrecise: P goes to the PV MMA as TWO 16-bit operands, P = P_hi + P_lo (P_lo = the rounding residual of P_hi,
// written to the upper 64 columns of the S tile, which are free once S is in registers), and O += P_hi V + P_lo V.
// That gives P fp32-like accuracy, i.e. the accuracy of the reference's CUDA-core FA1 kernel, which keeps P in
// fp32 (flashAttention.cu:107-135); with Q = K = V ~ N(0, 0.02) (the reference's own test inputs, main.cu:43-61)
// the 2^-12 rounding of a single fp16 P shows up as up to 3 % in the reference's symmetric-relative metric on
// outputs of magnitude 5e-6 (absolute error 6e-7), above its 2 % gate (main.cu:346).  Costs 1.5x the MMAs and four
// more instructions per score pair; used by fa_b200_forward_legacy and on request (fa_b200_params.precise).
template <int D, bool kBF16, bool kCausal, bool kPrecise = false>
__global__ void __launch_bounds__(kNumThreads, 1)
fa_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                    const __grid_constant__ CUtensorMap tmW, const FwdArgs a) {
  using T = FwdTraits<D>;
  constexpr bool kTwoIssuers = (FA_TWO_ISSUERS != 0);
  constexpr bool kSepP = kTwoIssuers && !kPrecise && D <= FA_SHARED_S_MAX_D;     // S0 | S1 | P0 | P1 | O0 | O1
  constexpr bool kSharedS = !kTwoIssuers && (FA_SHARED_S != 0) && !kPrecise && D <= FA_SHARED_S_MAX_D;
  // TMEM columns (fp32 columns; P is 16-bit, 64 columns per tile)
  constexpr uint32_t kColS1 = kSharedS ? 0u : uint32_t(kBlockN);              // S of tile 1 (tile 0: column 0)
  constexpr uint32_t kColP0 = kSharedS ? uint32_t(kBlockN) : (kSepP ? 2u * kBlockN : 0u);               // P of tile 0
  constexpr uint32_t kColP1 = kSharedS ? uint32_t(kBlockN + kBlockN / 2) : (kSepP ? 2u * kBlockN + kBlockN / 2 : uint32_t(kBlockN));
  constexpr uint32_t kColO = kSepP ? 3u * kBlockN : 2u * kBlockN;             // O_i at kColO + i * D
  static_assert(kColO + 2 * D <= 512, "TMEM map does not fit");
  constexpr int kStages = T::kStages;
  constexpr uint32_t kTileBytes = T::kTileBytes;
  constexpr uint32_t kBoxBytes = T::kBoxBytes;
  constexpr int kNumBoxes = T::kNumBoxes;
  constexpr int kBoxCols = T::kBoxCols;
  constexpr uint32_t kRowBytes = T::kRowBytes;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();   // 128B swizzle atoms need 1 KB alignment
  const uint32_t sQ = smem_base;
  const uint32_t sO = smem_base + 2 * kTileBytes;          // epilogue staging, one tile
  const uint32_t sKV = smem_base + 3 * kTileBytes;
  const uint32_t bars = smem_base + (3 + kStages) * kTileBytes;
  // barrier slots (8 bytes each)
  const uint32_t bar_q_full = bars;                        // [2]  TMA -> MMA
  const uint32_t bar_q_empty = bars + 16;                  // [2]  MMA -> TMA (all QK^T of the item issued+done)
  const uint32_t bar_s_full = bars + 32;                   // [2]  MMA -> softmax
  const uint32_t bar_o_full = bars + 48;                   // [2]  MMA -> softmax/epilogue (a PV has landed)
  const uint32_t bar_o_free = bars + 64;                   // [2]  epilogue -> MMA (O_i read out of TMEM)
  const uint32_t bar_ep_full = bars + 80;                  // [2]  softmax -> epilogue (1/l published)
  const uint32_t bar_ep_empty = bars + 96;                 // [2]  epilogue -> softmax (1/l slot consumed)
  const uint32_t bar_p_full = bars + 112;                  // [2 tiles][2 halves] softmax -> MMA (128 arrivals)
  const uint32_t bar_kv_full = bars + 144;                 // [kStages]
  const uint32_t bar_kv_empty = bars + 144 + 8 * kStages;  // [kStages]
  const uint32_t tmem_slot = bars + 144 + 16 * kStages;    // u32 written by tcgen05.alloc
  const uint32_t bar_pv_part = bars + 384;                 // [2]  MMA -> softmax (first part of a PV has landed)
  const uint32_t bar_s_free = bars + 400;                  // [2] softmax -> MMA (S copied into registers; 128 arrivals; shared-S map: one)
  const uint32_t bar_clc_full = bars + 416;                // [2]  CLC response landed (16 tx bytes)
  const uint32_t bar_clc_empty = bars + 432;               // [2]  all 14 consumer warps have read the response
  const uint32_t clc_resp = bars + 448;                    // [2] x 16 B
  const uint32_t s_inv_l = bars + 512;                     // [2 tiles][128 rows] fp32

  // warp index through a lane-0 shuffle: tells the compiler it is warp-uniform, so everything derived from it (tile
  // index, TMEM lane quarter, barrier addresses) can live in uniform registers instead of being recomputed from
  // threadIdx after every barrier wait
  const int warp = FA_UNIFORM_HINT ? __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0) : int(threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kClcConsumers = kTwoIssuers ? 15 : 14;   // producer lane + MMA warp(s) + 8 softmax warps + 4 epilogue warps
  // Every role walks the same item sequence: blockIdx.x first, then whatever the scheduler lane stole.
  // The response for item t+1 is requested at the start of item t into slot (t+1)&1 and read by every
  // consumer when it has finished item t.
  auto next_item = [&](int t_next, bool whole_warp = true) -> int {
    const uint32_t slot = uint32_t(t_next & 1);
    mbar_wait(bar_clc_full + 8 * slot, uint32_t((t_next - 1) >> 1) & 1u, 500);   // use k = (t_next-1)/2 of the slot
    const int id = clc_decode(clc_resp + 16 * slot);
    // same value in every lane: say so (not for the producer, which calls this from a single lane)
    return (FA_UNIFORM_HINT && whole_warp) ? __shfl_sync(0xffffffffu, id, 0) : id;
  };
  auto release_item_slot = [&](int t_next) {   // one arrival per consumer warp, after all its lanes have read
    mbar_arrive(bar_clc_empty + 8 * uint32_t(t_next & 1));
  };

  // ---- one-time setup
  if (warp == 12 && lane == 0) {
    prefetch_tensormap(&tmQ);
    prefetch_tensormap(&tmK);
    prefetch_tensormap(&tmV);
    prefetch_tensormap(&tmO);
    if (a.nsplit > 1) prefetch_tensormap(&tmW);
  }
  if (warp == 13 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_q_full + 8 * i, 1);
      mbar_init(bar_q_empty + 8 * i, 1);
      mbar_init(bar_s_full + 8 * i, 1);
      mbar_init(bar_o_full + 8 * i, 1);
      mbar_init(bar_pv_part + 8 * i, 1);
      mbar_init(bar_o_free + 8 * i, 128);
      mbar_init(bar_ep_full + 8 * i, 128);
      mbar_init(bar_ep_empty + 8 * i, 128);
      mbar_init(bar_p_full + 16 * i, 128);
      mbar_init(bar_p_full + 16 * i + 8, 128);
    }
    mbar_init(bar_s_free, 128);
    mbar_init(bar_s_free + 8, 128);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_kv_full + 8 * s, 1);
      mbar_init(bar_kv_empty + 8 * s, kTwoIssuers ? 2 : 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_clc_full + 8 * s, 1);
      mbar_init(bar_clc_empty + 8 * s, kClcConsumers);
    }
    fence_mbar_init();
  }
  if (warp == 14) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp >= 12) {
    setmaxnreg_dec<40>();
    if (warp == 12) {
      // =========================== TMA producer ===========================
      if (lane == 0) {
        int it = 0;                  // K/V ring position, runs across work items
        uint32_t nq[2] = {0u, 0u};   // Q_i loads issued so far (a tile is skipped when it has no K/V tile to visit)
        int w = blockIdx.x;
        for (int t = 0; w >= 0; ++t) {
          {  // scheduler: request the item after this one
            const uint32_t slot = uint32_t((t + 1) & 1);
            mbar_wait(bar_clc_empty + 8 * slot, (uint32_t(t >> 1) & 1u) ^ 1u, 120);   // use k = t/2 of the slot
            mbar_arrive_expect_tx(bar_clc_full + 8 * slot, 16);
            clc_try_cancel(clc_resp + 16 * slot, bar_clc_full + 8 * slot);
          }
          const WorkItem wi = get_item<kCausal>(a, w);
#ifdef FA_TRACE
          if (a.trace != nullptr && blockIdx.x < 160 && t < 40) {   // which CTA ran which item, and when
            unsigned long long now;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
            a.trace[4096 + (blockIdx.x * 40 + t) * 2] = w + 1;
            a.trace[4096 + (blockIdx.x * 40 + t) * 2 + 1] = (long long)now;
          }
#endif
          const int b_idx = wi.bh / a.H, h_idx = wi.bh - b_idx * a.H;
          auto load_tile = [&](const CUtensorMap* tm, unsigned int perm, uint32_t dst, uint32_t bar, int row0) {
            mbar_arrive_expect_tx(bar, kTileBytes);
#pragma unroll
            for (int h = 0; h < kNumBoxes; ++h)
              tma_load_tile(dst + h * kBoxBytes, tm, bar, h * kBoxCols, row0, h_idx, b_idx, perm);
          };
          // Q_i of the previous item must have been consumed by all of its QK^T MMAs
          if (wi.n_t0 > 0) {
            mbar_wait(bar_q_empty, (nq[0] & 1u) ^ 1u, 110);
            load_tile(&tmQ, a.perm_q, sQ, bar_q_full, wi.q0);
            ++nq[0];
          }
          for (int j = 0; j < wi.n_max; ++j) {
#pragma unroll
            for (int kv = 0; kv < 2; ++kv) {
              const int stage = it % kStages;
              const uint32_t ph = (it / kStages) & 1;
              mbar_wait(bar_kv_empty + 8 * stage, ph ^ 1, 100 + kv);
              load_tile(kv == 0 ? &tmK : &tmV, a.perm_kv, sKV + stage * kTileBytes, bar_kv_full + 8 * stage, (wi.kv_begin + j) * kBlockN);
              ++it;
              if (j == 0 && kv == 0 && wi.n_t1 > 0) {
                mbar_wait(bar_q_empty + 8, (nq[1] & 1u) ^ 1u, 111);
                load_tile(&tmQ, a.perm_q, sQ + kTileBytes, bar_q_full + 8, wi.q0 + kBlockM);
                ++nq[1];
              }
            }
          }
          w = next_item(t + 1, false);
          release_item_slot(t + 1);
        }
      }
      __syncwarp();
    } else if (warp == 13 || (kTwoIssuers && warp == 15)) {
      // =========================== MMA issuer(s) ===========================
      // The whole warp runs the (uniform) control flow and the barrier waits; one elected lane issues
      // the tcgen05.mma / tcgen05.commit instructions.  Keeping the flow warp-uniform lets the compiler
      // hold descriptors in uniform registers: the issue cost per MMA must stay well under the 64
      // cycles one 128x128x16 MMA takes on the tensor pipe.
      auto mbar_wait = [&](uint32_t bar, uint32_t parity, int tag) {
        if (FA_MMA_SPIN) fa::mbar_wait_spin(bar, parity, tag); else fa::mbar_wait(bar, parity, tag);
      };
      const uint32_t hi_qk = uint32_t(a.desc_hi_qk >> 32), hi_v = uint32_t(a.desc_hi_v >> 32);
      const uint32_t lo_qk = uint32_t(a.desc_hi_qk), lo_v = uint32_t(a.desc_hi_v);
      const uint32_t idesc_qk = a.idesc_qk, idesc_pv = a.idesc_pv;
      // S_i = Q_i K^T; then signal `bar_done`, and up to two more barriers (K stage release, Q_i release)
      auto issue_qk = [&](int i, int stage, uint32_t bar_done, uint32_t bar_rel_a, uint32_t bar_rel_b) {
        const uint32_t a_lo = lo_qk | ((sQ + i * kTileBytes) >> 4);
        const uint32_t b_lo = lo_qk | ((sKV + stage * kTileBytes) >> 4);
        const uint32_t d_tmem = tmem_base + (i ? kColS1 : 0u);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < ((FA_ABLATE & 64) ? D / 32 : D / 16); ++k) {
            const uint32_t off = ((k / 4) * kBoxBytes + (k % 4) * 32) >> 4;
            umma_ss(d_tmem, a_lo + off, hi_qk, b_lo + off, hi_qk, idesc_qk, k > 0 ? 1u : 0u);
          }
          umma_commit(bar_done);
          if (bar_rel_a) umma_commit(bar_rel_a);
          if (bar_rel_b) umma_commit(bar_rel_b);
        }
        __syncwarp();
      };
      // O_i += P_i V_j, issued in two parts (FA_P_FIRST_Q groups of 32 keys, then the rest): the first part starts as
      // soon as the softmax warpgroup has published those columns of P, while it is still exponentiating the rest.
      auto issue_pv = [&](int i, int stage, bool acc, uint32_t parity, uint32_t bar_done, uint32_t bar_release) {
        const uint32_t b_lo = lo_v | ((sKV + stage * kTileBytes) >> 4);
        const uint32_t p_tmem = tmem_base + (i ? kColP1 : kColP0);
        const uint32_t d_tmem = tmem_base + kColO + i * D;          // O_i
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(bar_p_full + 16 * i + 8 * h, parity, 212 + 2 * i + h);
          tc_fence_after();
          if (elect_one_sync()) {
#pragma unroll
            for (int k = (h ? 2 * FA_P_FIRST_Q : 0); k < (h ? 8 : 2 * FA_P_FIRST_Q); k += ((FA_ABLATE & 128) ? 2 : 1))
              umma_ts(d_tmem, p_tmem + k * 8, b_lo + k * (16 * kRowBytes / 16), hi_v, idesc_pv, (acc || k > 0) ? 1u : 0u);
            if constexpr (kPrecise) {   // O += P_lo V_j: the rounding residual of P, 64 columns further up
#pragma unroll
              for (int k = (h ? 2 * FA_P_FIRST_Q : 0); k < (h ? 8 : 2 * FA_P_FIRST_Q); ++k)
                umma_ts(d_tmem, p_tmem + kBlockN / 2 + k * 8, b_lo + k * (16 * kRowBytes / 16), hi_v, idesc_pv, 1u);
            }
            if (h == 0) {
              umma_commit(bar_pv_part + 8 * i);   // only waited for by the softmax warps' rare mid-tile rescale
            } else {
              umma_commit(bar_done);
              if (bar_release) umma_commit(bar_release);
            }
          }
          __syncwarp();
        }
      };
      auto stage_of = [&](int it) { return it % kStages; };
      auto phase_of = [&](int it) { return uint32_t((it / kStages) & 1); };

      if constexpr (kTwoIssuers) {
        // ---- one issuer per Q tile: warp 13 -> tile 0, warp 15 -> tile 1
        const int i = (warp == 15) ? 1 : 0;
        const uint32_t b_s_full = bar_s_full + 8 * i, b_o_full = bar_o_full + 8 * i, b_o_free = bar_o_free + 8 * i;
        const uint32_t b_q_full = bar_q_full + 8 * i, b_q_empty = bar_q_empty + 8 * i, b_s_free = bar_s_free + 8 * i;
        int it0 = 0;          // ring position of this item's K_0
        uint32_t cnt = 0;     // S/P/O steps of this tile so far (across items)
        uint32_t nq = 0;      // Q_i tiles consumed so far
        uint32_t ne = 0;      // epilogues of this tile started before this item
        // a stage this tile does not read still needs this issuer's arrival; arriving only after the stage has been seen
        // filled keeps the arrival in the right phase of kv_empty
        auto pass_stage = [&](int it) {
          mbar_wait(bar_kv_full + 8 * stage_of(it), phase_of(it), 240);
          if (lane == 0) mbar_arrive(bar_kv_empty + 8 * stage_of(it));
          __syncwarp();
        };
        // S_i of step `step` (counted across items) is in the softmax warpgroup's registers
        auto wait_s_copied = [&](uint32_t step) {
          if constexpr (kSepP) {
            mbar_wait(b_s_free, step & 1u, 231);
            tc_fence_after();
          }
        };
        int w = blockIdx.x;
        for (int t = 0; w >= 0; ++t) {
          const WorkItem wi = get_item<kCausal>(a, w);
          const int n_i = i ? wi.n_t1 : wi.n_t0, n_max = wi.n_max;
          const bool valid_i = i ? wi.valid1 : wi.valid0;
          if (n_max > 0) {
            if (n_i > 0) {      // S_i(0) = Q_i K_0^T
              mbar_wait(bar_kv_full + 8 * stage_of(it0), phase_of(it0), 200);
              mbar_wait(b_q_full, nq & 1u, 201 + i);
              tc_fence_after();
              if (cnt > 0) wait_s_copied(cnt - 1u);
              issue_qk(i, stage_of(it0), b_s_full, bar_kv_empty + 8 * stage_of(it0), n_i == 1 ? b_q_empty : 0u);
            } else {
              pass_stage(it0);
            }
          }
          for (int j = 0; j < n_max; ++j) {
            const int it_v = it0 + 2 * j + 1, it_k = it0 + 2 * j + 2;
            const bool has_next = (j + 1 < n_max);
            auto next_qk = [&]() {
              if (!has_next) return;
              if (j + 1 < n_i) {
                mbar_wait(bar_kv_full + 8 * stage_of(it_k), phase_of(it_k), 211);
                wait_s_copied(cnt + uint32_t(j));
                issue_qk(i, stage_of(it_k), b_s_full, bar_kv_empty + 8 * stage_of(it_k), j + 2 == n_i ? b_q_empty : 0u);
              } else {
                pass_stage(it_k);
              }
            };
            if constexpr (kSepP) next_qk();      // P has columns of its own: the score product does not wait for P V(j)
            if (j < n_i) {
              mbar_wait(bar_kv_full + 8 * stage_of(it_v), phase_of(it_v), 210);
              // the first PV of an item overwrites O_i: the epilogue must have read the previous item's O_i
              if (j == 0) mbar_wait(b_o_free, (ne & 1u) ^ 1u, 220 + i);
              issue_pv(i, stage_of(it_v), j > 0, (cnt + uint32_t(j)) & 1u, b_o_full, bar_kv_empty + 8 * stage_of(it_v));
            } else {
              pass_stage(it_v);
            }
            if constexpr (!kSepP) next_qk();     // P lies over S_i: Q_i K^T(j+1) follows P_i V(j)
          }
          it0 += 2 * n_max;
          cnt += uint32_t(n_i);
          nq += n_i > 0 ? 1u : 0u;
          ne += valid_i ? 1u : 0u;
          w = next_item(t + 1);
          __syncwarp();
          if (lane == 0) release_item_slot(t + 1);
        }
      } else {
      int it0 = 0;                    // ring position of this item's K_0
      uint32_t cnt[2] = {0u, 0u};     // S/P/O phase counters per tile (one per (tile, j) step, across items)
      uint32_t nq[2] = {0u, 0u};      // Q_i tiles consumed so far
      uint32_t ne[2] = {0u, 0u};      // epilogues of tile i started before this item (items where the tile is valid)
      uint32_t nqk = 0;               // shared-S protocol: score products issued so far (each one is followed by one s_free)
      // shared-S protocol: the buffer is free again once the softmax warpgroup that owns the previous S has it in registers
      auto wait_s_free = [&]() {
        if constexpr (kSharedS) {
          if (nqk > 0) {
            mbar_wait(bar_s_free, (nqk - 1u) & 1u, 230);
            tc_fence_after();
          }
          ++nqk;
        }
      };
      int w = blockIdx.x;
      for (int t = 0; w >= 0; ++t) {
        const WorkItem wi = get_item<kCausal>(a, w);
        const int n_t0 = wi.n_t0, n_t1 = wi.n_t1, n_max = wi.n_max;
        if (n_max > 0) {
          // S_i(0) = Q_i K_0^T
          mbar_wait(bar_kv_full + 8 * stage_of(it0), phase_of(it0), 200);
          const uint32_t rel = bar_kv_empty + 8 * stage_of(it0);
          if (n_t0 > 0) {
            mbar_wait(bar_q_full, nq[0] & 1u, 201);
            tc_fence_after();
            wait_s_free();
            issue_qk(0, stage_of(it0), bar_s_full, n_t1 > 0 ? 0u : rel, n_t0 == 1 ? bar_q_empty : 0u);
          }
          if (n_t1 > 0) {
            mbar_wait(bar_q_full + 8, nq[1] & 1u, 202);
            tc_fence_after();
            wait_s_free();
            issue_qk(1, stage_of(it0), bar_s_full + 8, rel, n_t1 == 1 ? bar_q_empty + 8 : 0u);
          }
        }
        for (int j = 0; j < n_max; ++j) {
          const int it_v = it0 + 2 * j + 1, it_k = it0 + 2 * j + 2;
          const bool has_next = (j + 1 < n_max);
          mbar_wait(bar_kv_full + 8 * stage_of(it_v), phase_of(it_v), 210);
          if (has_next) mbar_wait(bar_kv_full + 8 * stage_of(it_k), phase_of(it_k), 211);
          const uint32_t rel_v = bar_kv_empty + 8 * stage_of(it_v);
          const uint32_t rel_k = bar_kv_empty + 8 * stage_of(it_k);
          // n_t1 >= n_t0 whenever both tiles exist (the second tile sits lower in the causal triangle);
          // tile 1 is absent (n_t1 == 0) only for a ragged last block.
          if constexpr (kSharedS) {
            // QK^T of step j+1 goes first: it only needs the score buffer (the other tile's softmax has copied its S
            // out), not this tile's P V of step j, so S_i(j+1) is waiting when the softmax of S_i(j) ends.
            if (j + 1 < n_t0) {
              wait_s_free();
              issue_qk(0, stage_of(it_k), bar_s_full, j + 1 < n_t1 ? 0u : rel_k, j + 2 == n_t0 ? bar_q_empty : 0u);
            }
            if (j < n_t0) {
              if (j == 0) mbar_wait(bar_o_free, (ne[0] & 1u) ^ 1u, 220);
              issue_pv(0, stage_of(it_v), j > 0, (cnt[0] + j) & 1u, bar_o_full, j < n_t1 ? 0u : rel_v);
            }
            if (j + 1 < n_t1) {
              wait_s_free();
              issue_qk(1, stage_of(it_k), bar_s_full + 8, rel_k, j + 2 == n_t1 ? bar_q_empty + 8 : 0u);
            }
            if (j < n_t1) {
              if (j == 0) mbar_wait(bar_o_free + 8, (ne[1] & 1u) ^ 1u, 221);
              issue_pv(1, stage_of(it_v), j > 0, (cnt[1] + j) & 1u, bar_o_full + 8, rel_v);
            }
          } else {
            if (j < n_t0) {
              // the first PV of an item overwrites O_0: the epilogue must have read the previous item's O_0
              if (j == 0) mbar_wait(bar_o_free, (ne[0] & 1u) ^ 1u, 220);
              issue_pv(0, stage_of(it_v), j > 0, (cnt[0] + j) & 1u, bar_o_full, j < n_t1 ? 0u : rel_v);
            }
            if (j + 1 < n_t0)
              issue_qk(0, stage_of(it_k), bar_s_full, j + 1 < n_t1 ? 0u : rel_k, j + 2 == n_t0 ? bar_q_empty : 0u);
            if (j < n_t1) {
              if (j == 0) mbar_wait(bar_o_free + 8, (ne[1] & 1u) ^ 1u, 221);
              issue_pv(1, stage_of(it_v), j > 0, (cnt[1] + j) & 1u, bar_o_full + 8, rel_v);
            }
            if (j + 1 < n_t1)
              issue_qk(1, stage_of(it_k), bar_s_full + 8, rel_k, j + 2 == n_t1 ? bar_q_empty + 8 : 0u);
          }
        }
        it0 += 2 * n_max;
        cnt[0] += uint32_t(n_t0);
        cnt[1] += uint32_t(n_t1);
        nq[0] += n_t0 > 0 ? 1u : 0u;
        nq[1] += n_t1 > 0 ? 1u : 0u;
        ne[0] += wi.valid0 ? 1u : 0u;
        ne[1] += wi.valid1 ? 1u : 0u;
        w = next_item(t + 1);
        __syncwarp();
        if (lane == 0) release_item_slot(t + 1);
      }
      }   // one issuer
    }
  } else if (warp >= 8) {
    // =========================== correction / epilogue warpgroup ===========================
    // Final correction of the accumulator: O_i * (1/l) -> 16-bit -> 128B-swizzled staging tile -> TMA store,
    // running concurrently with the next work item.  1/l arrives from the softmax warpgroup through shared
    // memory; O_i is released back to the MMA warp as soon as it is in registers.
    setmaxnreg_dec<56>();
    const int wl = warp & 3;
    const int row_in_tile = wl * 32 + lane;
    const uint32_t lane_addr = uint32_t(wl * 32) << 16;
    uint32_t cnt[2] = {0u, 0u};
    uint32_t ne[2] = {0u, 0u};      // hand-offs of tile i received so far
    bool store_pending = false;
    int w = blockIdx.x;
    for (int t = 0; w >= 0; ++t) {
      const WorkItem wi = get_item<kCausal>(a, w);
      w = next_item(t + 1);     // read early: the slot is only recycled once every warp has let go of it
      __syncwarp();
      if (lane == 0) release_item_slot(t + 1);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (!(i ? wi.valid1 : wi.valid0)) continue;
        const int n_i = i ? wi.n_t1 : wi.n_t0;
        const uint32_t tO = tmem_base + lane_addr + kColO + i * D;
#if FA_EP_NAMED_BAR
        // the epilogue warpgroup waits for almost a whole work item here: on a hardware named barrier (the softmax
        // warpgroup arrives, this one syncs) it does not poll, so it leaves its SMSP's issue slots to the softmax warps
        asm volatile("bar.sync %0, %1;" ::"r"(3 + i), "r"(256) : "memory");
#else
        mbar_wait(bar_ep_full + 8 * i, ne[i] & 1u, 400 + i);
#endif
        ++ne[i];
        float inv_l;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(inv_l) : "r"(s_inv_l + uint32_t(i * 128 + row_in_tile) * 4u) : "memory");
        mbar_arrive(bar_ep_empty + 8 * i);
        if (n_i > 0) {
          mbar_wait(bar_o_full + 8 * i, (cnt[i] + uint32_t(n_i) - 1u) & 1u, 410 + i);
          tc_fence_after();
        }
        // staging tile free again?  (the previous TMA store must have finished reading it)
        if (store_pending) {
          if (row_in_tile == 0) tma_store_wait_read<0>();
          named_bar_sync(2, 128);
        }
#pragma unroll
        for (int q = 0; q < D / 32; ++q) {
          uint32_t orow[32];
          if (n_i > 0) {
            tmem_ld32(tO + q * 32, orow);
            tmem_wait_ld();
            if (q == D / 32 - 1) {        // O_i is in registers: hand the accumulator back to the MMA warp
              tc_fence_before();
              mbar_arrive(bar_o_free + 8 * i);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) orow[k] = 0u;
            if (q == D / 32 - 1) mbar_arrive(bar_o_free + 8 * i);
          }
          // 16-byte chunk (q*4 + v) of this row: box (chunk / chunks-per-row), swizzled inside the box row
          constexpr uint32_t kChunksPerRow = kRowBytes / 16;                 // 8 (128B swizzle) or 4 (64B swizzle)
          const uint32_t swz = (kRowBytes == 128) ? uint32_t(row_in_tile & 7) : uint32_t((row_in_tile >> 1) & 3);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint32_t wv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              wv[e] = pack2<kBF16>(__uint_as_float(orow[v * 8 + 2 * e]) * inv_l,
                                   __uint_as_float(orow[v * 8 + 2 * e + 1]) * inv_l);
            const uint32_t c16 = uint32_t(q * 4 + v);
            const uint32_t addr = sO + (c16 / kChunksPerRow) * kBoxBytes + row_in_tile * kRowBytes +
                                  (((c16 % kChunksPerRow) ^ swz) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(wv[0]), "r"(wv[1]),
                         "r"(wv[2]), "r"(wv[3]) : "memory");
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (row_in_tile == 0) {
#pragma unroll
          for (int h = 0; h < kNumBoxes; ++h) {
            if (wi.split_out)     // a partial: workspace tile (split, item, rows i*128..) instead of the caller's O
              tma_store_tile(&tmW, sO + h * kBoxBytes, h * kBoxCols, i * kBlockM, wi.ws_item, wi.split, a.perm_w);
            else
              tma_store_tile(&tmO, sO + h * kBoxBytes, h * kBoxCols, wi.q0 + i * kBlockM, wi.bh % a.H, wi.out_b, a.perm_o);
          }
          tma_store_commit();
        }
        store_pending = true;
        cnt[i] += uint32_t(n_i);
      }
    }
    if (store_pending && row_in_tile == 0) tma_store_wait<0>();
  } else {
    // =========================== softmax warpgroups ===========================
    setmaxnreg_inc<208>();
    const int i = warp >> 2;           // which Q tile
    const int wl = warp & 3;           // warp within the warpgroup == TMEM lane quarter
    const int row_in_tile = wl * 32 + lane;
    const uint32_t lane_addr = uint32_t(wl * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + (i ? kColS1 : 0u);
    const uint32_t tP = tmem_base + lane_addr + (i ? kColP1 : kColP0);   // 16-bit P, 64 columns (+ 64 of P_lo in precise mode)
    const uint32_t tO = tmem_base + lane_addr + kColO + i * D;
    const uint32_t b_s_full = bar_s_full + 8 * i;
    const uint32_t b_p_full = bar_p_full + 16 * i;
    const uint32_t b_o_full = bar_o_full + 8 * i;
    const float c = a.scale_log2;
    uint32_t cnt = 0;                  // S/P/O phase counter of this tile, across items
    uint32_t ne = 0;                   // hand-offs to the epilogue warpgroup so far
    int w = blockIdx.x;
    for (int t = 0; w >= 0; ++t) {
      const int w_cur = w;
      int n_i, limit, kv_begin;
      bool valid;
      {
        const WorkItem wi = get_item<kCausal>(a, w_cur);
        valid = i ? wi.valid1 : wi.valid0;
        n_i = i ? wi.n_t1 : wi.n_t0;
        kv_begin = wi.kv_begin;
        // last visible key index for this row
        limit = a.Nkv - 1;
        if (kCausal) limit = min(limit, wi.q0 + i * kBlockM + row_in_tile + a.causal_off);
      }
      w = next_item(t + 1);     // read early: the slot is only recycled once every warp has let go of it
      __syncwarp();
      if (lane == 0) release_item_slot(t + 1);
      if (!valid) continue;

      float m_ref = -INFINITY;   // max used in the exponents (lags the true max by <= threshold)
      float m_true = -INFINITY;  // true running row max (scaled, log2 units)
      float l = 0.f;             // sum of 2^(t - m_ref)

      for (int j = 0; j < n_i; ++j) {
        const uint32_t par = (cnt + uint32_t(j)) & 1u;
        mbar_wait(b_s_full, par, 300 + i);
        tc_fence_after();
        if (row_in_tile == 0 && t == 0) FA_TRACE_EV(j, 4 * i + 0);
        uint32_t sr[4][32];
#pragma unroll
        for (int q = 0; q < 4; ++q) tmem_ld32(tS + q * 32, sr[q]);
        tmem_wait_ld();
        if constexpr (kSharedS || kSepP) {
          // S is in registers: the score buffer may take the next score product (shared-S map: the other tile's) ...
          tc_fence_before();
          mbar_arrive(bar_s_free + (kSepP ? 8 * i : 0));
          // ... and before P of this step overwrites P of the previous one, P V of the previous step must have read it
          // (with P written over its own S tile this was implied by S(j) being there at all)
          if (cnt + uint32_t(j) > 0u) {
            mbar_wait(b_o_full, par ^ 1u, 340 + i);
            tc_fence_after();
          }
        }
        if (row_in_tile == 0 && t == 0) FA_TRACE_EV(j, 4 * i + 1);

        const float2 c2 = make_float2(c, c);
        // p = 2^(s*c + neg_m) for the 32 keys of group q, with packed (2-wide) fp32 math; FA_EMU_PAIRS_OF_4 of every 4
        // pairs take the polynomial path, the rest MUFU.EX2.  P (16-bit) goes to tP (over S_i, or columns of its own).
        auto exp_group = [&](int q, const float2 neg_m2, float2 (&lsum2)[4], const float2 neg_late2, int late_from = 16) {
          uint32_t pk[16];
          [[maybe_unused]] uint32_t pl[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            float2 x = make_float2(__uint_as_float(sr[q][2 * k]), __uint_as_float(sr[q][2 * k + 1]));
            if (!(FA_ABLATE & 2)) x = __ffma2_rn(x, c2, k < late_from ? neg_m2 : neg_late2);
            float2 pv;
            if (D <= 64 ? ((k & 7) < FA_EMU_PAIRS_OF_8_D64) : ((k & 3) < FA_EMU_PAIRS_OF_4)) {
              pv = ex2_emulated(x);
            } else if ((FA_ABLATE & 4) && (k & 1)) {
              pv = x;
            } else {
              pv.x = ex2_approx(x.x);
              pv.y = ex2_approx(x.y);
            }
            if (!(FA_ABLATE & 1)) lsum2[k & 3] = __fadd2_rn(lsum2[k & 3], pv);
            else if (k == 0) lsum2[0] = __fadd2_rn(lsum2[0], pv);
            pk[k] = (FA_ABLATE & 32) ? (__float_as_uint(pv.x) ^ (__float_as_uint(pv.y) >> 16)) : pack2<kBF16>(pv);
            if constexpr (kPrecise) {
              const float2 hi = unpack2<kBF16>(pk[k]);
              pl[k] = pack2<kBF16>(__fadd2_rn(pv, make_float2(-hi.x, -hi.y)));
            }
          }
          tmem_st16(tP + q * 16, pk);
          if constexpr (kPrecise) tmem_st16(tP + kBlockN / 2 + q * 16, pl);   // P_lo over columns 64..127
        };
        auto publish = [&](int part, bool already) {   // P columns written so far are visible to the MMA warp
          if (!(FA_ABLATE & 16)) tmem_wait_st();
          tc_fence_before();
          if (!already) mbar_arrive(b_p_full + 8 * part);
          if (row_in_tile == 0 && t == 0) FA_TRACE_EV(j, 4 * i + 2 + part);
        };
        auto total = [](const float2 (&v)[4]) {
          const float2 r = __fadd2_rn(__fadd2_rn(v[0], v[1]), __fadd2_rn(v[2], v[3]));
          return r.x + r.y;
        };
        // O_i (this warp's 32 rows) *= alpha, in TMEM
        auto rescale_o = [&](float alpha) {
#pragma unroll
          for (int q = 0; q < D / 32; ++q) {
            uint32_t orow[32];
            tmem_ld32(tO + q * 32, orow);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 32; ++k) orow[k] = __float_as_uint(__uint_as_float(orow[k]) * alpha);
            tmem_st32(tO + q * 32, orow);
          }
        };
        auto group_max = [&](int q0, int q1) {   // four independent chains (a chain of dependent FMNMX3 costs ~4 clk each)
          float mxp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int q = q0; q < q1; ++q)
#pragma unroll
            for (int k = 0; k < 32; k += 8)
#pragma unroll
              for (int u = 0; u < 4; ++u)
                mxp[u] = fmaxf(mxp[u], fmaxf(__uint_as_float(sr[q][k + 2 * u]), __uint_as_float(sr[q][k + 2 * u + 1])));
          return fmaxf(fmaxf(mxp[0], mxp[1]), fmaxf(mxp[2], mxp[3]));
        };

        // masking: key index > limit -> -inf (diagonal tiles of causal runs, ragged last tile)
        const int lim_local = limit - (kv_begin + j) * kBlockN;
        const bool masked = __any_sync(0xffffffffu, lim_local < kBlockN - 1);

        // ---- fast path: no row-max pass, no branch in the common case.  The tile is exponentiated against the
        // reference max it inherited, and each thread hands a part of P to the MMA warp only if the part's row sum proves
        // every p of the part finite and harmless (p <= sum <= 2^15) - a predicated arrival, so the instruction stream
        // stays straight.  Only when a row of the warp failed (one vote at the end of the tile) is anything redone:
        //   * a row failed in the first part: no PV MMA of this tile can have started (its arrival is missing), so the
        //     whole tile goes through the exact path below, which arrives for the threads that have not yet;
        //   * rows failed in the second part only: wait for the first part's PV MMAs, rescale O and l to the new max,
        //     exponentiate the second part again.
        bool done = false;
        bool arrived0 = false, arrived1 = false;   // this thread's arrivals on p_full[0 / 1] for this tile
        if (FA_FAST_SOFTMAX && D >= FA_FAST_MIN_D && !kPrecise && j > 0 && !masked && !a.need_stats) {
          float2 neg_m2 = make_float2(-m_ref, -m_ref);     // finite: the row saw a whole unmasked tile before
          float2 ls0[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
          float2 ls1[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
          for (int q = 0; q < FA_P_FIRST_Q; ++q) exp_group(q, neg_m2, ls0, neg_m2);
          const float sum0 = total(ls0);
          arrived0 = sum0 <= kFastSumLimit;
          tmem_wait_st();
          tc_fence_before();
          // The arrival has no consumer, so the compiler would let it sink below all of the second part's exponentials
          // (and the first PV MMAs would start that much later).  Its token is folded, as an opaque +0.0, into the
          // reference value all but the first four pairs of the second part use: those four pairs (64 clk of MUFU)
          // cover the store wait, the rest has to follow the arrival.
          const uint32_t tok = mbar_arrive_if_token(b_p_full, arrived0);
          const float neg_late = -m_ref + __uint_as_float(tok & a.zero);
          const float2 neg_late2 = make_float2(neg_late, neg_late);
          if (row_in_tile == 0 && t == 0) FA_TRACE_EV(j, 4 * i + 2);
#pragma unroll
          for (int q = FA_P_FIRST_Q; q < 4; ++q) exp_group(q, neg_m2, ls1, neg_late2, q == FA_P_FIRST_Q ? 4 : 0);
          float sum1 = total(ls1);
          arrived1 = arrived0 && (sum1 <= kFastSumLimit);
          tmem_wait_st();
          tc_fence_before();
          if (arrived1) mbar_arrive(b_p_full + 8);
          if (row_in_tile == 0 && t == 0) FA_TRACE_EV(j, 4 * i + 3);
          if (!__any_sync(0xffffffffu, !arrived1)) {
            l += sum0 + sum1;
            done = true;
          } else if (!__any_sync(0xffffffffu, !arrived0)) {
            // second part only
            const float m_new = arrived1 ? m_ref : fmaxf(m_ref, group_max(FA_P_FIRST_Q, 4) * c);
            m_true = fmaxf(m_true, m_new);
            const float alpha = ex2_approx(m_ref - m_new);
            m_ref = m_new;
            l = (l + sum0) * alpha;
            mbar_wait(bar_pv_part + 8 * i, par, 320 + i);   // the first part's MMAs have landed in O_i
            tc_fence_after();
            rescale_o(alpha);
            neg_m2 = make_float2(-m_ref, -m_ref);
#pragma unroll
            for (int u = 0; u < 4; ++u) ls1[u] = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = FA_P_FIRST_Q; q < 4; ++q) exp_group(q, neg_m2, ls1, neg_m2);
            l += total(ls1);
            tmem_wait_st();
            tc_fence_before();
            if (!arrived1) mbar_arrive(b_p_full + 8);
            done = true;
          }
        }

        if (!done) {
          // ---- exact path: mask, row max, lazy rescale, exponentials
          if (masked) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
              for (int k = 0; k < 32; ++k)
                if (q * 32 + k > lim_local) sr[q][k] = 0xff800000u;  // -inf
          }
          const float m_new = (FA_ABLATE & 8) ? fmaxf(m_true, __uint_as_float(sr[0][0]) * c) : fmaxf(m_true, group_max(0, 4) * c);
          m_true = m_new;
          if (row_in_tile == 0 && t == 0) FA_TRACE_EV(j, 14 + i);

          if (j == 0) {
            m_ref = m_new;
          } else {
            const bool moved = (m_new - m_ref) > kRescaleThreshold;
            if (__any_sync(0xffffffffu, moved)) {
              // rescale this warp's 32 rows of O (and l) to the new reference max
              const float alpha = (m_new == -INFINITY) ? 1.f : ex2_approx(m_ref - m_new);
              m_ref = m_new;
              l *= alpha;
              mbar_wait(b_o_full, par ^ 1u, 310 + i);  // PV(j-1) has landed in TMEM
              tc_fence_after();
              rescale_o(alpha);
            }
          }
          const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
          const float2 neg_m2 = make_float2(neg_m, neg_m);
          float2 lsum2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            exp_group(q, neg_m2, lsum2, neg_m2);
            if (q == FA_P_FIRST_Q - 1) publish(0, arrived0);
            if (q == 3) publish(1, arrived1);
          }
          l += total(lsum2);
        }
      }
      cnt += uint32_t(n_i);

      // ---- hand the row normaliser to the epilogue warpgroup, write the row statistics, move on
      if (limit < kv_begin * kBlockN) l = 0.f;   // row sees no key of this item: the polynomial exp2 returns 2^-127, not 0
      const float inv_l = (l > 0.f) ? (1.f / l) : 0.f;
      mbar_wait(bar_ep_empty + 8 * i, (ne & 1u) ^ 1u, 330 + i);   // slot of the previous hand-off consumed
      ++ne;
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_inv_l + uint32_t(i * 128 + row_in_tile) * 4u), "f"(inv_l) : "memory");
#if FA_EP_NAMED_BAR
      asm volatile("bar.arrive %0, %1;" ::"r"(3 + i), "r"(256) : "memory");
#else
      mbar_arrive(bar_ep_full + 8 * i);
#endif
      const WorkItem wi = get_item<kCausal>(a, w_cur);   // recomputed here to keep it out of the hot loop's registers
      const int row = wi.q0 + i * kBlockM + row_in_tile;
      if (row < a.Nq) {
        const float ln2 = 0.6931471805599453f;
        const bool any = l > 0.f;
        if (wi.split_out) {
          const long long off = ((long long)wi.split * a.num_ws_items + wi.ws_item) * (2 * kBlockM) + i * kBlockM + row_in_tile;
          a.ws_lse[off] = any ? fmaf(m_ref, ln2, logf(l)) : -INFINITY;
          a.ws_m[off] = any ? m_true * ln2 : -INFINITY;
        } else {
          const long long off = (long long)wi.out_b * a.stat_stride_b + (long long)(wi.bh % a.H) * a.stat_stride_h + row;
          if (a.lse) a.lse[off] = any ? fmaf(m_ref, ln2, logf(l)) : -INFINITY;
          if (a.m) a.m[off] = any ? m_true * ln2 : -INFINITY;
          if (a.l) a.l[off] = any ? l * ex2_approx(m_ref - m_true) : 0.f;
        }
      }
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 14) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace fa
