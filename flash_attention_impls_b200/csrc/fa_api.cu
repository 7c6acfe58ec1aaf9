// fa_api.cu — host layer behind the C ABI of include/fa_b200.h.
//
// Replaces the reference's host launch / dispatch level:
//   code/cuda_fa1/main.cu:377-385 (grid/block/shmem math + raw <<<>>> launch)
//   code/cutlass_cuda_fa1/run/flash_attn_cutlass.cu:457-544 (templated launcher + head_dim switch)
//   code/cutlass_cuda_fa1/run/flash_attn_unified.cu:545-617
// Validates arguments, encodes the four TMA tensor maps, picks the kernel instantiation and
// enqueues it.  Allocates nothing, never synchronises, returns a status code.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/fa_b200.h"
#include "fa_fwd_sm100.cuh"
#include "fa_bwd_sm100.cuh"

namespace fa {
int launch_item_combine(const void* O_part, const float* lse_part, const float* m_part, void* O, float* lse, float* l, float* m,
                        const FwdArgs& a, int d, long long o_sb, long long o_sh, long long o_sn, long long st_sb, long long st_sh,
                        int dtype, bool causal, cudaStream_t stream);
int launch_split_combine(const void* O_part, const float* lse_part, const float* m_part, void* O, float* lse,
                         float* l, float* m, int nsplit, long long rows, int d, int H, int N, long long o_sb,
                         long long o_sh, long long o_sn, long long st_sb, long long st_sh, int dtype,
                         cudaStream_t stream);
}

namespace {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  static const bool verbose = getenv("FA_B200_VERBOSE") != nullptr;   // read once per process
  if (verbose) fprintf(stderr, "fa_b200: %s\n", g_err);
  return code;
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// [B, H, rows, d] tensor view with element strides (sb, sh, sn), d contiguous, described as a 4-D TMA tensor
// {d, a1, a2, a3}.  The three outer axes (rows, heads, batch) are sorted by stride so the map's strides ascend
// (a [B,N,H,d] view has head stride < row stride); *perm records, 2 bits per axis, which of row (0) / head (1) /
// batch (2) each outer axis carries, and the kernel orders its coordinates accordingly.  The box is
// 64 columns x 128 rows x 1 x 1 with the 128-byte swizzle the UMMA descriptors expect (d = 32: 32 columns, 64-byte
// swizzle).  Out-of-range rows read as zero and are clipped on store, which is what makes ragged N work.
int encode_tmap(CUtensorMap* tm, unsigned* perm, const void* base, int dtype, int d, long long rows, long long H,
              long long B, long long sn, long long sh, long long sb) {
  // `d` is the tensor's real head_dim; the kernel instantiation (and with it the box) is the next of 32 / 64 / 128.
  // Columns of a box beyond d are out of bounds: TMA reads them as zeros and clips them on store, which is all a
  // head_dim between the instantiated sizes needs (zero columns add nothing to q.k and produce zero output columns).
  const int dk = d <= 32 ? 32 : (d <= 64 ? 64 : 128);
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(FA_B200_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  struct Axis { long long size, stride; unsigned role; cuuint32_t box; };
  Axis ax[3] = {{rows, sn, 0u, 128u}, {H, sh, 1u, 1u}, {B, sb, 2u, 1u}};
  // insertion sort by stride; size-1 axes go last (their stride is irrelevant and may be anything)
  auto key = [](const Axis& x) { return x.size == 1 ? (1LL << 62) : x.stride; };
  for (int i = 1; i < 3; ++i)
    for (int j = i; j > 0 && key(ax[j]) < key(ax[j - 1]); --j) std::swap(ax[j], ax[j - 1]);
  cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)ax[0].size, (cuuint64_t)ax[1].size, (cuuint64_t)ax[2].size};
  cuuint64_t strides[3];
  cuuint32_t box[4] = {(cuuint32_t)(dk >= 64 ? 64 : 32), ax[0].box, ax[1].box, ax[2].box};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  unsigned long long prev_extent = (unsigned long long)d * 2;
  *perm = 0;
  for (int i = 0; i < 3; ++i) {
    // a size-1 axis never advances: give it any legal stride (the extent so far)
    unsigned long long st = ax[i].size == 1 ? prev_extent : (unsigned long long)ax[i].stride * 2;
    strides[i] = st;
    prev_extent = std::max(prev_extent, st * (unsigned long long)ax[i].size);
    *perm |= ax[i].role << (2 * i);
  }
  CUresult r = enc(tm, dtype == FA_B200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                   4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   dk >= 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FA_B200_ERR_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return FA_B200_OK;
}

// Tensor maps are pure functions of (pointer, shape, strides, dtype), and callers launch the same tensors over and
// over (a training loop, the ring driver's recv slots, the host pipeline's staging buffers), so the encoded maps are
// kept in a small per-thread cache: 32 entries, round-robin replacement, no locking.  A hit
// costs a scan of at most 32 keys instead of a driver call; on a 67 us launch (BASELINE c2) four encodes are
// measurable host time.  The map only holds address arithmetic, so a freed-and-reallocated buffer at the same
// address with the same geometry is still described correctly.
struct TmapKey {
  const void* base;
  long long rows, H, B, sn, sh, sb;
  int dtype, d;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && H == o.H && B == o.B && sn == o.sn && sh == o.sh && sb == o.sb &&
           dtype == o.dtype && d == o.d;
  }
};
struct TmapEntry { TmapKey key; CUtensorMap map; unsigned perm; bool valid; };
constexpr int kTmapCacheSize = 32;
std::atomic<uint64_t> g_tmap_hits{0}, g_tmap_misses{0};

int make_tmap(CUtensorMap* tm, unsigned* perm, const void* base, int dtype, int d, long long rows, long long H,
              long long B, long long sn, long long sh, long long sb) {
  thread_local TmapEntry cache[kTmapCacheSize] = {};
  thread_local int next = 0;
  const TmapKey key{base, rows, H, B, sn, sh, sb, dtype, d};
  for (int i = 0; i < kTmapCacheSize; ++i)
    if (cache[i].valid && cache[i].key == key) {
      *tm = cache[i].map;
      *perm = cache[i].perm;
      g_tmap_hits.fetch_add(1, std::memory_order_relaxed);
      return FA_B200_OK;
    }
  const int rc = encode_tmap(tm, perm, base, dtype, d, rows, H, B, sn, sh, sb);
  if (rc) return rc;
  g_tmap_misses.fetch_add(1, std::memory_order_relaxed);
  TmapEntry& e = cache[next];
  next = (next + 1) % kTmapCacheSize;
  e.key = key; e.map = *tm; e.perm = *perm; e.valid = true;
  return FA_B200_OK;
}

int check_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(FA_B200_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  static std::atomic<int> cc_cache[64];
  int cc = (dev < 64) ? cc_cache[dev].load() : 0;
  if (cc == 0) {
    int major = 0, minor = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (e != cudaSuccess) return fail(FA_B200_ERR_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    cc = major * 10 + minor;
    if (dev < 64) cc_cache[dev].store(cc);
  }
  if (cc / 10 != 10)
    return fail(FA_B200_ERR_ARCH, "device %d is sm_%d; this library only runs on sm_100 (B200) and has no fallback",
                dev, cc);
  return FA_B200_OK;
}


template <int D, bool kBF16, bool kCausal, bool kPrecise>
int launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to, const CUtensorMap& tw,
           const fa::FwdArgs& args, long long grid, cudaStream_t stream) {
  auto kern = fa::fa_fwd_sm100_kernel<D, kBF16, kCausal, kPrecise>;
  constexpr int smem = fa::FwdTraits<D>::kSmemBytes;
  // opt-in to > 48 KB dynamic shared memory: once per device per instantiation
  static std::atomic<unsigned long long> configured{0};
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(configured.load() & bit)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(FA_B200_ERR_CUDA, "cudaFuncSetAttribute(smem=%d): %s", smem, cudaGetErrorString(e));
    configured.fetch_or(bit);
  }
#ifdef FA_TRACE
  // bring-up only: dump the clock64 timeline of CTA 0 to $FA_B200_TRACE after a synchronous run
  fa::FwdArgs targs = args;
  static long long* trace_dev = nullptr;
  const char* trace_path = getenv("FA_B200_TRACE");
  if (trace_path) {
    const size_t trace_n = 4096 + 160 * 40 * 2;
    if (!trace_dev) cudaMalloc(&trace_dev, trace_n * sizeof(long long));
    cudaMemsetAsync(trace_dev, 0, trace_n * sizeof(long long), stream);
    targs.trace = trace_dev;
  }
  kern<<<dim3((unsigned)grid), dim3(fa::kNumThreads), smem, stream>>>(tq, tk, tv, to, tw, targs);
  if (trace_path) {
    static long long h[4096 + 160 * 40 * 2];
    cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(trace_path, "w")) {
      for (int j = 0; j < 32; ++j) {
        for (int ev = 0; ev < 16; ++ev) fprintf(f, "%lld ", h[j * 16 + ev]);
        fprintf(f, "\n");
      }
      fprintf(f, "#items cta t item globaltimer_ns\n");
      for (int c = 0; c < 160; ++c)
        for (int t = 0; t < 40; ++t)
          if (h[4096 + (c * 40 + t) * 2]) fprintf(f, "I %d %d %lld %lld\n", c, t, h[4096 + (c * 40 + t) * 2] - 1, h[4096 + (c * 40 + t) * 2 + 1]);
      fclose(f);
    }
  }
#else
  kern<<<dim3((unsigned)grid), dim3(fa::kNumThreads), smem, stream>>>(tq, tk, tv, to, tw, args);
#endif
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(FA_B200_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1);
  return FA_B200_OK;
}

template <int D, bool kBF16, bool kCausal>
int launch_bwd_dq(const CUtensorMap& tq, const CUtensorMap& tdo, const CUtensorMap& tk, const CUtensorMap& tv,
                  const CUtensorMap& tdq, const fa::BwdArgs& args, long long grid, cudaStream_t stream) {
  auto kern = fa::fa_bwd_dq_sm100_kernel<D, kBF16, kCausal>;
  constexpr int smem = fa::BwdTraits<D>::kSmem2Bytes;
  static std::atomic<unsigned long long> configured{0};
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(configured.load() & bit)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(FA_B200_ERR_CUDA, "cudaFuncSetAttribute(bwd dq, smem=%d): %s", smem, cudaGetErrorString(e));
    configured.fetch_or(bit);
  }
  kern<<<dim3((unsigned)grid), dim3(fa::kBwdThreads), smem, stream>>>(tq, tdo, tk, tv, tdq, args);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(FA_B200_ERR_CUDA, "backward dQ kernel launch: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1);
  return FA_B200_OK;
}

template <int D, bool kBF16, bool kCausal>
int launch_bwd_dkdv(const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tq, const CUtensorMap& tdo,
                    const CUtensorMap& tdk, const CUtensorMap& tdv, const fa::BwdArgs& args, long long grid, cudaStream_t stream) {
  auto kern = fa::fa_bwd_dkdv_sm100_kernel<D, kBF16, kCausal>;
  constexpr int smem = fa::BwdTraits<D>::kSmem2Bytes;
  static std::atomic<unsigned long long> configured{0};
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(configured.load() & bit)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(FA_B200_ERR_CUDA, "cudaFuncSetAttribute(bwd dkdv, smem=%d): %s", smem, cudaGetErrorString(e));
    configured.fetch_or(bit);
  }
  kern<<<dim3((unsigned)grid), dim3(fa::kBwdThreads), smem, stream>>>(tk, tv, tq, tdo, tdk, tdv, args);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(FA_B200_ERR_CUDA, "backward dK/dV kernel launch: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1);
  return FA_B200_OK;
}

// FA_B200_GROUP_HEADS (tuning knob of the causal item order; 0 = the L2-sized default) is read once per process.
std::atomic<long long> g_group_heads{-1};   // -1: not initialised yet
long long group_heads_override() {
  long long v = g_group_heads.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* s = getenv("FA_B200_GROUP_HEADS");
    v = s ? std::max(0LL, strtoll(s, nullptr, 0)) : 0LL;
    g_group_heads.store(v, std::memory_order_relaxed);
  }
  return v;
}

// Bring-up overrides (UMMA descriptors, split count) exist only in -DFA_B200_DEBUG builds: a stray environment
// variable must not be able to corrupt a descriptor in the shipped library.
#ifdef FA_B200_DEBUG
unsigned long long env_u64(const char* name, unsigned long long dflt) {
  const char* s = getenv(name);
  return s ? strtoull(s, nullptr, 0) : dflt;
}
#endif

// Split-KV policy for a launch of `items` work items of `n_kv_tiles` K/V tiles each.  Returns the number of key-axis
// splits and sets *split_begin to the list position from which items are split:
//   * far fewer items than SMs (the reference's (1,1,N,64) sweep): every item is split so the machine fills once;
//   * non-causal launches of equal items whose last wave fills at most half of the SMs (512 items on 148 SMs = 3.46
//     waves: c3 sharded over 8 GPUs, 0.433 -> 0.406 ms) and whose items are at least 24 K/V tiles long: only that tail
//     is split, so it costs 1/nsplit of a round instead of a whole one.  (Causal items are ordered longest-first, their
//     tail is short already; at BASELINE c2's 8 tiles per item the combine launch costs more than the split saves:
//     0.0661 -> 0.0680 ms, profiles/r02_tail_split.log.)
int choose_nsplit(long long items, int n_kv_tiles, int sms, bool causal, long long* split_begin) {
  *split_begin = 0;
  long long n = 1;
  if (n_kv_tiles >= 8) {
    if (items * 2 <= sms) {
      n = std::min<long long>(std::min<long long>(sms / items, n_kv_tiles / 4), 32);   // >= 4 tiles per split
    } else if (!causal && items > sms && n_kv_tiles >= 24) {   // shorter items: the combine launch costs what the split saves
      const long long tail = items % sms;
      if (tail > 0 && tail * 2 <= sms) {
        n = std::min<long long>(std::min<long long>(sms / tail, n_kv_tiles / 4), 8);
        if (n > 1) *split_begin = items - tail;
      }
    }
  }
#ifdef FA_B200_DEBUG
  n = (long long)env_u64("FA_B200_NSPLIT", (unsigned long long)n);
#endif
  n = std::max<long long>(1, std::min<long long>(n, n_kv_tiles));
  if (n == 1) *split_begin = 0;
  return (int)n;
}

// partial O (16-bit) + lse + m for `ws_items` items of 256 rows, `nsplit` times
size_t split_workspace_bytes(int nsplit, long long ws_items, int d) {
  return nsplit <= 1 ? 0 : (size_t)nsplit * ws_items * 2 * fa::kBlockM * ((size_t)d * 2 + 2 * sizeof(float));
}

// SMs of the current device (queried once per device; 148 on B200, which is also the answer when no device is
// visible - the host-only introspection calls must work on a CPU box).
int sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int n = (dev < 64) ? cache[dev].load() : 0;
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (dev < 64) cache[dev].store(n);
  }
  return n;
}

// Shape / scheduling part of the kernel arguments (shared by the launch path and fa_b200_work_item).
void fill_schedule(fa::FwdArgs& a, long long BH, int Nq, int Nkv, int d) {
  const long long num_q_blocks = (Nq + 2 * fa::kBlockM - 1) / (2 * fa::kBlockM);
  a.Nq = Nq;
  a.Nkv = Nkv;
  a.causal_off = Nkv - Nq;
  a.num_q_blocks = (int)num_q_blocks;
  a.num_items = (int)(BH * num_q_blocks);
  a.num_bh = (int)BH;
  // heads whose K and V (2 * N_kv * d * 2 bytes each) fit in about half of the 126 MB L2 together
  const long long kv_bytes_per_head = 4LL * Nkv * d;
  long long g = (64LL << 20) / (kv_bytes_per_head > 0 ? kv_bytes_per_head : 1);
  g = std::max<long long>(1, std::min<long long>(g, BH));
  if (group_heads_override() > 0) g = group_heads_override();
  a.group_heads = (int)std::max<long long>(1, std::min<long long>(g, BH));
  a.nsplit = 1;
  a.split_begin = 0;
  a.num_ws_items = 0;
  a.tiles_per_split = (Nkv + fa::kBlockN - 1) / fa::kBlockN;
}

// Switch the schedule to `nsplit` key-axis splits of the items from list position `split_begin` on.
void apply_split(fa::FwdArgs& a, int nsplit, long long split_begin) {
  const int n_kv_tiles = (a.Nkv + fa::kBlockN - 1) / fa::kBlockN;
  a.nsplit = nsplit;
  a.tiles_per_split = (n_kv_tiles + nsplit - 1) / nsplit;
  a.split_begin = (int)split_begin;
  a.num_ws_items = a.num_items - (int)split_begin;
  a.num_items = (int)split_begin + a.num_ws_items * nsplit;
}

}  // namespace

extern "C" {

int fa_b200_work_item(int B, int H, int N, int N_kv, int d, int causal, int index, int* bh, int* q0,
                      int* tiles0, int* tiles1) {
  if (B <= 0 || H <= 0 || N <= 0 || N_kv < 0 || d < 8 || d > 128 || (d % 8)) {
    fail(FA_B200_ERR_SHAPE, "work_item: bad shape");
    return 0;
  }
  fa::FwdArgs a{};
  fill_schedule(a, (long long)B * H, N, N_kv ? N_kv : N, d);
  if (index < 0 || index >= a.num_items) {
    fail(FA_B200_ERR_SHAPE, "work_item: index %d outside [0, %d)", index, a.num_items);
    return 0;
  }
  const fa::WorkItem wi = causal ? fa::get_item<true>(a, index) : fa::get_item<false>(a, index);
  if (bh) *bh = wi.bh;
  if (q0) *q0 = wi.q0;
  if (tiles0) *tiles0 = wi.n_t0;
  if (tiles1) *tiles1 = wi.n_t1;
  return a.num_items;
}

size_t fa_b200_workspace_bytes(int B, int H, int N, int N_kv, int d) {
  if (B <= 0 || H <= 0 || N <= 0 || N_kv < 0 || d < 8 || d > 128 || (d % 8)) return 0;
  const int Nkv = N_kv ? N_kv : N;
  const long long items = (long long)B * H * ((N + 2 * fa::kBlockM - 1) / (2 * fa::kBlockM));
  // the causal flag is not an argument: report the larger (non-causal) need, which also covers the causal policy
  long long sb_nc = 0, sb_c = 0;
  const int n_nc = choose_nsplit(items, (Nkv + fa::kBlockN - 1) / fa::kBlockN, sm_count(), false, &sb_nc);
  const int n_c = choose_nsplit(items, (Nkv + fa::kBlockN - 1) / fa::kBlockN, sm_count(), true, &sb_c);
  return std::max(split_workspace_bytes(n_nc, items - sb_nc, d), split_workspace_bytes(n_c, items - sb_c, d));
}

int fa_b200_forward(const fa_b200_params* p) {
  if (!p) return fail(FA_B200_ERR_NULL, "params is NULL");
  if (!p->Q || !p->K || !p->V || !p->O) return fail(FA_B200_ERR_NULL, "Q, K, V and O must be non-NULL");
  if (p->B <= 0 || p->H <= 0 || p->N <= 0 || p->N_kv < 0)
    return fail(FA_B200_ERR_SHAPE, "bad shape B=%d H=%d N=%d N_kv=%d", p->B, p->H, p->N, p->N_kv);
  if (p->d < 8 || p->d > 128 || (p->d % 8))
    return fail(FA_B200_ERR_HEAD_DIM, "unsupported head_dim=%d (supported: multiples of 8 up to 128)", p->d);
  if (p->dtype != FA_B200_FP16 && p->dtype != FA_B200_BF16)
    return fail(FA_B200_ERR_DTYPE, "unsupported dtype=%d (0 = fp16, 1 = bf16)", p->dtype);
  const long long BH = (long long)p->B * p->H;
  const int Nq = p->N, Nkv = p->N_kv ? p->N_kv : p->N;
  const int d = p->d;
  // kernel instantiation: the next of 32 / 64 / 128 (the reference's dispatch set, flash_attn_cutlass.cu:530-542); a
  // head_dim in between runs it with the surplus columns read as zeros by TMA and clipped on store (see encode_tmap) -
  // the reference's FA1 kernel takes any d <= 128 (flashAttention.cu:86) and its Triton path D % 16 == 0 (FA2-triton.py:177)
  const int dk = d <= 32 ? 32 : (d <= 64 ? 64 : 128);
  const long long num_q_blocks = (Nq + 2 * fa::kBlockM - 1) / (2 * fa::kBlockM);
  if (BH * num_q_blocks > 0x7fffffffLL) return fail(FA_B200_ERR_SHAPE, "B*H*ceil(N/256) exceeds the grid limit");
  // strides: 0 => dense [B,H,N,d] default
  struct Str { long long b, h, n; };
  auto resolve = [&](long long sb, long long sh, long long sn, long long rows) {
    Str r;
    r.n = sn ? sn : d;
    r.h = sh ? sh : rows * r.n;
    r.b = sb ? sb : (long long)p->H * r.h;
    return r;
  };
  const Str qs = resolve(p->q_stride_b, p->q_stride_h, p->q_stride_n, Nq);
  const Str ks = resolve(p->kv_stride_b, p->kv_stride_h, p->kv_stride_n, Nkv);
  const Str os = resolve(p->o_stride_b, p->o_stride_h, p->o_stride_n, Nq);
  const long long ssh = p->stat_stride_h ? p->stat_stride_h : (long long)Nq;
  const long long ssb = p->stat_stride_b ? p->stat_stride_b : (long long)p->H * ssh;
  auto mis = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) != 0; };
  if (mis(p->Q) || mis(p->K) || mis(p->V) || mis(p->O))
    return fail(FA_B200_ERR_ALIGNMENT, "Q, K, V, O base pointers must be 16-byte aligned");
  for (const Str* t : {&qs, &ks, &os}) {
    if ((t->b % 8) || (t->h % 8) || (t->n % 8))
      return fail(FA_B200_ERR_ALIGNMENT, "batch / head / row strides must be multiples of 8 elements");
    if (t->b <= 0 || t->h <= 0 || t->n < d)
      return fail(FA_B200_ERR_SHAPE, "strides must be positive and the row stride at least d");
  }
  if (ssh < Nq || ssb <= 0) return fail(FA_B200_ERR_SHAPE, "statistics head stride smaller than N");
  int rc = check_device();
  if (rc) return rc;

  // split-KV: only with a large enough caller-provided workspace
  long long split_begin = 0;
  int nsplit = choose_nsplit(BH * num_q_blocks, (Nkv + fa::kBlockN - 1) / fa::kBlockN, sm_count(), p->causal != 0, &split_begin);
  const long long ws_items = BH * num_q_blocks - split_begin;
  if (nsplit > 1 && (!p->workspace || p->workspace_bytes < split_workspace_bytes(nsplit, ws_items, d) ||
                     (reinterpret_cast<uintptr_t>(p->workspace) & 15u)))
    nsplit = 1;
  const long long ws_rows = ws_items * 2 * fa::kBlockM;     // rows of one split's partial
  char* ws_o = static_cast<char*>(p->workspace);
  float* ws_lse = nsplit > 1 ? reinterpret_cast<float*>(ws_o + (size_t)nsplit * ws_rows * d * 2) : nullptr;
  float* ws_m = nsplit > 1 ? ws_lse + (size_t)nsplit * ws_rows : nullptr;

  CUtensorMap tq, tk, tv, to, tw;
  unsigned perm_q = 0, perm_kv = 0, perm_v = 0, perm_o = 0, perm_w = 0;
  if ((rc = make_tmap(&tq, &perm_q, p->Q, p->dtype, d, Nq, p->H, p->B, qs.n, qs.h, qs.b))) return rc;
  if ((rc = make_tmap(&tk, &perm_kv, p->K, p->dtype, d, Nkv, p->H, p->B, ks.n, ks.h, ks.b))) return rc;
  if ((rc = make_tmap(&tv, &perm_v, p->V, p->dtype, d, Nkv, p->H, p->B, ks.n, ks.h, ks.b))) return rc;
  if ((rc = make_tmap(&to, &perm_o, p->O, p->dtype, d, Nq, p->H, p->B, os.n, os.h, os.b))) return rc;
  if (nsplit > 1) {   // partials go to the workspace, laid out by item: [nsplit][ws_items][256 rows][d]
    if ((rc = make_tmap(&tw, &perm_w, ws_o, p->dtype, d, 2 * fa::kBlockM, ws_items, nsplit, d, 2LL * fa::kBlockM * d, ws_rows * d)))
      return rc;
  } else {
    tw = to;
  }

  const float scale = (p->softmax_scale != 0.f) ? p->softmax_scale : 1.0f / sqrtf((float)d);
  fa::FwdArgs a{};
  a.lse = p->lse;
  a.l = p->l;
  a.m = p->m;
  fill_schedule(a, BH, Nq, Nkv, d);
  a.scale_log2 = scale * 1.4426950408889634f;
  a.need_stats = (p->l || p->m) ? 1 : 0;   // l / m need the exact row max of every tile (no fast softmax path)
  a.stat_stride_b = ssb;
  a.stat_stride_h = ssh;
  a.H = p->H;
  a.perm_q = perm_q;
  a.perm_kv = perm_kv;
  a.perm_o = perm_o;
  a.perm_w = perm_w;
  const unsigned fmt = (p->dtype == FA_B200_BF16) ? 1u : 0u;
  if (dk >= 64) {
    // Q, K: K-major, 128B swizzle: 8-row groups 1024 B apart (SBO); LBO unused for swizzled K-major.
    a.desc_hi_qk = fa::umma_desc_hi_bits(16, 1024, 2);
    // V: MN-major (d contiguous), 128B swizzle: 64-column halves one box (16 KB) apart (LBO),
    // 8-key groups 1024 B apart (SBO).
    a.desc_hi_v = fa::umma_desc_hi_bits(fa::FwdTraits<128>::kBoxBytes, 1024, 2);
  } else {
    // d = 32: 64-byte rows, 64B swizzle (layout type 4): 8-row groups 512 B apart
    a.desc_hi_qk = fa::umma_desc_hi_bits(16, 512, 4);
    a.desc_hi_v = fa::umma_desc_hi_bits(fa::FwdTraits<32>::kBoxBytes, 512, 4);
  }
  a.idesc_qk = fa::umma_idesc(fmt, 0, 0, 128, 128);
  a.idesc_pv = fa::umma_idesc(fmt, 0, 1, 128, (unsigned)dk);
#ifdef FA_B200_DEBUG   // bring-up overrides
  a.desc_hi_qk = env_u64("FA_B200_DESC_HI_QK", a.desc_hi_qk);
  a.desc_hi_v = env_u64("FA_B200_DESC_HI_V", a.desc_hi_v);
  a.idesc_qk = (unsigned)env_u64("FA_B200_IDESC_QK", a.idesc_qk);
  a.idesc_pv = (unsigned)env_u64("FA_B200_IDESC_PV", a.idesc_pv);
#endif

  cudaStream_t stream = reinterpret_cast<cudaStream_t>(p->stream);
  // one CTA per work item; resident CTAs steal the not-yet-launched ones (cluster launch control), so the
  // kernel behaves as a persistent kernel with a dynamic hardware scheduler
  if (nsplit > 1) {
    apply_split(a, nsplit, split_begin);
    a.ws_lse = ws_lse;
    a.ws_m = ws_m;
  }
  const long long grid = a.num_items;
  const bool bf16 = p->dtype == FA_B200_BF16;
  const bool causal = p->causal != 0;
  const bool precise = p->precise != 0;
#define FA_LAUNCH(D_, BF_, C_)                                                             \
  rc = precise ? launch<D_, BF_, C_, true>(tq, tk, tv, to, tw, a, grid, stream)            \
               : launch<D_, BF_, C_, false>(tq, tk, tv, to, tw, a, grid, stream)
  if (dk == 128) {
    if (bf16) { if (causal) FA_LAUNCH(128, true, true); else FA_LAUNCH(128, true, false); }
    else      { if (causal) FA_LAUNCH(128, false, true); else FA_LAUNCH(128, false, false); }
  } else if (dk == 32) {
    if (bf16) { if (causal) FA_LAUNCH(32, true, true); else FA_LAUNCH(32, true, false); }
    else      { if (causal) FA_LAUNCH(32, false, true); else FA_LAUNCH(32, false, false); }
  } else {
    if (bf16) { if (causal) FA_LAUNCH(64, true, true); else FA_LAUNCH(64, true, false); }
    else      { if (causal) FA_LAUNCH(64, false, true); else FA_LAUNCH(64, false, false); }
  }
#undef FA_LAUNCH
  if (rc || nsplit == 1) return rc;
  return fa::launch_item_combine(ws_o, ws_lse, ws_m, p->O, p->lse, p->l, p->m, a, d, os.b, os.h, os.n, ssb, ssh, p->dtype, causal,
                                 stream);
}

int fa_b200_forward_legacy(const void* Q, const void* K, const void* V, void* O, float* l, float* m,
                           int B, int H, int N, int d, int M, void* stream) {
  (void)M;  // FA1 tile-size knob (flashAttention.cu:17-18); tile shape is fixed by the hardware here
  fa_b200_params p;
  memset(&p, 0, sizeof(p));
  p.Q = Q; p.K = K; p.V = V; p.O = O; p.l = l; p.m = m;
  p.B = B; p.H = H; p.N = N; p.d = d;
  p.dtype = FA_B200_FP16;
  p.stream = stream;
  // the reference's FA1 kernel keeps P in fp32 (flashAttention.cu:107-135) and its driver gates on a relative
  // metric (main.cu:346): match that accuracy (see fa_b200_params.precise)
  p.precise = 1;
  return fa_b200_forward(&p);
}

int fa_b200_forward_fp16(const void* Q, const void* K, const void* V, void* O, int batch_size,
                         int num_heads, int seq_len, int head_dim, void* stream) {
  fa_b200_params p;
  memset(&p, 0, sizeof(p));
  p.Q = Q; p.K = K; p.V = V; p.O = O;
  p.B = batch_size; p.H = num_heads; p.N = seq_len; p.d = head_dim;
  p.dtype = FA_B200_FP16;
  p.stream = stream;
  return fa_b200_forward(&p);
}

int fa_b200_backward(const fa_b200_bwd_params* p) {
  if (!p) return fail(FA_B200_ERR_NULL, "params is NULL");
  if (!p->Q || !p->K || !p->V || !p->O || !p->dO || !p->lse || !p->dQ || !p->dK || !p->dV || !p->delta)
    return fail(FA_B200_ERR_NULL, "backward: Q, K, V, O, dO, lse, dQ, dK, dV and delta must be non-NULL");
  if (p->B <= 0 || p->H <= 0 || p->N <= 0 || p->N_kv < 0)
    return fail(FA_B200_ERR_SHAPE, "B, H, N must be positive and N_kv >= 0 (got %d, %d, %d, %d)", p->B, p->H, p->N, p->N_kv);
  const int d = p->d;
  if (d != 32 && d != 64 && d != 128) return fail(FA_B200_ERR_HEAD_DIM, "unsupported head_dim %d (supported: 32, 64, 128)", d);
  if (p->dtype != FA_B200_FP16 && p->dtype != FA_B200_BF16) return fail(FA_B200_ERR_DTYPE, "dtype must be FA_B200_FP16 or FA_B200_BF16");
  const void* ptrs[] = {p->Q, p->K, p->V, p->O, p->dO, p->dQ, p->dK, p->dV};
  for (const void* q : ptrs)
    if (reinterpret_cast<uintptr_t>(q) & 15u) return fail(FA_B200_ERR_ALIGNMENT, "backward: tensors must be 16-byte aligned");
  const long long BH = (long long)p->B * p->H;
  const int Nq = p->N, Nkv = p->N_kv ? p->N_kv : p->N;
  const int q_tiles = (Nq + fa::kBlockM - 1) / fa::kBlockM, kv_tiles = (Nkv + fa::kBlockN - 1) / fa::kBlockN;
  if (BH * std::max(q_tiles, kv_tiles) > 0x7fffffffLL) return fail(FA_B200_ERR_SHAPE, "backward: too many tiles");
  // strides per tensor group: 0 => dense [B,H,rows,d]
  struct Str { long long b, h, n; };
  auto resolve = [&](const int64_t* st, long long rows) {
    Str r;
    r.n = st[2] ? st[2] : d;
    r.h = st[1] ? st[1] : rows * r.n;
    r.b = st[0] ? st[0] : (long long)p->H * r.h;
    return r;
  };
  const Str qs = resolve(p->q_stride, Nq), ks = resolve(p->kv_stride, Nkv), os = resolve(p->o_stride, Nq),
            gs = resolve(p->do_stride, Nq), dqs = resolve(p->dq_stride, Nq), dks = resolve(p->dkv_stride, Nkv);
  for (const Str* t : {&qs, &ks, &os, &gs, &dqs, &dks}) {
    if ((t->b % 8) || (t->h % 8) || (t->n % 8))
      return fail(FA_B200_ERR_ALIGNMENT, "backward: batch / head / row strides must be multiples of 8 elements");
    if (t->b <= 0 || t->h <= 0 || t->n < d)
      return fail(FA_B200_ERR_SHAPE, "backward: strides must be positive and the row stride at least d");
  }
  int rc = check_device();
  if (rc) return rc;

  CUtensorMap tq, tk, tv, tdo, tdq, tdk, tdv;
  fa::BwdArgs a{};
  unsigned perm_v = 0, perm_dv = 0;
  if ((rc = make_tmap(&tq, &a.perm_q, p->Q, p->dtype, d, Nq, p->H, p->B, qs.n, qs.h, qs.b))) return rc;
  if ((rc = make_tmap(&tdo, &a.perm_do, p->dO, p->dtype, d, Nq, p->H, p->B, gs.n, gs.h, gs.b))) return rc;
  if ((rc = make_tmap(&tk, &a.perm_kv, p->K, p->dtype, d, Nkv, p->H, p->B, ks.n, ks.h, ks.b))) return rc;
  if ((rc = make_tmap(&tv, &perm_v, p->V, p->dtype, d, Nkv, p->H, p->B, ks.n, ks.h, ks.b))) return rc;
  if ((rc = make_tmap(&tdq, &a.perm_dq, p->dQ, p->dtype, d, Nq, p->H, p->B, dqs.n, dqs.h, dqs.b))) return rc;
  if ((rc = make_tmap(&tdk, &a.perm_dkv, p->dK, p->dtype, d, Nkv, p->H, p->B, dks.n, dks.h, dks.b))) return rc;
  if ((rc = make_tmap(&tdv, &perm_dv, p->dV, p->dtype, d, Nkv, p->H, p->B, dks.n, dks.h, dks.b))) return rc;
  if (perm_v != a.perm_kv || perm_dv != a.perm_dkv) return fail(FA_B200_ERR_DRIVER, "backward: inconsistent tensor-map axis order");
  const float scale = (p->softmax_scale != 0.f) ? p->softmax_scale : 1.0f / sqrtf((float)d);
  a.lse = p->lse;
  a.delta = p->delta;
  a.Nq = Nq; a.Nkv = Nkv; a.H = p->H;
  a.causal_off = Nkv - Nq;
  a.q_tiles = q_tiles; a.kv_tiles = kv_tiles;
  a.scale = scale;
  a.scale_log2 = scale * 1.4426950408889634f;
  const unsigned fmt = (p->dtype == FA_B200_BF16) ? 1u : 0u;
  if (d >= 64) {
    a.desc_k = fa::umma_desc_hi_bits(16, 1024, 2);
    a.desc_mn = fa::umma_desc_hi_bits(fa::FwdTraits<128>::kBoxBytes, 1024, 2);
  } else {
    a.desc_k = fa::umma_desc_hi_bits(16, 512, 4);
    a.desc_mn = fa::umma_desc_hi_bits(fa::FwdTraits<32>::kBoxBytes, 512, 4);
  }
  a.idesc_ss = fa::umma_idesc(fmt, 0, 0, 128, 128);
  a.idesc_ts = fa::umma_idesc(fmt, 0, 1, 128, (unsigned)d);
  a.idesc_ss_half = fa::umma_idesc(fmt, 0, 0, 128, 64);

  cudaStream_t stream = reinterpret_cast<cudaStream_t>(p->stream);
  const bool bf16 = p->dtype == FA_B200_BF16;
  const bool causal = p->causal != 0;
  // 1. delta = rowsum(dO o O)
  {
    const long long rows = BH * Nq;
    const long long total = rows * (d / 8);
    long long blocks = std::min<long long>((total + 255) / 256, (long long)sm_count() * 8);
    auto is_dense = [&](const Str& t) { return t.n == d && t.h == (long long)Nq * d && t.b == (long long)p->H * Nq * d; };
    const bool dense = is_dense(os) && is_dense(gs);
#define FA_DELTA(BF_, DN_)                                                                                          \
  fa::bwd_delta_kernel<BF_, DN_><<<(unsigned)blocks, 256, 0, stream>>>((const char*)p->O, (const char*)p->dO, p->delta, rows, d, \
                                                                       p->H, Nq, os.b, os.h, os.n, gs.b, gs.h, gs.n)
    if (bf16) { if (dense) FA_DELTA(true, true); else FA_DELTA(true, false); }
    else      { if (dense) FA_DELTA(false, true); else FA_DELTA(false, false); }
#undef FA_DELTA
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(FA_B200_ERR_CUDA, "delta kernel launch: %s", cudaGetErrorString(e));
    g_launches.fetch_add(1);
  }
  // 2. dQ kernel: one CTA per (b,h,query tile);  3. dK/dV kernel: one CTA per (b,h,key tile)
#define FA_BWD(D_, BF_, C_)                                                                               \
  do {                                                                                                    \
    rc = launch_bwd_dq<D_, BF_, C_>(tq, tdo, tk, tv, tdq, a, BH * q_tiles, stream);                       \
    if (!rc) rc = launch_bwd_dkdv<D_, BF_, C_>(tk, tv, tq, tdo, tdk, tdv, a, BH * kv_tiles, stream);      \
  } while (0)
  if (d == 128) {
    if (bf16) { if (causal) FA_BWD(128, true, true); else FA_BWD(128, true, false); }
    else      { if (causal) FA_BWD(128, false, true); else FA_BWD(128, false, false); }
  } else if (d == 32) {
    if (bf16) { if (causal) FA_BWD(32, true, true); else FA_BWD(32, true, false); }
    else      { if (causal) FA_BWD(32, false, true); else FA_BWD(32, false, false); }
  } else {
    if (bf16) { if (causal) FA_BWD(64, true, true); else FA_BWD(64, true, false); }
    else      { if (causal) FA_BWD(64, false, true); else FA_BWD(64, false, false); }
  }
#undef FA_BWD
  return rc;
}

void fa_b200_set_group_heads(int heads) { g_group_heads.store(heads > 0 ? heads : 0); }
uint64_t fa_b200_launch_count(void) { return g_launches.load(); }
void fa_b200_tmap_cache_stats(uint64_t* hits, uint64_t* misses) {
  if (hits) *hits = g_tmap_hits.load();
  if (misses) *misses = g_tmap_misses.load();
}
const char* fa_b200_last_error(void) { return g_err; }
int fa_b200_version(void) { return (FA_B200_VERSION_MAJOR << 16) | FA_B200_VERSION_MINOR; }

const char* fa_b200_status_string(int s) {
  switch (s) {
    case FA_B200_OK: return "ok";
    case FA_B200_ERR_NULL: return "null pointer";
    case FA_B200_ERR_SHAPE: return "bad shape";
    case FA_B200_ERR_HEAD_DIM: return "unsupported head_dim";
    case FA_B200_ERR_DTYPE: return "unsupported dtype";
    case FA_B200_ERR_ALIGNMENT: return "bad alignment";
    case FA_B200_ERR_ARCH: return "not an sm_100 device";
    case FA_B200_ERR_CUDA: return "CUDA error";
    case FA_B200_ERR_DRIVER: return "driver / tensor-map error";
    default: return "unknown status";
  }
}

}  // extern "C"

// Used by fa_merge.cu so that every kernel of the library is counted in fa_b200_launch_count().
namespace fa {
void count_launch() { g_launches.fetch_add(1); }
int api_fail(int code, const char* msg) { return fail(code, "%s", msg); }
int api_check_device() { return check_device(); }
int api_sm_count() { return sm_count(); }
}  // namespace fa
