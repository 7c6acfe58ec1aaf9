// fa_ring.cu — ring attention behind the C ABI (SURVEY.md section 8b "ownership" row, section 8e; BASELINE configs[4]).
//
// The reference has no multi-GPU path; this is the north star's long-sequence mode.  One process (or host thread)
// per GPU owns one `fa_b200_ring` handle.  Every rank keeps its K/V block in a buffer the peers of the node have
// mapped (CUDA IPC); a forward call runs P steps, step s attending the local queries to the block of rank
// (r - s) mod P.  What makes it B200-shaped:
//
//   * NVSwitch gives every GPU full bandwidth to every peer, so a block never hops around the ring: the rank that
//     needs it PULLS it straight from its owner with the copy engines (cudaMemcpyAsync on a side stream).  No SM is
//     used - the attention kernel holds every SM with a persistent CTA.
//   * The pulls run through a window of TWO receive slots: the block of step s+1 lands while step s computes and the
//     pull of step s+2 waits for step s's kernel, so the footprint is the ring's O(N/P): one published block + two
//     receive slots, whatever P is.
//   * Cross-rank ordering uses 32-bit sequence flags in the mapped memory instead of a collective: "my block of call
//     n is published" (ready) and "I have pulled your block of call n" (ack) are 4-byte DMA writes into the peer's
//     flag array, and the consumer waits with cuStreamWaitValue32 on its OWN memory - again no SM, no host sync, no
//     NCCL kernel that would have to squeeze in between the persistent attention launches.
//   * The partial of every step goes into a slot of a preallocated stack and ONE HBM-bound pass merges them with
//     their logsumexp at the end (fa_merge.cu).
//
// Causal runs use the zig-zag partition (rank r owns sequence chunks r and 2P-1-r), which turns every step into a
// plain or square-causal call on (strided) row ranges with equal work on every rank - see the step loop.
#include <cuda.h>
#include <cuda_runtime.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/fa_b200.h"

namespace fa {
int api_fail(int code, const char* msg);
int api_check_device();
}  // namespace fa

namespace {

constexpr int kMaxWorld = 64;
constexpr uint64_t kMagic = 0x46413242474e4952ull;   // "FA2BGNIR"
constexpr size_t kFlagBytes = 2 * kMaxWorld * sizeof(uint32_t);   // ready[64] | ack[64]

struct RingExport {          // what fa_b200_ring_export writes: FA_B200_RING_EXPORT_BYTES
  unsigned char ipc[64];     // cudaIpcMemHandle_t of the arena
  uint64_t magic;
  uint64_t arena_ptr;        // the owner's address (used as is when the importer is the same process)
  uint64_t block_bytes;
  int32_t pid, device, rank, world;
  unsigned char pad[FA_B200_RING_EXPORT_BYTES - 64 - 3 * 8 - 4 * 4];
};
static_assert(sizeof(RingExport) == FA_B200_RING_EXPORT_BYTES, "export blob size");

using WaitValue32Fn = CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
using MemsetD32AsyncFn = CUresult (*)(CUdeviceptr, unsigned int, size_t, CUstream);

template <typename Fn>
Fn driver_fn(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  return reinterpret_cast<Fn>(p);
}

#define FA_TRY(call)                                                                      \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) return fa::api_fail(FA_B200_ERR_CUDA, cudaGetErrorString(e_)); \
  } while (0)
#define FA_TRY_DRV(call, what)                                                                        \
  do {                                                                                                \
    CUresult r_ = (call);                                                                             \
    if (r_ != CUDA_SUCCESS) {                                                                         \
      char msg_[96];                                                                                  \
      snprintf(msg_, sizeof(msg_), "%s failed with CUresult %d", what, (int)r_);                      \
      return fa::api_fail(FA_B200_ERR_DRIVER, msg_);                                                  \
    }                                                                                                 \
  } while (0)

}  // namespace

struct fa_b200_ring {
  int world, rank, B, H, n_local, d, dtype, device;
  size_t blk;                       // bytes of one K (or V) block: B*H*n_local*d*2
  char* arena;                      // K | V | flags, exported
  size_t arena_bytes;
  char* peer[kMaxWorld];            // mapped arenas (peer[rank] == arena)
  bool peer_ipc[kMaxWorld];         // opened with cudaIpcOpenMemHandle (must be closed)
  char* recv[2];                    // the two receive slots, K | V each
  char* o_parts;                    // [world][B,H,n_local,d] 16-bit partial outputs
  float* lse_parts;                 // [world][B,H,n_local]
  uint32_t* seq_src;                // [2] device words holding the sequence number of call n (slot n & 1)
  cudaStream_t cs;                  // copy stream
  cudaEvent_t ev_fork, ev_pull[2], ev_done[2];
  bool done_recorded[2];
  uint32_t seq;
  uint32_t ack_waited;               // the call number whose "previous block consumed" waits are already enqueued
  bool connected;
  WaitValue32Fn wait_value;
  MemsetD32AsyncFn memset_d32;
  // optional timeline of the most recent call: per step begin / kv_ready / attn_done, then the combine
  bool profile;
  std::vector<cudaEvent_t> marks;
  int n_marks;
  size_t owned_bytes;

  uint32_t* ready_flags(char* a) const { return reinterpret_cast<uint32_t*>(a + 2 * blk); }
  uint32_t* ack_flags(char* a) const { return ready_flags(a) + kMaxWorld; }
};

extern "C" {

int fa_b200_ring_create(int world, int rank, int B, int H, int n_local, int d, int dtype, fa_b200_ring** out) {
  if (!out) return fa::api_fail(FA_B200_ERR_NULL, "ring_create: out is NULL");
  *out = nullptr;
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world)
    return fa::api_fail(FA_B200_ERR_SHAPE, "ring_create: need 1 <= world <= 64 and 0 <= rank < world");
  if (B <= 0 || H <= 0 || n_local <= 0) return fa::api_fail(FA_B200_ERR_SHAPE, "ring_create: B, H, n_local must be positive");
  if (d != 32 && d != 64 && d != 128) return fa::api_fail(FA_B200_ERR_HEAD_DIM, "ring_create: unsupported head_dim (32, 64, 128)");
  if (dtype != FA_B200_FP16 && dtype != FA_B200_BF16) return fa::api_fail(FA_B200_ERR_DTYPE, "ring_create: bad dtype");
  int rc = fa::api_check_device();
  if (rc) return rc;
  fa_b200_ring* r = new (std::nothrow) fa_b200_ring();
  if (!r) return fa::api_fail(FA_B200_ERR_CUDA, "ring_create: out of host memory");
  r->world = world; r->rank = rank; r->B = B; r->H = H; r->n_local = n_local; r->d = d; r->dtype = dtype;
  r->blk = (size_t)B * H * n_local * d * 2;
  r->arena_bytes = 2 * r->blk + kFlagBytes;
  r->seq = 0; r->ack_waited = 0; r->connected = (world == 1); r->profile = false; r->n_marks = 0; r->owned_bytes = 0;
  cudaGetDevice(&r->device);
  for (int i = 0; i < kMaxWorld; ++i) { r->peer[i] = nullptr; r->peer_ipc[i] = false; }
  r->wait_value = driver_fn<WaitValue32Fn>("cuStreamWaitValue32");
  r->memset_d32 = driver_fn<MemsetD32AsyncFn>("cuMemsetD32Async");
  if (!r->wait_value || !r->memset_d32) {
    delete r;
    return fa::api_fail(FA_B200_ERR_DRIVER, "ring_create: cuStreamWaitValue32 / cuMemsetD32Async not available in this driver");
  }
  cudaError_t e = cudaSuccess;
  const size_t stat = (size_t)B * H * n_local * sizeof(float);
  if (world > 1) {
    e = cudaMalloc((void**)&r->arena, r->arena_bytes);
    if (e == cudaSuccess) e = cudaMemset(r->arena + 2 * r->blk, 0, kFlagBytes);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaMalloc((void**)&r->recv[i], 2 * r->blk);
    if (e == cudaSuccess) e = cudaMalloc((void**)&r->o_parts, (size_t)world * r->blk);
    if (e == cudaSuccess) e = cudaMalloc((void**)&r->lse_parts, (size_t)world * stat);
    if (e == cudaSuccess) e = cudaMalloc((void**)&r->seq_src, 2 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(r->seq_src, 0, 2 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&r->cs, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->ev_fork, cudaEventDisableTiming);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
      e = cudaEventCreateWithFlags(&r->ev_pull[i], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->ev_done[i], cudaEventDisableTiming);
      r->done_recorded[i] = false;
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();   // the zeroed flags are in place before anyone can connect
    r->owned_bytes = r->arena_bytes + 4 * r->blk + (size_t)world * (r->blk + stat) + 8;
    r->peer[rank] = r->arena;
  }
  if (e != cudaSuccess) {
    fa_b200_ring_destroy(r);
    return fa::api_fail(FA_B200_ERR_CUDA, cudaGetErrorString(e));
  }
  *out = r;
  return FA_B200_OK;
}

int fa_b200_ring_export(const fa_b200_ring* r, unsigned char blob[FA_B200_RING_EXPORT_BYTES]) {
  if (!r || !blob) return fa::api_fail(FA_B200_ERR_NULL, "ring_export: NULL argument");
  RingExport x;
  memset(&x, 0, sizeof(x));
  x.magic = kMagic;
  x.arena_ptr = reinterpret_cast<uint64_t>(r->arena);
  x.block_bytes = r->blk;
  x.pid = (int32_t)getpid();
  x.device = r->device; x.rank = r->rank; x.world = r->world;
  if (r->world > 1) {
    cudaIpcMemHandle_t h;
    FA_TRY(cudaIpcGetMemHandle(&h, r->arena));
    memcpy(x.ipc, &h, 64);
  }
  memcpy(blob, &x, sizeof(x));
  return FA_B200_OK;
}

int fa_b200_ring_connect(fa_b200_ring* r, const unsigned char* blobs) {
  if (!r || !blobs) return fa::api_fail(FA_B200_ERR_NULL, "ring_connect: NULL argument");
  if (r->connected) return FA_B200_OK;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != r->device) return fa::api_fail(FA_B200_ERR_CUDA, "ring_connect: handle belongs to another device");
  for (int p = 0; p < r->world; ++p) {
    RingExport x;
    memcpy(&x, blobs + (size_t)p * FA_B200_RING_EXPORT_BYTES, sizeof(x));
    if (x.magic != kMagic || x.rank != p || x.world != r->world || x.block_bytes != r->blk)
      return fa::api_fail(FA_B200_ERR_SHAPE, "ring_connect: export blobs must be in rank order and come from handles of the same shape");
    if (p == r->rank) continue;
    if (x.pid == (int32_t)getpid()) {
      // several ranks in one process (one host thread per GPU, or the single-GPU protocol test): no IPC needed
      if (x.device != r->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(x.device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) return fa::api_fail(FA_B200_ERR_CUDA, cudaGetErrorString(e));
      }
      r->peer[p] = reinterpret_cast<char*>(x.arena_ptr);
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, x.ipc, 64);
      void* ptr = nullptr;
      FA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
      r->peer[p] = static_cast<char*>(ptr);
      r->peer_ipc[p] = true;
    }
  }
  r->connected = true;
  return FA_B200_OK;
}

int fa_b200_ring_kv_buffers(fa_b200_ring* r, void** k_buf, void** v_buf) {
  if (!r || !k_buf || !v_buf) return fa::api_fail(FA_B200_ERR_NULL, "ring_kv_buffers: NULL argument");
  *k_buf = r->world > 1 ? r->arena : nullptr;
  *v_buf = r->world > 1 ? r->arena + r->blk : nullptr;
  return FA_B200_OK;
}

// Enqueues on `stream` the wait for "every peer has pulled the block I published in the previous call".  forward()
// does this itself before it overwrites the publish buffers; a caller that writes K/V straight into those buffers
// (fa_b200_ring_kv_buffers) calls it before the kernel that produces the next K/V.
static int wait_consumed(fa_b200_ring* r, cudaStream_t st) {
  const uint32_t next = r->seq + 1;
  if (r->world == 1 || r->ack_waited == next) return FA_B200_OK;
  for (int p = 0; p < r->world; ++p)
    if (p != r->rank)
      FA_TRY_DRV(r->wait_value(st, reinterpret_cast<CUdeviceptr>(r->ack_flags(r->arena) + p), next - 1, CU_STREAM_WAIT_VALUE_GEQ),
                 "cuStreamWaitValue32(ack)");
  r->ack_waited = next;
  return FA_B200_OK;
}

int fa_b200_ring_wait_consumed(fa_b200_ring* r, void* stream) {
  if (!r) return fa::api_fail(FA_B200_ERR_NULL, "ring_wait_consumed: NULL handle");
  if (!r->connected) return fa::api_fail(FA_B200_ERR_SHAPE, "ring_wait_consumed: call fa_b200_ring_connect first");
  return wait_consumed(r, reinterpret_cast<cudaStream_t>(stream));
}

size_t fa_b200_ring_device_bytes(const fa_b200_ring* r) { return r ? r->owned_bytes : 0; }

int fa_b200_ring_set_profile(fa_b200_ring* r, int on) {
  if (!r) return fa::api_fail(FA_B200_ERR_NULL, "ring_set_profile: NULL handle");
  r->profile = on != 0;
  const size_t want = r->profile ? (size_t)(3 * r->world + 2) : 0;
  while (r->marks.size() < want) {
    cudaEvent_t ev;
    FA_TRY(cudaEventCreate(&ev));
    r->marks.push_back(ev);
  }
  r->n_marks = 0;
  return FA_B200_OK;
}

int fa_b200_ring_timeline(fa_b200_ring* r, float* ms, int cap) {
  if (!r || !ms) { fa::api_fail(FA_B200_ERR_NULL, "ring_timeline: NULL argument"); return -1; }
  int n = 0;
  for (int i = 1; i < r->n_marks && n < cap; ++i, ++n)
    if (cudaEventElapsedTime(&ms[n], r->marks[0], r->marks[i]) != cudaSuccess) { cudaGetLastError(); return -1; }
  return n;
}

int fa_b200_ring_forward(fa_b200_ring* r, const void* Q, const void* K, const void* V, void* O, float* lse,
                         int causal, float softmax_scale, void* stream) {
  if (!r || !Q || !K || !V || !O) return fa::api_fail(FA_B200_ERR_NULL, "ring_forward: NULL argument");
  if (!r->connected) return fa::api_fail(FA_B200_ERR_SHAPE, "ring_forward: call fa_b200_ring_connect first");
  if (causal && (r->n_local & 1) && r->world > 1)
    return fa::api_fail(FA_B200_ERR_SHAPE, "ring_forward: causal (zig-zag) rings need an even local length");
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != r->device) return fa::api_fail(FA_B200_ERR_CUDA, "ring_forward: handle belongs to another device");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int P = r->world, me = r->rank, Nl = r->n_local, d = r->d, half = Nl / 2;
  const long long sh = (long long)Nl * d, sb = (long long)r->H * sh;   // dense [B,H,Nl,d] strides (elements)
  const size_t stat = (size_t)r->B * r->H * Nl;
  int mark = 0;
  auto note = [&]() -> cudaError_t {
    if (!r->profile || mark >= (int)r->marks.size()) return cudaSuccess;
    return cudaEventRecord(r->marks[mark++], st);
  };

  // One local attention call on row ranges of dense [B,H,Nl,d] buffers: queries [q0, q0+nq), keys [0, nkv).
  auto attend = [&](const void* k, const void* v, int q0, int nq, int nkv, bool csl, void* o_dst, float* lse_dst) {
    fa_b200_params p;
    memset(&p, 0, sizeof(p));
    p.Q = static_cast<const char*>(Q) + (size_t)q0 * d * 2;
    p.K = k; p.V = v;
    p.O = static_cast<char*>(o_dst) + (size_t)q0 * d * 2;
    p.lse = lse_dst ? lse_dst + q0 : nullptr;
    p.B = r->B; p.H = r->H; p.N = nq; p.N_kv = nkv; p.d = d; p.dtype = r->dtype;
    p.causal = csl ? 1 : 0;
    p.softmax_scale = softmax_scale;
    p.q_stride_b = sb; p.q_stride_h = sh; p.q_stride_n = d;
    p.kv_stride_b = sb; p.kv_stride_h = sh; p.kv_stride_n = d;
    p.o_stride_b = sb; p.o_stride_h = sh; p.o_stride_n = d;
    p.stat_stride_b = (long long)r->H * Nl; p.stat_stride_h = Nl;
    p.stream = st;
    return fa_b200_forward(&p);
  };

  if (P == 1) {
    FA_TRY(note());
    int rc = attend(K, V, 0, Nl, Nl, causal != 0, O, lse);
    if (rc) return rc;
    FA_TRY(note());
    r->n_marks = mark;
    return FA_B200_OK;
  }

  FA_TRY(note());   // mark 0: call start
  // 1. my published block of the previous call has been pulled by everyone (acks arrive long before this point)
  int rc = wait_consumed(r, st);
  if (rc) return rc;
  const uint32_t seq = ++r->seq;
  uint32_t* seq_word = r->seq_src + (seq & 1u);
  // 2. publish (skipped when the caller already keeps K/V in the ring's own buffers, fa_b200_ring_kv_buffers)
  if (K != r->arena) FA_TRY(cudaMemcpyAsync(r->arena, K, r->blk, cudaMemcpyDeviceToDevice, st));
  if (V != r->arena + r->blk) FA_TRY(cudaMemcpyAsync(r->arena + r->blk, V, r->blk, cudaMemcpyDeviceToDevice, st));
  // 3. tell every peer: 4-byte DMA write of the sequence number into its ready[me]
  FA_TRY_DRV(r->memset_d32(reinterpret_cast<CUdeviceptr>(seq_word), seq, 1, st), "cuMemsetD32Async");
  for (int p = 0; p < P; ++p)
    if (p != me)
      FA_TRY(cudaMemcpyAsync(r->ready_flags(r->peer[p]) + me, seq_word, sizeof(uint32_t), cudaMemcpyDefault, st));
  FA_TRY(cudaEventRecord(r->ev_fork, st));
  FA_TRY(cudaStreamWaitEvent(r->cs, r->ev_fork, 0));
  // every slot row a step does not compute keeps lse = -inf and is skipped by the combine
  FA_TRY_DRV(r->memset_d32(reinterpret_cast<CUdeviceptr>(r->lse_parts), 0xff800000u, (size_t)P * stat, st), "cuMemsetD32Async");

  // Pull of step s into receive slot s & 1 (enqueued on the copy stream).
  auto enqueue_pull = [&](int s) -> int {
    const int src = (me - s + P) % P, slot = s & 1;
    if (r->done_recorded[slot]) FA_TRY(cudaStreamWaitEvent(r->cs, r->ev_done[slot], 0));   // last reader of the slot
    FA_TRY_DRV(r->wait_value(r->cs, reinterpret_cast<CUdeviceptr>(r->ready_flags(r->arena) + src), seq, CU_STREAM_WAIT_VALUE_GEQ),
               "cuStreamWaitValue32(ready)");
    FA_TRY(cudaMemcpyAsync(r->recv[slot], r->peer[src], 2 * r->blk, cudaMemcpyDefault, r->cs));
    FA_TRY(cudaEventRecord(r->ev_pull[slot], r->cs));
    FA_TRY(cudaMemcpyAsync(r->ack_flags(r->peer[src]) + me, seq_word, sizeof(uint32_t), cudaMemcpyDefault, r->cs));
    return FA_B200_OK;
  };

  if ((rc = enqueue_pull(1))) return rc;
  for (int s = 0; s < P; ++s) {
    const int src = (me - s + P) % P, slot = s & 1;
    FA_TRY(note());
    if (s > 0) FA_TRY(cudaStreamWaitEvent(st, r->ev_pull[slot], 0));
    FA_TRY(note());
    const void* k = s == 0 ? K : r->recv[slot];
    const void* v = s == 0 ? V : r->recv[slot] + r->blk;
    char* o_s = r->o_parts + (size_t)s * r->blk;
    float* l_s = r->lse_parts + (size_t)s * stat;
    if (!causal || src == me) {
      rc = attend(k, v, 0, Nl, Nl, causal != 0, o_s, l_s);          // own block: square causal on [chunk r ; chunk 2P-1-r]
    } else if (src < me) {
      rc = attend(k, v, 0, Nl, half, false, o_s, l_s);              // every local query sees only the first half (chunk src)
    } else {
      rc = attend(k, v, half, Nl - half, Nl, false, o_s, l_s);      // only the second half of the queries sees the block
    }
    if (rc) return rc;
    if (s > 0) {
      FA_TRY(cudaEventRecord(r->ev_done[slot], st));
      r->done_recorded[slot] = true;
    }
    FA_TRY(note());
    if (s + 2 < P) {
      // slot (s+2)&1 == slot of step s (for s >= 1) or the slot nobody used yet in this call (s == 0)
      if ((rc = enqueue_pull(s + 2))) return rc;
    }
  }
  rc = fa_b200_combine_partials(r->o_parts, r->lse_parts, P, O, lse, (int64_t)stat, d, r->dtype, st);
  if (rc) return rc;
  FA_TRY(note());
  r->n_marks = mark;
  return FA_B200_OK;
}

int fa_b200_ring_destroy(fa_b200_ring* r) {
  if (!r) return FA_B200_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(r->device);
  cudaDeviceSynchronize();
  for (int p = 0; p < kMaxWorld; ++p)
    if (r->peer_ipc[p] && r->peer[p]) cudaIpcCloseMemHandle(r->peer[p]);
  if (r->arena) cudaFree(r->arena);
  for (int i = 0; i < 2; ++i) {
    if (r->recv[i]) cudaFree(r->recv[i]);
    if (r->ev_pull[i]) cudaEventDestroy(r->ev_pull[i]);
    if (r->ev_done[i]) cudaEventDestroy(r->ev_done[i]);
  }
  if (r->o_parts) cudaFree(r->o_parts);
  if (r->lse_parts) cudaFree(r->lse_parts);
  if (r->seq_src) cudaFree(r->seq_src);
  if (r->cs) cudaStreamDestroy(r->cs);
  if (r->ev_fork) cudaEventDestroy(r->ev_fork);
  for (cudaEvent_t ev : r->marks) cudaEventDestroy(ev);
  cudaGetLastError();
  cudaSetDevice(prev);
  delete r;
  return FA_B200_OK;
}

}  // extern "C"
