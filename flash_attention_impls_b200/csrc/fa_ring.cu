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
//   * Cross-rank ordering uses interprocess CUDA EVENTS instead of a collective: "my block of call n is published"
//     (ready) and "I have finished pulling in call n" (pulled) are events the owner records on its streams and the
//     peers wait for with cudaStreamWaitEvent - no SM, no NCCL kernel that would have to squeeze in between the
//     persistent attention launches, no device-side spinning.  An interprocess event wait only sees records that
//     were ISSUED before it, so the ranks also keep one host-visible sequence counter each (a page of POSIX shared
//     memory): a rank spins on the peer's counter (microseconds: all ranks enter the call together) until the peer
//     has enqueued the record of this call, then enqueues its wait.  The device side stays fully asynchronous.
//     (Stream memory operations - cuStreamWaitValue32 on flags in the mapped memory - would need no host counter, but
//     on this pool's driver 580 the host call that enqueues work BEHIND a pending wait does not return until the value
//     arrives, which deadlocks a rank that enqueues its own wait before its peer's write; the first version of this
//     file hung that way.  Probe: tools/micro/memops_probe.cu, output in profiles/r02_memops_probe.log.)
//   * The partial of every step goes into a slot of a preallocated stack and ONE HBM-bound pass merges them with
//     their logsumexp at the end (fa_merge.cu).
//
// Causal runs use the zig-zag partition (rank r owns sequence chunks r and 2P-1-r), which turns every step into a
// plain or square-causal call on (strided) row ranges with equal work on every rank - see the step loop.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/fa_b200.h"

namespace fa {
int api_fail(int code, const char* msg);
int api_check_device();
int launch_fill_u32(uint32_t* dst, uint32_t value, long long count, cudaStream_t stream);   // fa_merge.cu
}  // namespace fa

namespace {

constexpr int kMaxWorld = 64;
constexpr uint64_t kMagic = 0x46413242474e4952ull;   // "FA2BGNIR"

// One page of shared host memory per rank: the call numbers whose event records this rank has already ISSUED.
struct HostSync {
  std::atomic<uint32_t> ready_seq;     // ev_ready[n & 1] of call n has been recorded (enqueued) on the owner's stream
  std::atomic<uint32_t> pulled_seq;    // ev_pulled[n & 1] of call n has been recorded: every pull of call n is enqueued
};

struct RingExport {          // what fa_b200_ring_export writes: FA_B200_RING_EXPORT_BYTES
  uint64_t magic;
  uint64_t arena_ptr;        // the owner's addresses, used as they are when the importer is the same process
  uint64_t self_ptr;
  uint64_t block_bytes;
  int32_t pid, device, rank, world;
  unsigned char ipc_mem[64];        // cudaIpcMemHandle_t of the arena
  unsigned char ipc_event[4][64];   // cudaIpcEventHandle_t: ready[0], ready[1], pulled[0], pulled[1]
  char shm_name[64];
  unsigned char pad[FA_B200_RING_EXPORT_BYTES - 4 * 8 - 4 * 4 - 64 - 4 * 64 - 64];
};
static_assert(sizeof(RingExport) == FA_B200_RING_EXPORT_BYTES, "export blob size");
static_assert(sizeof(cudaIpcEventHandle_t) == 64 && sizeof(cudaIpcMemHandle_t) == 64, "IPC handle sizes");

#define FA_TRY(call)                                                                      \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) return fa::api_fail(FA_B200_ERR_CUDA, cudaGetErrorString(e_)); \
  } while (0)

// Host wait for a peer's sequence counter (the peer has entered the same call).  Bounded: a rank that never shows up
// is reported instead of hanging the caller.
bool spin_until(const std::atomic<uint32_t>* word, uint32_t seq, double timeout_s) {
  if ((int32_t)(word->load(std::memory_order_acquire) - seq) >= 0) return true;
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned it = 0;; ++it) {
    if ((int32_t)(word->load(std::memory_order_acquire) - seq) >= 0) return true;
    if ((it & 63) == 63) {
      sched_yield();
      if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) return false;
    }
  }
}

}  // namespace

struct fa_b200_ring {
  int world, rank, B, H, n_local, d, dtype, device;
  size_t blk;                       // bytes of one K (or V) block: B*H*n_local*d*2
  char* arena;                      // K | V, exported
  char* peer[kMaxWorld];            // mapped arenas (peer[rank] == arena)
  bool peer_ipc[kMaxWorld];         // opened through CUDA IPC (must be closed)
  cudaEvent_t ev_ready[2], ev_pulled[2];              // mine, interprocess
  cudaEvent_t peer_ready[kMaxWorld][2], peer_pulled[kMaxWorld][2];
  HostSync* hs;                     // mine (shared memory)
  const HostSync* peer_hs[kMaxWorld];
  char shm_name[64];
  char* recv[2];                    // the two receive slots, K | V each
  char* o_parts;                    // [world][B,H,n_local,d] 16-bit partial outputs
  float* lse_parts;                 // [world][B,H,n_local]
  cudaStream_t cs;                  // copy stream
  cudaEvent_t ev_pull[2], ev_done[2];
  bool done_recorded[2];
  uint32_t seq;
  uint32_t consumed_waited;         // the call number whose "previous block consumed" waits are already enqueued
  bool connected;
  double timeout_s;
  // optional timeline of the most recent call: per step begin / kv_ready / attn_done, then the combine
  bool profile;
  std::vector<cudaEvent_t> marks;
  int n_marks;
  size_t owned_bytes;
};

extern "C" {

int fa_b200_ring_create(int world, int rank, int B, int H, int n_local, int d, int dtype, fa_b200_ring** out) {
  if (!out) return fa::api_fail(FA_B200_ERR_NULL, "ring_create: out is NULL");
  *out = nullptr;
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world)
    return fa::api_fail(FA_B200_ERR_SHAPE, "ring_create: need 1 <= world <= 64 and 0 <= rank < world");
  if (B <= 0 || H <= 0 || n_local <= 0) return fa::api_fail(FA_B200_ERR_SHAPE, "ring_create: B, H, n_local must be positive");
  if (d < 8 || d > 128 || (d % 8)) return fa::api_fail(FA_B200_ERR_HEAD_DIM, "ring_create: unsupported head_dim (multiples of 8 up to 128)");
  if (dtype != FA_B200_FP16 && dtype != FA_B200_BF16) return fa::api_fail(FA_B200_ERR_DTYPE, "ring_create: bad dtype");
  int rc = fa::api_check_device();
  if (rc) return rc;
  fa_b200_ring* r = new (std::nothrow) fa_b200_ring();
  if (!r) return fa::api_fail(FA_B200_ERR_CUDA, "ring_create: out of host memory");
  r->world = world; r->rank = rank; r->B = B; r->H = H; r->n_local = n_local; r->d = d; r->dtype = dtype;
  r->blk = (size_t)B * H * n_local * d * 2;
  r->seq = 0; r->consumed_waited = 0; r->connected = (world == 1); r->profile = false; r->n_marks = 0; r->owned_bytes = 0;
  r->timeout_s = 60.0;
  if (const char* t = getenv("FA_B200_RING_TIMEOUT_S")) r->timeout_s = atof(t) > 0 ? atof(t) : r->timeout_s;
  cudaGetDevice(&r->device);
  if (world == 1) {
    *out = r;
    return FA_B200_OK;
  }
  // host-visible sequence counters in POSIX shared memory
  static std::atomic<unsigned> counter{0};
  snprintf(r->shm_name, sizeof(r->shm_name), "/fa_b200_ring_%d_%d_%u", (int)getpid(), rank, counter.fetch_add(1));
  const char* err = nullptr;
  int fd = shm_open(r->shm_name, O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0 || ftruncate(fd, 4096) != 0) err = "ring_create: shm_open / ftruncate failed";
  void* m = err ? MAP_FAILED : mmap(nullptr, 4096, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  if (fd >= 0) close(fd);
  if (!err && m == MAP_FAILED) err = "ring_create: mmap of the shared counters failed";
  if (err) {
    r->shm_name[0] = 0;
    fa_b200_ring_destroy(r);
    return fa::api_fail(FA_B200_ERR_CUDA, err);
  }
  r->hs = new (m) HostSync();
  r->hs->ready_seq.store(0);
  r->hs->pulled_seq.store(0);
  r->peer_hs[rank] = r->hs;

  cudaError_t e = cudaSuccess;
  const size_t stat = (size_t)B * H * n_local * sizeof(float);
  e = cudaMalloc((void**)&r->arena, 2 * r->blk);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaMalloc((void**)&r->recv[i], 2 * r->blk);
  if (e == cudaSuccess) e = cudaMalloc((void**)&r->o_parts, (size_t)world * r->blk);
  if (e == cudaSuccess) e = cudaMalloc((void**)&r->lse_parts, (size_t)world * stat);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&r->cs, cudaStreamNonBlocking);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&r->ev_pull[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->ev_done[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->ev_ready[i], cudaEventDisableTiming | cudaEventInterprocess);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->ev_pulled[i], cudaEventDisableTiming | cudaEventInterprocess);
    r->done_recorded[i] = false;
  }
  r->owned_bytes = 6 * r->blk + (size_t)world * (r->blk + stat);
  r->peer[rank] = r->arena;
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
  x.self_ptr = reinterpret_cast<uint64_t>(r);
  x.block_bytes = r->blk;
  x.pid = (int32_t)getpid();
  x.device = r->device; x.rank = r->rank; x.world = r->world;
  if (r->world > 1) {
    cudaIpcMemHandle_t h;
    FA_TRY(cudaIpcGetMemHandle(&h, r->arena));
    memcpy(x.ipc_mem, &h, 64);
    const cudaEvent_t evs[4] = {r->ev_ready[0], r->ev_ready[1], r->ev_pulled[0], r->ev_pulled[1]};
    for (int i = 0; i < 4; ++i) {
      cudaIpcEventHandle_t eh;
      FA_TRY(cudaIpcGetEventHandle(&eh, evs[i]));
      memcpy(x.ipc_event[i], &eh, 64);
    }
    memcpy(x.shm_name, r->shm_name, sizeof(x.shm_name));
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
      const fa_b200_ring* o = reinterpret_cast<const fa_b200_ring*>(x.self_ptr);
      if (x.device != r->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(x.device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) return fa::api_fail(FA_B200_ERR_CUDA, cudaGetErrorString(e));
      }
      r->peer[p] = o->arena;
      for (int i = 0; i < 2; ++i) { r->peer_ready[p][i] = o->ev_ready[i]; r->peer_pulled[p][i] = o->ev_pulled[i]; }
      r->peer_hs[p] = o->hs;
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, x.ipc_mem, 64);
      void* ptr = nullptr;
      FA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
      r->peer[p] = static_cast<char*>(ptr);
      r->peer_ipc[p] = true;
      for (int i = 0; i < 4; ++i) {
        cudaIpcEventHandle_t eh;
        memcpy(&eh, x.ipc_event[i], 64);
        FA_TRY(cudaIpcOpenEventHandle(i < 2 ? &r->peer_ready[p][i] : &r->peer_pulled[p][i - 2], eh));
      }
      x.shm_name[sizeof(x.shm_name) - 1] = 0;
      int fd = shm_open(x.shm_name, O_RDONLY, 0);
      void* m = fd < 0 ? MAP_FAILED : mmap(nullptr, 4096, PROT_READ, MAP_SHARED, fd, 0);
      if (fd >= 0) close(fd);
      if (m == MAP_FAILED) return fa::api_fail(FA_B200_ERR_CUDA, "ring_connect: cannot map a peer's shared counters (ranks on different hosts?)");
      r->peer_hs[p] = static_cast<const HostSync*>(m);
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

// Orders `st` behind "every peer has pulled the block I published in the previous call".  forward() does this itself
// before it overwrites the publish buffers; a caller that writes K/V straight into those buffers
// (fa_b200_ring_kv_buffers) calls it before the kernel that produces the next K/V.
static int wait_consumed(fa_b200_ring* r, cudaStream_t st) {
  const uint32_t next = r->seq + 1, prev = r->seq;
  if (r->world == 1 || r->consumed_waited == next) return FA_B200_OK;
  if (prev > 0)
    for (int p = 0; p < r->world; ++p) {
      if (p == r->rank) continue;
      if (!spin_until(&r->peer_hs[p]->pulled_seq, prev, r->timeout_s))
        return fa::api_fail(FA_B200_ERR_CUDA, "ring: a peer did not finish enqueuing the previous ring forward (timeout)");
      FA_TRY(cudaStreamWaitEvent(st, r->peer_pulled[p][prev & 1u], 0));
    }
  r->consumed_waited = next;
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
  // 1. the block I published in the previous call has been pulled by everyone (true long before this point)
  int rc = wait_consumed(r, st);
  if (rc) return rc;
  const uint32_t seq = ++r->seq;
  const int par = int(seq & 1u);
  // 2. publish (skipped when the caller already keeps K/V in the ring's own buffers, fa_b200_ring_kv_buffers),
  //    record "ready", and tell the peers' hosts that the record of this call has been issued
  if (K != r->arena) FA_TRY(cudaMemcpyAsync(r->arena, K, r->blk, cudaMemcpyDeviceToDevice, st));
  if (V != r->arena + r->blk) FA_TRY(cudaMemcpyAsync(r->arena + r->blk, V, r->blk, cudaMemcpyDeviceToDevice, st));
  FA_TRY(cudaEventRecord(r->ev_ready[par], st));
  r->hs->ready_seq.store(seq, std::memory_order_release);
  // every slot row a step does not compute keeps lse = -inf and is skipped by the combine
  if ((rc = fa::launch_fill_u32(reinterpret_cast<uint32_t*>(r->lse_parts), 0xff800000u, (long long)P * stat, st))) return rc;

  // Pull of step s into receive slot s & 1 (enqueued on the copy stream): wait for the slot's last reader and for the
  // owner's "ready" of THIS call, copy K|V in one piece, record.
  auto enqueue_pull = [&](int s) -> int {
    const int src = (me - s + P) % P, slot = s & 1;
    if (!spin_until(&r->peer_hs[src]->ready_seq, seq, r->timeout_s))
      return fa::api_fail(FA_B200_ERR_CUDA, "ring_forward: a peer did not enter this ring forward (timeout); every rank must call it");
    if (r->done_recorded[slot]) FA_TRY(cudaStreamWaitEvent(r->cs, r->ev_done[slot], 0));
    FA_TRY(cudaStreamWaitEvent(r->cs, r->peer_ready[src][par], 0));
    FA_TRY(cudaMemcpyAsync(r->recv[slot], r->peer[src], 2 * r->blk, cudaMemcpyDefault, r->cs));
    FA_TRY(cudaEventRecord(r->ev_pull[slot], r->cs));
    if (s == P - 1) {   // the last pull of the call: "I am done pulling"
      FA_TRY(cudaEventRecord(r->ev_pulled[par], r->cs));
      r->hs->pulled_seq.store(seq, std::memory_order_release);
    }
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
    // window of two slots: the pull of step s+2 reuses the slot of step s (s >= 1) or the still unused one (s == 0)
    if (s + 2 < P && (rc = enqueue_pull(s + 2))) return rc;
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
  for (int p = 0; p < kMaxWorld; ++p) {
    if (!r->peer_ipc[p]) continue;
    if (r->peer[p]) cudaIpcCloseMemHandle(r->peer[p]);
    for (int i = 0; i < 2; ++i) {
      if (r->peer_ready[p][i]) cudaEventDestroy(r->peer_ready[p][i]);
      if (r->peer_pulled[p][i]) cudaEventDestroy(r->peer_pulled[p][i]);
    }
    if (r->peer_hs[p]) munmap(const_cast<HostSync*>(r->peer_hs[p]), 4096);
  }
  if (r->arena) cudaFree(r->arena);
  for (int i = 0; i < 2; ++i) {
    if (r->recv[i]) cudaFree(r->recv[i]);
    if (r->ev_pull[i]) cudaEventDestroy(r->ev_pull[i]);
    if (r->ev_done[i]) cudaEventDestroy(r->ev_done[i]);
    if (r->ev_ready[i]) cudaEventDestroy(r->ev_ready[i]);
    if (r->ev_pulled[i]) cudaEventDestroy(r->ev_pulled[i]);
  }
  if (r->o_parts) cudaFree(r->o_parts);
  if (r->lse_parts) cudaFree(r->lse_parts);
  if (r->cs) cudaStreamDestroy(r->cs);
  for (cudaEvent_t ev : r->marks) cudaEventDestroy(ev);
  if (r->hs) munmap(r->hs, 4096);
  if (r->shm_name[0]) shm_unlink(r->shm_name);
  cudaGetLastError();
  cudaSetDevice(prev);
  delete r;
  return FA_B200_OK;
}

}  // extern "C"
