/*
 * fa_b200.h — C ABI of the B200 (sm_100a) FlashAttention forward library (libfa_b200.so).
 *
 * This is the drop-in boundary for the CUDA launch path of santiweide/flash-attention-impls.
 * Every entry point takes plain device pointers and sizes; nothing here depends on torch,
 * CUTLASS or C++.  Tensors are the reference's layout: contiguous row-major [B, H, N, d]
 * (reference: code/cuda_fa1/flashAttention.cu:30, offset ((b*H+h)*N)*d) and per-row
 * statistics [B, H, N] fp32 (flashAttention.cu:31).
 *
 * Reference interfaces replaced (paths relative to the reference repository):
 *   - code/cuda_fa1/flashAttention.h:8-11   __global__ flash_attention_forward(Q,K,V,O,l,m,B,H,N,d,M)
 *     and its raw <<<grid,block,shmem>>> call sites code/cuda_fa1/main.cu:297-303, :444
 *       -> fa_b200_forward_legacy()
 *   - code/cutlass_cuda_fa1/run/flash_attn_cutlass.cu:519-544  flash_attention_cutlass_dispatch
 *   - code/cutlass_cuda_fa1/run/flash_attn_unified.cu:545-617  flash_attention_forward_dispatch,
 *     flash_attention_small_tile_dispatch, attention_reference_dispatch
 *       -> fa_b200_forward_fp16()  (same argument list; the C++-linkage shims with the reference's
 *          exact names live in flash_attention_impls_b200/csrc/fa_ref_shims.cu)
 *   - code/triton_fa2/FA2-triton.py:173-205,240-244  flash_attention(q,k,v,causal) -> (o, m, l)
 *       -> fa_b200_forward() with causal=1 and the l/m/lse outputs
 *
 * All calls are enqueue-only on `stream`, allocate nothing, never synchronise, never exit(),
 * and are safe to call from several host threads (one per GPU).  There is NO CPU fallback: on a
 * device that is not sm_100 the calls return FA_B200_ERR_ARCH.
 */
#ifndef FA_B200_H_
#define FA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FA_B200_VERSION_MAJOR 0
#define FA_B200_VERSION_MINOR 6

/* status codes (0 = ok).  The reference returns void and prints to stderr
 * (flash_attn_cutlass.cu:540-542, :510-514); the C ABI returns codes instead. */
enum fa_b200_status {
  FA_B200_OK = 0,
  FA_B200_ERR_NULL = 1,        /* a required pointer is NULL */
  FA_B200_ERR_SHAPE = 2,       /* B,H,N <= 0, N_kv < 0, B*H too large */
  FA_B200_ERR_HEAD_DIM = 3,    /* forward: d not a multiple of 8 in [8,128] (the reference: any d <= 128, flashAttention.cu:86;
                                  D % 16 == 0, FA2-triton.py:177; {32,64,128} in its dispatchers); backward: d not in {32,64,128} */
  FA_B200_ERR_DTYPE = 4,       /* dtype not FA_B200_FP16 / FA_B200_BF16 */
  FA_B200_ERR_ALIGNMENT = 5,   /* base pointers must be 16-byte aligned, strides multiples of 8 elements */
  FA_B200_ERR_ARCH = 6,        /* current device is not compute capability 10.x */
  FA_B200_ERR_CUDA = 7,        /* a CUDA runtime/driver call failed; see fa_b200_last_error() */
  FA_B200_ERR_DRIVER = 8       /* cuTensorMapEncodeTiled unavailable or failed */
};

enum fa_b200_dtype {
  FA_B200_FP16 = 0,            /* __half, the reference's dtype (flashAttention.h:8-11) */
  FA_B200_BF16 = 1             /* __nv_bfloat16 (new; BASELINE configs c3-c5) */
};

/* Full parameter block.  Zero-initialise, then fill.  Fields the reference has no notion of
 * (N_kv, strides, causal, lse) default to the reference behaviour when left 0 / NULL. */
typedef struct fa_b200_params {
  const void* Q;     /* [B,H,N,d]    dtype */
  const void* K;     /* [B,H,N_kv,d] dtype */
  const void* V;     /* [B,H,N_kv,d] dtype */
  void* O;           /* [B,H,N,d]    dtype */
  float* lse;        /* optional [B,H,N]: natural-log logsumexp of the scaled scores, = m + ln(l) */
  float* l;          /* optional [B,H,N]: sum_j exp(s_ij - m_i)    (flashAttention.cu:115-120,137) */
  float* m;          /* optional [B,H,N]: max_j s_ij, s = q.k*scale (flashAttention.cu:138)
                        (asking for l or m keeps the kernel on its exact-row-max path for every tile; O and lse alone
                        let it skip the row-max pass where a tile provably needs none - same results to rounding) */
  int B, H, N, d;
  int N_kv;          /* 0 => N (self-attention, the only case the reference has) */
  int dtype;         /* enum fa_b200_dtype */
  int causal;        /* 0 / 1.  Mask rule col > row + (N_kv - N) -> -inf; with N_kv == N this is
                        FA2-triton.py:70-73 (square, top-left aligned) */
  float softmax_scale; /* 0 => 1/sqrt(d) (flashAttention.cu:96, flash_attn_cutlass.cu:470) */
  /* Element strides of the batch, head and row (sequence) axes; the d axis is always contiguous.  A stride left
   * 0 takes its dense [B,H,N,d] default (row: d, head: N*d, batch: H*N*d; statistics: head N, batch H*N), so a
   * zeroed block means the reference's layout (flashAttention.cu:30).  Non-default strides describe views:
   * row sub-ranges of a longer sequence (the ring driver), or a [B,N,H,d] tensor (row stride H*d, head stride
   * d) without a copy - the reference's Triton path takes arbitrary strides the same way
   * (FA2-triton.py:190-195).  K and V share their strides; lse, l and m share theirs (rows contiguous).
   * Every stride must be a multiple of 8 elements (TMA: 16 bytes). */
  int64_t q_stride_b, q_stride_h, q_stride_n;
  int64_t kv_stride_b, kv_stride_h, kv_stride_n;
  int64_t o_stride_b, o_stride_h, o_stride_n;
  int64_t stat_stride_b, stat_stride_h;
  void* stream;      /* cudaStream_t; NULL => legacy default stream, as in the reference */
  /* Optional device scratch for split-KV scheduling: launches with far fewer (b,h,256-row) work items than SMs
   * (e.g. the reference's (1,1,N,64) sweep, report/pmph-a6.tex:282-286) are cut along the key axis into partial
   * results that a second kernel combines with their logsumexp; so is the last, at most half-filled wave of a
   * non-causal launch of long equal items (c3 sharded over 8 GPUs: 512 items on 148 SMs).  Size it with
   * fa_b200_workspace_bytes(); NULL or too small => the plain single-pass schedule.  The callee still allocates
   * nothing. */
  void* workspace;
  size_t workspace_bytes;
  /* 1: feed P to the P.V tensor-core product as two 16-bit operands (P = P_hi + P_lo, the rounding residual),
   * which gives P fp32-like accuracy - what the reference's CUDA-core FA1 kernel has (flashAttention.cu:107-135) -
   * at about 1.4x the time.  0 (default): one 16-bit P, the usual tensor-core FlashAttention accuracy (well inside
   * the 2e-3 max-abs gate; differs from an fp32-P result by ~2^-12 relative per product term).
   * fa_b200_forward_legacy sets it, so that the reference's own driver keeps passing its 2 % relative gate
   * (main.cu:346) on its near-zero test outputs. */
  int precise;
} fa_b200_params;

/* Primary entry point. */
int fa_b200_forward(const fa_b200_params* p);

/* Scratch bytes fa_b200_forward would like for this shape (0 when it would not split). */
size_t fa_b200_workspace_bytes(int B, int H, int N, int N_kv, int d);

/* Same argument list as the reference kernel flash_attention_forward (flashAttention.h:8-11)
 * plus the stream; call this where the reference does `flash_attention_forward<<<grid,block,
 * shmem>>>(Q,K,V,O,l,m,B,H,N,d,M)`.  fp16, non-causal.  M (the FA1 "SRAM size" knob that picks
 * Bc/Br, flashAttention.cu:17-18) is accepted and ignored.  l and m get the reference's meaning. */
int fa_b200_forward_legacy(const void* Q, const void* K, const void* V, void* O,
                           float* l, float* m, int B, int H, int N, int d, int M,
                           void* stream);

/* Same argument list as flash_attention_cutlass_dispatch (flash_attn_cutlass.cu:519-529) and the
 * three dispatchers of flash_attn_unified.cu:545-617.  fp16, non-causal, O only. */
int fa_b200_forward_fp16(const void* Q, const void* K, const void* V, void* O,
                         int batch_size, int num_heads, int seq_len, int head_dim,
                         void* stream);

/* Merge two attention partials over disjoint key sets (ring attention):
 *   lse' = log(exp(lse_acc) + exp(lse_part));  O_acc' = O_acc*exp(lse_acc-lse') + O_part*exp(lse_part-lse')
 * O_acc is fp32 [rows, d] (row stride d), O_part is dtype [rows, d]; lse_* are fp32 [rows].
 * Rows whose lse_part is -inf are left unchanged.  In place on (O_acc, lse_acc). */
int fa_b200_merge_partial(float* O_acc, float* lse_acc, const void* O_part, const float* lse_part,
                          int64_t rows, int d, int dtype, void* stream);

/* Merge `nparts` attention partials over disjoint key sets in one pass (what the ring driver runs once at the end
 * of a forward; same arithmetic as the split-KV combine):
 *   lse = log(sum_s exp(lse_s));   O = sum_s exp(lse_s - lse) * O_s
 * O_parts is dtype [nparts, rows, d], lse_parts fp32 [nparts, rows]; O is dtype [rows, d], lse fp32 [rows] (optional).
 * A partial whose lse_s is -inf for a row is skipped for that row (its O_s row is never read, so it may be
 * uninitialised); rows with no finite partial give O = 0, lse = -inf.  HBM-bound: reads (finite partials) * 2 bytes and
 * writes 2 bytes per element, against 10 bytes per element and partial for the incremental fa_b200_merge_partial. */
int fa_b200_combine_partials(const void* O_parts, const float* lse_parts, int nparts, void* O, float* lse,
                             int64_t rows, int d, int dtype, void* stream);

/* Final cast of the fp32 ring accumulator to dtype: O[rows,d] = (dtype) O_acc[rows,d]. */
int fa_b200_cast_output(void* O, const float* O_acc, int64_t rows, int d, int dtype, void* stream);

/* ---- backward (reference: code/triton_fa2/FA2-triton.py:98-170, :207-237) ---------------------
 * dQ, dK, dV of O = softmax(Q K^T scale [+ causal mask]) V for an upstream gradient dO, from the forward's O and
 * logsumexp (for the reference's saved statistics: lse = m + ln l, FA2-triton.py:203).  Q, O, dO, dQ are [B,H,N,d] and
 * K, V, dK, dV [B,H,N_kv,d] views with free batch / head / row strides, as the reference's backward takes them
 * (`q.stride(i)` ... `dV.stride(i)`, FA2-triton.py:219-227); the causal mask is the forward's (col > row + N_kv - N).
 * Three launches: delta = rowsum(dO o O) (HBM-bound), a dQ kernel and a dK/dV kernel (tcgen05; see
 * flash_attention_impls_b200/csrc/fa_bwd_sm100.cuh).  Every output element has one writer: no atomics (the reference
 * accumulates dK/dV with fp16 atomic adds, :164-167), deterministic, outputs are overwritten, not accumulated into.
 * `delta` is caller-provided fp32 scratch of B*H*N elements; lse and delta are dense [B,H,N].  d in {32, 64, 128}. */
typedef struct fa_b200_bwd_params {
  const void* Q;     /* [B,H,N,d] dtype */
  const void* K;     /* [B,H,N_kv,d] */
  const void* V;
  const void* O;     /* forward output */
  const void* dO;    /* upstream gradient */
  const float* lse;  /* [B,H,N] forward logsumexp (-inf for a row that saw no key) */
  void* dQ;          /* [B,H,N,d] dtype, written */
  void* dK;          /* [B,H,N_kv,d] dtype, written */
  void* dV;
  float* delta;      /* [B,H,N] fp32 scratch */
  int B, H, N, d;
  int dtype;         /* enum fa_b200_dtype */
  int causal;
  float softmax_scale; /* 0 => 1/sqrt(d) */
  void* stream;
  /* ---- everything below may stay zero: dense tensors, N_kv == N (the only case version 0.4 had) */
  int N_kv;          /* 0 => N */
  /* element strides {batch, head, row} per tensor group, 0 => the dense default; d is contiguous; multiples of 8 */
  int64_t q_stride[3];     /* Q */
  int64_t kv_stride[3];    /* K and V */
  int64_t o_stride[3];     /* O */
  int64_t do_stride[3];    /* dO */
  int64_t dq_stride[3];    /* dQ */
  int64_t dkv_stride[3];   /* dK and dV */
} fa_b200_bwd_params;
int fa_b200_backward(const fa_b200_bwd_params* p);

/* ---- host-buffer path (what bench.py's `e2e` figure times) -------------------------------------------------
 * The reference's drivers keep Q, K, V on the host and copy them over before launching (main.cu:403-405,
 * test_flash_attn.cu:86-104).  This is that path as one call: pinned host Q,K,V -> device -> kernel -> pinned host
 * O (+ lse).  The (b,h) slices are independent, so the work is cut into `chunks` groups of slices and pipelined
 * on three streams (H2D / compute / D2H) through double-buffered device staging owned by the context; staging is
 * allocated once in fa_b200_host_ctx_create, never per call.  fa_b200_forward_host() enqueues one whole forward
 * and returns; fa_b200_host_ctx_sync() waits for everything enqueued on the context. */
typedef struct fa_b200_host_ctx fa_b200_host_ctx;
int fa_b200_host_ctx_create(int B, int H, int N, int d, int dtype, int causal, int chunks, fa_b200_host_ctx** out);
int fa_b200_forward_host(fa_b200_host_ctx* ctx, const void* q_host, const void* k_host, const void* v_host,
                         void* o_host, float* lse_host /* may be NULL */);
int fa_b200_host_ctx_sync(fa_b200_host_ctx* ctx);
/* Device time (CUDA events) from the first H2D copy to the end of the last D2H copy of the most recent
 * fa_b200_forward_host call; call after fa_b200_host_ctx_sync. */
int fa_b200_host_ctx_elapsed_ms(fa_b200_host_ctx* ctx, float* ms);
void fa_b200_host_ctx_destroy(fa_b200_host_ctx* ctx);

/* ---- peer memory over NVLink (ring attention transport) ----------------------------------------------------
 * One process per GPU: a rank publishes its K/V block in a buffer it exports with CUDA IPC; the other ranks of
 * the node map it and PULL the block with the copy engines (cudaMemcpyAsync over NVLink/NVSwitch), which needs no
 * SM - the attention kernel occupies every SM with a persistent CTA, so SM-based send/recv kernels would only run
 * in the gaps between launches.  fa_b200_peer_alloc: cudaMalloc + export (handle is 64 opaque bytes);
 * fa_b200_peer_open: map a peer's buffer (peer access is enabled lazily); fa_b200_copy_async: DMA copy on
 * `stream` between any two device pointers. */
int fa_b200_peer_alloc(size_t bytes, void** dev_ptr, unsigned char handle[64]);
int fa_b200_peer_free(void* dev_ptr);
int fa_b200_peer_open(const unsigned char handle[64], void** dev_ptr);
int fa_b200_peer_close(void* dev_ptr);
int fa_b200_copy_async(void* dst, const void* src, size_t bytes, void* stream);

/* ---- ring attention behind the C ABI (SURVEY.md section 8b ownership row, 8e; BASELINE configs[4]) -----------------
 * One handle per GPU (one process or host thread per GPU), single node.  The sequence is split over `world` ranks;
 * rank r holds Q, K, V rows [B,H,n_local,d] (dense).  A forward runs `world` steps; step s attends the local queries
 * to the K/V block of rank (r - s) mod world, which is PULLED from its owner's exported buffer by the copy engines
 * (no SM, no NCCL kernel) through a window of two receive slots, ordered by interprocess CUDA events (no collective,
 * no device-side spinning; the ranks' hosts exchange one sequence counter each through a page of POSIX shared memory,
 * so a forward may hold the calling host thread for the few microseconds until every peer has ENTERED the same
 * call - one host thread per rank).  The partials are merged with their logsumexp.  Causal rings use the zig-zag partition: the local rows of rank r are sequence chunks r and
 * 2*world-1-r (of 2*world equal chunks), in that order; non-causal rings accept any equal partition.
 *
 * Life cycle:  create (allocates everything the handle will ever use: one published K|V block, two receive slots,
 * the partial stack - fa_b200_ring_device_bytes reports it; forward never allocates)  ->  export (a 512-byte blob)
 * -> the CALLER exchanges the blobs between ranks by any means (MPI, torch.distributed, a file)  ->  connect (maps
 * the peers; blobs in rank order)  ->  forward, any number of times, collectively (every rank the same number of
 * calls)  ->  destroy (collective: no rank may still be pulling; synchronise the ranks first).
 * Ranks living in the same process (one host thread per GPU) connect without IPC. */
#define FA_B200_RING_EXPORT_BYTES 512
typedef struct fa_b200_ring fa_b200_ring;
int fa_b200_ring_create(int world, int rank, int B, int H, int n_local, int d, int dtype, fa_b200_ring** out);
int fa_b200_ring_export(const fa_b200_ring* ring, unsigned char blob[FA_B200_RING_EXPORT_BYTES]);
int fa_b200_ring_connect(fa_b200_ring* ring, const unsigned char* blobs /* world * FA_B200_RING_EXPORT_BYTES */);
/* O [B,H,n_local,d] dtype and lse [B,H,n_local] fp32 (optional) for the local rows.  Enqueue-only on `stream` (plus the
 * handle's own copy stream); softmax_scale 0 => 1/sqrt(d). */
int fa_b200_ring_forward(fa_b200_ring* ring, const void* Q, const void* K, const void* V, void* O, float* lse,
                         int causal, float softmax_scale, void* stream);
/* Zero-copy publish: the handle's own K and V buffers (NULL for world == 1).  A caller that produces K/V straight
 * into them and passes these pointers to forward() saves the publish copy; it must enqueue
 * fa_b200_ring_wait_consumed on the producing stream first (the peers may still be pulling the previous block). */
int fa_b200_ring_kv_buffers(fa_b200_ring* ring, void** k_buf, void** v_buf);
int fa_b200_ring_wait_consumed(fa_b200_ring* ring, void* stream);
/* Device memory the handle owns, in bytes (published block + two receive slots + partial stack + flags). */
size_t fa_b200_ring_device_bytes(const fa_b200_ring* ring);
/* Optional CUDA-event timeline of the most recent forward (diagnostics; costs a few event records per step):
 * after set_profile(ring, 1) and a synchronised forward, timeline() writes milliseconds since the call started for
 * [step0 begin, step0 K/V ready, step0 kernel done, step1 begin, ..., combine done] and returns the count (-1 on error). */
int fa_b200_ring_set_profile(fa_b200_ring* ring, int on);
int fa_b200_ring_timeline(fa_b200_ring* ring, float* ms, int cap);
int fa_b200_ring_destroy(fa_b200_ring* ring);

/* Introspection of the tile scheduler (host-only, no GPU needed): decodes work item `index` of the launch that
 * fa_b200_forward would make for this shape - which (b*H+h) slice, first query row, and how many 128-key K/V
 * tiles each of its two 128-row Q tiles visits (0 = tile skipped).  Returns the number of work items
 * (= grid size); on error returns 0 and sets fa_b200_last_error().  The item list is head-major; causal launches order items longest-first inside groups
 * of heads whose K/V fit the L2 together. */
int fa_b200_work_item(int B, int H, int N, int N_kv, int d, int causal, int index, int* bh, int* q0,
                      int* tiles0, int* tiles1);

/* Tuning knob of the causal item order: heads per longest-first group (0 = the default, as many heads as keep
 * their K/V in about half of the L2 together).  Process-wide; its initial value is read ONCE from the environment
 * variable FA_B200_GROUP_HEADS.  Results never depend on it, only the order in which work items are visited. */
void fa_b200_set_group_heads(int heads);

/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
uint64_t fa_b200_launch_count(void);

/* Hits / misses of the per-thread TMA tensor-map cache (keyed on pointer, shape, strides, dtype). */
void fa_b200_tmap_cache_stats(uint64_t* hits, uint64_t* misses);

/* Message for the last non-OK status returned on the calling thread ("" if none). */
const char* fa_b200_last_error(void);

/* Human-readable name of a status code. */
const char* fa_b200_status_string(int status);

/* (major << 16) | minor */
int fa_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FA_B200_H_ */
