// fa_host.cu — the host-buffer path behind the C ABI: pinned host Q,K,V -> H2D -> kernel -> D2H -> pinned host O.
//
// Replaces what the reference's drivers do around the launch (code/cuda_fa1/main.cu:392-405 cudaMalloc +
// cudaMemcpy H2D of Q,K,V; :262-275, :417-421 D2H of results) with a three-stream pipeline over groups of
// (b,h) slices, so the PCIe transfers of chunk c+1 / c-1 overlap the tensor-core work of chunk c.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/fa_b200.h"

namespace fa {
int api_fail(int code, const char* msg);
int api_check_device();
}  // namespace fa

struct fa_b200_host_ctx {
  int B, H, N, d, dtype, causal;
  int device;
  std::vector<std::pair<int, int>> ranges;   // [begin, end) of (b*H+h) slices per chunk
  int max_chunk;                             // slices in the largest chunk
  // double-buffered device staging
  void* dq[2]; void* dk[2]; void* dv[2]; void* dout[2]; float* dlse[2];
  cudaStream_t s_h2d, s_comp, s_d2h;
  cudaEvent_t ev_in[2], ev_comp[2], ev_out[2];   // per slot: inputs landed / kernel done / outputs copied
  bool used[2];
  cudaEvent_t ev_first, ev_last;             // timing marks of the most recent call
};

namespace {
#define FA_TRY(call)                                                             \
  do {                                                                           \
    cudaError_t e_ = (call);                                                     \
    if (e_ != cudaSuccess) return fa::api_fail(FA_B200_ERR_CUDA, cudaGetErrorString(e_)); \
  } while (0)
}  // namespace

extern "C" int fa_b200_host_ctx_create(int B, int H, int N, int d, int dtype, int causal, int chunks,
                                       fa_b200_host_ctx** out) {
  if (!out) return fa::api_fail(FA_B200_ERR_NULL, "host_ctx_create: out is NULL");
  *out = nullptr;
  if (B <= 0 || H <= 0 || N <= 0) return fa::api_fail(FA_B200_ERR_SHAPE, "host_ctx_create: bad shape");
  if (d < 8 || d > 128 || (d % 8)) return fa::api_fail(FA_B200_ERR_HEAD_DIM, "host_ctx_create: unsupported head_dim (multiples of 8 up to 128)");
  if (dtype != FA_B200_FP16 && dtype != FA_B200_BF16) return fa::api_fail(FA_B200_ERR_DTYPE, "host_ctx_create: bad dtype");
  int rc = fa::api_check_device();
  if (rc) return rc;
  fa_b200_host_ctx* c = new (std::nothrow) fa_b200_host_ctx();
  if (!c) return fa::api_fail(FA_B200_ERR_CUDA, "host_ctx_create: out of host memory");
  c->B = B; c->H = H; c->N = N; c->d = d; c->dtype = dtype; c->causal = causal;
  cudaGetDevice(&c->device);
  const int BH = B * H;
  chunks = std::max(1, std::min(chunks, BH));
  const int base = BH / chunks, rem = BH % chunks;
  int s = 0;
  for (int i = 0; i < chunks; ++i) {
    const int n = base + (i < rem ? 1 : 0);
    c->ranges.emplace_back(s, s + n);
    s += n;
  }
  c->max_chunk = base + (rem ? 1 : 0);
  const size_t tile = (size_t)c->max_chunk * N * d * 2, stat = (size_t)c->max_chunk * N * sizeof(float);
  for (int i = 0; i < 2; ++i) {
    c->dq[i] = c->dk[i] = c->dv[i] = c->dout[i] = nullptr; c->dlse[i] = nullptr; c->used[i] = false;
  }
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    if (e == cudaSuccess) e = cudaMalloc(&c->dq[i], tile);
    if (e == cudaSuccess) e = cudaMalloc(&c->dk[i], tile);
    if (e == cudaSuccess) e = cudaMalloc(&c->dv[i], tile);
    if (e == cudaSuccess) e = cudaMalloc(&c->dout[i], tile);
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->dlse[i], stat);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_comp[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&c->ev_first);
  if (e == cudaSuccess) e = cudaEventCreate(&c->ev_last);
  if (e != cudaSuccess) {
    fa_b200_host_ctx_destroy(c);
    return fa::api_fail(FA_B200_ERR_CUDA, cudaGetErrorString(e));
  }
  *out = c;
  return FA_B200_OK;
}

extern "C" int fa_b200_forward_host(fa_b200_host_ctx* c, const void* q_host, const void* k_host, const void* v_host,
                                    void* o_host, float* lse_host) {
  if (!c || !q_host || !k_host || !v_host || !o_host) return fa::api_fail(FA_B200_ERR_NULL, "forward_host: NULL argument");
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != c->device) return fa::api_fail(FA_B200_ERR_CUDA, "forward_host: context belongs to another device");
  const size_t row = (size_t)c->N * c->d * 2;        // bytes of one (b,h) slice
  const size_t srow = (size_t)c->N * sizeof(float);
  for (size_t ci = 0; ci < c->ranges.size(); ++ci) {
    const int b = c->ranges[ci].first, n = c->ranges[ci].second - c->ranges[ci].first, s = (int)(ci & 1);
    // H2D: the slot's inputs are free once the kernel that last read them has finished
    if (c->used[s]) FA_TRY(cudaStreamWaitEvent(c->s_h2d, c->ev_comp[s], 0));
    if (ci == 0) FA_TRY(cudaEventRecord(c->ev_first, c->s_h2d));
    FA_TRY(cudaMemcpyAsync(c->dq[s], (const char*)q_host + (size_t)b * row, (size_t)n * row, cudaMemcpyHostToDevice, c->s_h2d));
    FA_TRY(cudaMemcpyAsync(c->dk[s], (const char*)k_host + (size_t)b * row, (size_t)n * row, cudaMemcpyHostToDevice, c->s_h2d));
    FA_TRY(cudaMemcpyAsync(c->dv[s], (const char*)v_host + (size_t)b * row, (size_t)n * row, cudaMemcpyHostToDevice, c->s_h2d));
    FA_TRY(cudaEventRecord(c->ev_in[s], c->s_h2d));
    // compute: needs the inputs, and the slot's outputs must have been copied out
    FA_TRY(cudaStreamWaitEvent(c->s_comp, c->ev_in[s], 0));
    if (c->used[s]) FA_TRY(cudaStreamWaitEvent(c->s_comp, c->ev_out[s], 0));
    fa_b200_params p;
    memset(&p, 0, sizeof(p));
    p.Q = c->dq[s]; p.K = c->dk[s]; p.V = c->dv[s]; p.O = c->dout[s]; p.lse = c->dlse[s];
    p.B = 1; p.H = n; p.N = c->N; p.d = c->d; p.dtype = c->dtype; p.causal = c->causal;
    p.stream = c->s_comp;
    int rc = fa_b200_forward(&p);
    if (rc) return rc;
    FA_TRY(cudaEventRecord(c->ev_comp[s], c->s_comp));
    // D2H
    FA_TRY(cudaStreamWaitEvent(c->s_d2h, c->ev_comp[s], 0));
    FA_TRY(cudaMemcpyAsync((char*)o_host + (size_t)b * row, c->dout[s], (size_t)n * row, cudaMemcpyDeviceToHost, c->s_d2h));
    if (lse_host)
      FA_TRY(cudaMemcpyAsync((char*)lse_host + (size_t)b * srow, c->dlse[s], (size_t)n * srow, cudaMemcpyDeviceToHost, c->s_d2h));
    FA_TRY(cudaEventRecord(c->ev_out[s], c->s_d2h));
    c->used[s] = true;
  }
  FA_TRY(cudaEventRecord(c->ev_last, c->s_d2h));
  return FA_B200_OK;
}

extern "C" int fa_b200_host_ctx_sync(fa_b200_host_ctx* c) {
  if (!c) return fa::api_fail(FA_B200_ERR_NULL, "host_ctx_sync: NULL context");
  FA_TRY(cudaStreamSynchronize(c->s_d2h));
  FA_TRY(cudaStreamSynchronize(c->s_comp));
  FA_TRY(cudaStreamSynchronize(c->s_h2d));
  return FA_B200_OK;
}

extern "C" int fa_b200_host_ctx_elapsed_ms(fa_b200_host_ctx* c, float* ms) {
  if (!c || !ms) return fa::api_fail(FA_B200_ERR_NULL, "host_ctx_elapsed_ms: NULL argument");
  FA_TRY(cudaEventElapsedTime(ms, c->ev_first, c->ev_last));
  return FA_B200_OK;
}

extern "C" void fa_b200_host_ctx_destroy(fa_b200_host_ctx* c) {
  if (!c) return;
  for (int i = 0; i < 2; ++i) {
    if (c->dq[i]) cudaFree(c->dq[i]);
    if (c->dk[i]) cudaFree(c->dk[i]);
    if (c->dv[i]) cudaFree(c->dv[i]);
    if (c->dout[i]) cudaFree(c->dout[i]);
    if (c->dlse[i]) cudaFree(c->dlse[i]);
    if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
    if (c->ev_comp[i]) cudaEventDestroy(c->ev_comp[i]);
    if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
  }
  if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
  if (c->s_comp) cudaStreamDestroy(c->s_comp);
  if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
  if (c->ev_first) cudaEventDestroy(c->ev_first);
  if (c->ev_last) cudaEventDestroy(c->ev_last);
  delete c;
}

// ---------------------------------------------------------------------------------------------- peer memory
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");

extern "C" int fa_b200_peer_alloc(size_t bytes, void** dev_ptr, unsigned char handle[64]) {
  if (!dev_ptr || !handle || bytes == 0) return fa::api_fail(FA_B200_ERR_NULL, "peer_alloc: bad argument");
  int rc = fa::api_check_device();
  if (rc) return rc;
  FA_TRY(cudaMalloc(dev_ptr, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, *dev_ptr);
  if (e != cudaSuccess) {
    cudaFree(*dev_ptr);
    *dev_ptr = nullptr;
    return fa::api_fail(FA_B200_ERR_CUDA, cudaGetErrorString(e));
  }
  memcpy(handle, &h, 64);
  return FA_B200_OK;
}

extern "C" int fa_b200_peer_free(void* dev_ptr) {
  if (dev_ptr) FA_TRY(cudaFree(dev_ptr));
  return FA_B200_OK;
}

extern "C" int fa_b200_peer_open(const unsigned char handle[64], void** dev_ptr) {
  if (!dev_ptr || !handle) return fa::api_fail(FA_B200_ERR_NULL, "peer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  FA_TRY(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return FA_B200_OK;
}

extern "C" int fa_b200_peer_close(void* dev_ptr) {
  if (dev_ptr) FA_TRY(cudaIpcCloseMemHandle(dev_ptr));
  return FA_B200_OK;
}

extern "C" int fa_b200_copy_async(void* dst, const void* src, size_t bytes, void* stream) {
  if (!dst || !src) return fa::api_fail(FA_B200_ERR_NULL, "copy_async: NULL pointer");
  FA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, reinterpret_cast<cudaStream_t>(stream)));
  return FA_B200_OK;
}
