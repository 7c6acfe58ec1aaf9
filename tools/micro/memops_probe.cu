// memops_probe.cu — does cuStreamWaitValue32 behave as the ring driver (csrc/fa_ring.cu) needs on this driver?
// One process, one GPU.  Each case: stream A waits on a 32-bit flag in device memory, stream B raises it; the host
// polls both streams with a deadline instead of synchronising, so a case that never completes is reported, not hung on.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/micro/memops_probe tools/micro/memops_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)
#define CKD(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { printf("driver error %d at %s:%d\n", (int)r_, __FILE__, __LINE__); exit(2); } } while (0)

__global__ void spin_kernel(long long clocks) {
  const long long t0 = clock64();
  while (clock64() - t0 < clocks) {}
}

static bool wait_done(cudaStream_t a, cudaStream_t b, double seconds) {
  const auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    const cudaError_t qa = cudaStreamQuery(a), qb = cudaStreamQuery(b);
    if (qa == cudaSuccess && qb == cudaSuccess) return true;
    if ((qa != cudaErrorNotReady && qa != cudaSuccess) || (qb != cudaErrorNotReady && qb != cudaSuccess)) {
      printf("  stream error %s / %s\n", cudaGetErrorString(qa), cudaGetErrorString(qb));
      return false;
    }
    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > seconds) return false;
    std::this_thread::sleep_for(std::chrono::milliseconds(1));
  }
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);   // a hung case must not swallow the lines before it
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  int v = 0;
  CUdevice dev;
  CKD(cuDeviceGet(&dev, 0));
  cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_CAN_USE_STREAM_WAIT_VALUE_NOR, dev);
  printf("attr CAN_USE_STREAM_WAIT_VALUE_NOR_V2 = %d\n", v);
  cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_CAN_USE_64_BIT_STREAM_MEM_OPS, dev);
  printf("attr CAN_USE_64_BIT_STREAM_MEM_OPS_V2 = %d\n", v);
  cudaStream_t A, B;
  CK(cudaStreamCreateWithFlags(&A, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&B, cudaStreamNonBlocking));
  uint32_t *flag, *src, *big;
  CK(cudaMalloc(&flag, 256));
  CK(cudaMalloc(&src, 256));
  CK(cudaMalloc(&big, 64 << 20));
  uint32_t h = 0;

  for (int mode = 0; mode < 5; ++mode) {
    const char* names[] = {"raise with cuStreamWriteValue32", "raise with 4-byte cudaMemcpyAsync D2D from a word set by cuMemsetD32Async",
                           "raise with cuMemsetD32Async", "as case 1, waiter enqueued first and a kernel behind the wait",
                           "as case 1 with CU_STREAM_WAIT_VALUE_FLUSH"};
    CK(cudaMemset(flag, 0, 256));
    CK(cudaMemset(src, 0, 256));
    CK(cudaDeviceSynchronize());
    const uint32_t seq = 7;
    unsigned int wflags = CU_STREAM_WAIT_VALUE_GEQ | (mode == 4 ? CU_STREAM_WAIT_VALUE_FLUSH : 0);
    // waiter first (as the ring does: rank 0's whole call is enqueued before rank 1's)
    printf("case %d: enqueuing the wait ...\n", mode);
    const auto c0 = std::chrono::steady_clock::now();
    CKD(cuStreamWaitValue32((CUstream)A, (CUdeviceptr)flag, seq, wflags));
    printf("  cuStreamWaitValue32 returned after %.3f ms\n",
           std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - c0).count());
    if (mode == 3) spin_kernel<<<1, 32, 0, A>>>(1000);
    CK(cudaMemcpyAsync(big, big + (8 << 20), 32 << 20, cudaMemcpyDeviceToDevice, A));
    // some work on B before the raise
    spin_kernel<<<1, 32, 0, B>>>(2000000);
    if (mode == 0) {
      CKD(cuStreamWriteValue32((CUstream)B, (CUdeviceptr)flag, seq, 0));
    } else if (mode == 2) {
      CKD(cuMemsetD32Async((CUdeviceptr)flag, seq, 1, (CUstream)B));
    } else {
      CKD(cuMemsetD32Async((CUdeviceptr)src, seq, 1, (CUstream)B));
      CK(cudaMemcpyAsync(flag, src, 4, cudaMemcpyDefault, B));
    }
    const bool ok = wait_done(A, B, 5.0);
    CK(cudaMemcpy(&h, flag, 4, cudaMemcpyDeviceToHost));
    printf("case %d (%s): %s, flag = %u\n", mode, names[mode], ok ? "completed" : "DID NOT COMPLETE in 5 s", h);
    if (!ok) {   // release the waiter from the host so the next case starts clean
      h = seq;
      cudaMemcpyAsync(flag, &h, 4, cudaMemcpyHostToDevice, B);
      if (!wait_done(A, B, 5.0)) { printf("  could not release the waiter; giving up\n"); return 1; }
    }
  }
  printf("probe done\n");
  return 0;
}
