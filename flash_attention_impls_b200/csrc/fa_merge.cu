// fa_merge.cu — logsumexp merge of attention partials for ring attention (SURVEY.md section 8e;
// the reference has no multi-GPU path, this is new work named by BASELINE.json's north_star).
//
//   lse' = log(e^lse_a + e^lse_b),  O' = O_a e^(lse_a - lse') + O_b e^(lse_b - lse')
//
// HBM-bound element-wise kernels: one warp per row group, 16-byte vector loads/stores, grid-stride.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/fa_b200.h"
#include "fa_fwd_sm100.cuh"

namespace fa {
void count_launch();
int api_fail(int code, const char* msg);
int api_check_device();
int api_sm_count();

template <bool kBF16>
__device__ __forceinline__ uint32_t pack2f(float a, float b) {
  if constexpr (kBF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
}

// One thread handles 8 consecutive elements of a row (16 B of the 16-bit partial, 32 B of fp32).
template <bool kBF16>
__global__ void __launch_bounds__(256)
merge_partial_kernel(float* __restrict__ O_acc, float* __restrict__ lse_acc,
                     const uint4* __restrict__ O_part, const float* __restrict__ lse_part,
                     long long rows, int d) {
  const int vec_per_row = d / 8;
  const long long total = rows * vec_per_row;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / vec_per_row;
    const int v = (int)(idx - row * vec_per_row);
    const float la = lse_acc[row], lb = lse_part[row];
    if (lb == -INFINITY) continue;            // nothing to add for this row
    const float mx = fmaxf(la, lb);
    const float ea = (la == -INFINITY) ? 0.f : __expf(la - mx);
    const float eb = __expf(lb - mx);
    const float inv = 1.f / (ea + eb);
    const float wa = ea * inv, wb = eb * inv;
    const uint4 pv = O_part[idx];
    float4* acc = reinterpret_cast<float4*>(O_acc + row * d + v * 8);
    float4 a0 = acc[0], a1 = acc[1];
    const float2 p0 = unpack2<kBF16>(pv.x), p1 = unpack2<kBF16>(pv.y), p2 = unpack2<kBF16>(pv.z),
                 p3 = unpack2<kBF16>(pv.w);
    a0.x = a0.x * wa + p0.x * wb; a0.y = a0.y * wa + p0.y * wb;
    a0.z = a0.z * wa + p1.x * wb; a0.w = a0.w * wa + p1.y * wb;
    a1.x = a1.x * wa + p2.x * wb; a1.y = a1.y * wa + p2.y * wb;
    a1.z = a1.z * wa + p3.x * wb; a1.w = a1.w * wa + p3.y * wb;
    acc[0] = a0; acc[1] = a1;
  }
}

// lse_acc update runs after every thread of the merge has read the old value: separate tiny kernel.
__global__ void __launch_bounds__(256)
merge_lse_kernel(float* __restrict__ lse_acc, const float* __restrict__ lse_part, long long rows) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows;
       r += (long long)gridDim.x * blockDim.x) {
    const float la = lse_acc[r], lb = lse_part[r];
    if (lb == -INFINITY) continue;
    const float mx = fmaxf(la, lb);
    const float ea = (la == -INFINITY) ? 0.f : __expf(la - mx);
    lse_acc[r] = mx + __logf(ea + __expf(lb - mx));
  }
}

template <bool kBF16>
__global__ void __launch_bounds__(256)
cast_output_kernel(uint4* __restrict__ O, const float4* __restrict__ O_acc, long long nvec) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < nvec;
       idx += (long long)gridDim.x * blockDim.x) {
    const float4 a0 = O_acc[2 * idx], a1 = O_acc[2 * idx + 1];
    uint4 o;
    o.x = pack2f<kBF16>(a0.x, a0.y); o.y = pack2f<kBF16>(a0.z, a0.w);
    o.z = pack2f<kBF16>(a1.x, a1.y); o.w = pack2f<kBF16>(a1.z, a1.w);
    O[idx] = o;
  }
}

// Split-KV combine: partial (O_s, lse_s, m_s), s < nsplit, of the same rows over disjoint key ranges ->
//   lse = log sum_s e^lse_s,  O = sum_s e^(lse_s - lse) O_s,  m = max_s m_s,  l = e^(lse - m).
// Partials are dense [nsplit][rows][d] / [nsplit][rows]; the outputs use the caller's strides.
// One thread per 8 consecutive elements of a row.
template <bool kBF16>
__global__ void __launch_bounds__(256)
split_combine_kernel(const uint4* __restrict__ O_part, const float* __restrict__ lse_part,
                     const float* __restrict__ m_part, char* __restrict__ O, float* __restrict__ lse,
                     float* __restrict__ l, float* __restrict__ m, int nsplit, long long rows, int d, int H, int N,
                     long long o_sb, long long o_sh, long long o_sn, long long st_sb, long long st_sh) {
  const int vec_per_row = d / 8;
  const long long total = rows * vec_per_row;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / vec_per_row;
    const int v = (int)(idx - row * vec_per_row);
    float mx = -INFINITY;
    for (int s = 0; s < nsplit; ++s) mx = fmaxf(mx, lse_part[s * rows + row]);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float sum = 0.f;
    if (mx != -INFINITY) {
      for (int s = 0; s < nsplit; ++s) {
        const float ls = lse_part[s * rows + row];
        if (ls == -INFINITY) continue;
        const float w = __expf(ls - mx);
        sum += w;
        const uint4 pv = O_part[(s * rows + row) * vec_per_row + v];
        const float2 p0 = unpack2<kBF16>(pv.x), p1 = unpack2<kBF16>(pv.y), p2 = unpack2<kBF16>(pv.z),
                     p3 = unpack2<kBF16>(pv.w);
        acc[0] += w * p0.x; acc[1] += w * p0.y; acc[2] += w * p1.x; acc[3] += w * p1.y;
        acc[4] += w * p2.x; acc[5] += w * p2.y; acc[6] += w * p3.x; acc[7] += w * p3.y;
      }
    }
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    const long long b = row / ((long long)H * N);
    const long long rem = row - b * (long long)H * N;
    const long long h = rem / N, n = rem - h * N;
    uint4 o;
    o.x = pack2f<kBF16>(acc[0] * inv, acc[1] * inv); o.y = pack2f<kBF16>(acc[2] * inv, acc[3] * inv);
    o.z = pack2f<kBF16>(acc[4] * inv, acc[5] * inv); o.w = pack2f<kBF16>(acc[6] * inv, acc[7] * inv);
    *reinterpret_cast<uint4*>(O + (b * o_sb + h * o_sh + n * o_sn + v * 8) * 2) = o;
    if (v == 0) {
      const long long so = b * st_sb + h * st_sh + n;
      const float lse_v = sum > 0.f ? mx + __logf(sum) : -INFINITY;
      if (lse) lse[so] = lse_v;
      if (m || l) {
        float mm = -INFINITY;
        for (int s = 0; s < nsplit; ++s) mm = fmaxf(mm, m_part[s * rows + row]);
        if (m) m[so] = mm;
        if (l) l[so] = sum > 0.f ? __expf(lse_v - mm) : 0.f;
      }
    }
  }
}

// Combine of the forward kernel's split items (FwdArgs::nsplit > 1): partials are laid out by item,
// O_part [nsplit][num_ws_items][256][d], lse_part / m_part [nsplit][num_ws_items][256]; item t of the workspace is list
// position split_begin + t, i.e. rows q0 .. q0+255 of slice bh (item_coords).  Same arithmetic as split_combine_kernel;
// the outputs use the caller's strides.  One thread per 8 consecutive elements of a row.
template <bool kBF16, bool kCausal>
__global__ void __launch_bounds__(256)
item_combine_kernel(const uint4* __restrict__ O_part, const float* __restrict__ lse_part, const float* __restrict__ m_part,
                    char* __restrict__ O, float* __restrict__ lse, float* __restrict__ l, float* __restrict__ m,
                    const FwdArgs a, int d, long long o_sb, long long o_sh, long long o_sn, long long st_sb, long long st_sh) {
  const int vec_per_row = d / 8;
  const int rows_per_item = 2 * kBlockM;
  const long long part_rows = (long long)a.num_ws_items * rows_per_item;
  const long long total = part_rows * vec_per_row;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long prow = idx / vec_per_row;              // row inside one split's partial
    const int v = (int)(idx - prow * vec_per_row);
    const int t = (int)(prow / rows_per_item), r = (int)(prow - (long long)t * rows_per_item);
    int bh, qb;
    item_coords<kCausal>(a, a.split_begin + t, bh, qb);
    const int n = qb * rows_per_item + r;
    if (n >= a.Nq) continue;
    float mx = -INFINITY;
    for (int s = 0; s < a.nsplit; ++s) mx = fmaxf(mx, lse_part[s * part_rows + prow]);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float sum = 0.f;
    if (mx != -INFINITY) {
      for (int s = 0; s < a.nsplit; ++s) {
        const float ls = lse_part[s * part_rows + prow];
        if (ls == -INFINITY) continue;
        const float w = __expf(ls - mx);
        sum += w;
        const uint4 pv = O_part[(s * part_rows + prow) * vec_per_row + v];
        const float2 p0 = unpack2<kBF16>(pv.x), p1 = unpack2<kBF16>(pv.y), p2 = unpack2<kBF16>(pv.z),
                     p3 = unpack2<kBF16>(pv.w);
        acc[0] += w * p0.x; acc[1] += w * p0.y; acc[2] += w * p1.x; acc[3] += w * p1.y;
        acc[4] += w * p2.x; acc[5] += w * p2.y; acc[6] += w * p3.x; acc[7] += w * p3.y;
      }
    }
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    const long long b = bh / a.H, h = bh - b * a.H;
    uint4 o;
    o.x = pack2f<kBF16>(acc[0] * inv, acc[1] * inv); o.y = pack2f<kBF16>(acc[2] * inv, acc[3] * inv);
    o.z = pack2f<kBF16>(acc[4] * inv, acc[5] * inv); o.w = pack2f<kBF16>(acc[6] * inv, acc[7] * inv);
    *reinterpret_cast<uint4*>(O + (b * o_sb + h * o_sh + n * o_sn + v * 8) * 2) = o;
    if (v == 0) {
      const long long so = b * st_sb + h * st_sh + n;
      const float lse_v = sum > 0.f ? mx + __logf(sum) : -INFINITY;
      if (lse) lse[so] = lse_v;
      if (m || l) {
        float mm = -INFINITY;
        for (int s = 0; s < a.nsplit; ++s) mm = fmaxf(mm, m_part[s * part_rows + prow]);
        if (m) m[so] = mm;
        if (l) l[so] = sum > 0.f ? __expf(lse_v - mm) : 0.f;
      }
    }
  }
}

// 32-bit pattern fill (the ring driver's "lse = -inf" reset of its partial stack)
__global__ void __launch_bounds__(256) fill_u32_kernel(uint32_t* __restrict__ dst, uint32_t value, long long count) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    dst[i] = value;
}

// launched by fa_api.cu after the split-KV forward kernel
int launch_split_combine(const void* O_part, const float* lse_part, const float* m_part, void* O, float* lse,
                         float* l, float* m, int nsplit, long long rows, int d, int H, int N, long long o_sb,
                         long long o_sh, long long o_sn, long long st_sb, long long st_sh, int dtype,
                         cudaStream_t stream);

inline unsigned grid_for(long long work_items) {
  long long blocks = (work_items + 255) / 256;
  const long long cap = (long long)api_sm_count() * 8;   // 8 resident 256-thread CTAs per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}
int launch_split_combine(const void* O_part, const float* lse_part, const float* m_part, void* O, float* lse,
                         float* l, float* m, int nsplit, long long rows, int d, int H, int N, long long o_sb,
                         long long o_sh, long long o_sn, long long st_sb, long long st_sh, int dtype,
                         cudaStream_t stream) {
  const long long total = rows * (d / 8);
  if (dtype == FA_B200_BF16)
    split_combine_kernel<true><<<grid_for(total), 256, 0, stream>>>((const uint4*)O_part, lse_part, m_part, (char*)O, lse, l,
                                                                    m, nsplit, rows, d, H, N, o_sb, o_sh, o_sn, st_sb, st_sh);
  else
    split_combine_kernel<false><<<grid_for(total), 256, 0, stream>>>((const uint4*)O_part, lse_part, m_part, (char*)O, lse, l,
                                                                     m, nsplit, rows, d, H, N, o_sb, o_sh, o_sn, st_sb, st_sh);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return api_fail(FA_B200_ERR_CUDA, cudaGetErrorString(e));
  count_launch();
  return FA_B200_OK;
}
int launch_item_combine(const void* O_part, const float* lse_part, const float* m_part, void* O, float* lse, float* l, float* m,
                        const FwdArgs& a, int d, long long o_sb, long long o_sh, long long o_sn, long long st_sb, long long st_sh,
                        int dtype, bool causal, cudaStream_t stream) {
  const long long total = (long long)a.num_ws_items * 2 * kBlockM * (d / 8);
  const unsigned grid = grid_for(total);
#define FA_IC(BF_, C_)                                                                                                   \
  item_combine_kernel<BF_, C_><<<grid, 256, 0, stream>>>((const uint4*)O_part, lse_part, m_part, (char*)O, lse, l, m, a, d, o_sb, \
                                                         o_sh, o_sn, st_sb, st_sh)
  if (dtype == FA_B200_BF16) { if (causal) FA_IC(true, true); else FA_IC(true, false); }
  else                       { if (causal) FA_IC(false, true); else FA_IC(false, false); }
#undef FA_IC
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return api_fail(FA_B200_ERR_CUDA, cudaGetErrorString(e));
  count_launch();
  return FA_B200_OK;
}

int launch_fill_u32(uint32_t* dst, uint32_t value, long long count, cudaStream_t stream) {
  fill_u32_kernel<<<grid_for(count), 256, 0, stream>>>(dst, value, count);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return api_fail(FA_B200_ERR_CUDA, cudaGetErrorString(e));
  count_launch();
  return FA_B200_OK;
}
}  // namespace fa

extern "C" int fa_b200_merge_partial(float* O_acc, float* lse_acc, const void* O_part, const float* lse_part,
                                     int64_t rows, int d, int dtype, void* stream) {
  if (!O_acc || !lse_acc || !O_part || !lse_part) return fa::api_fail(FA_B200_ERR_NULL, "merge: NULL pointer");
  if (rows <= 0 || d <= 0 || d % 8) return fa::api_fail(FA_B200_ERR_SHAPE, "merge: rows > 0 and d % 8 == 0 required");
  if (dtype != FA_B200_FP16 && dtype != FA_B200_BF16) return fa::api_fail(FA_B200_ERR_DTYPE, "merge: bad dtype");
  if ((reinterpret_cast<uintptr_t>(O_acc) | reinterpret_cast<uintptr_t>(O_part)) & 15)
    return fa::api_fail(FA_B200_ERR_ALIGNMENT, "merge: O_acc and O_part must be 16-byte aligned");
  int rc = fa::api_check_device();
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = rows * (d / 8);
  if (dtype == FA_B200_BF16)
    fa::merge_partial_kernel<true><<<fa::grid_for(total), 256, 0, s>>>(O_acc, lse_acc, (const uint4*)O_part, lse_part, rows, d);
  else
    fa::merge_partial_kernel<false><<<fa::grid_for(total), 256, 0, s>>>(O_acc, lse_acc, (const uint4*)O_part, lse_part, rows, d);
  fa::merge_lse_kernel<<<fa::grid_for(rows), 256, 0, s>>>(lse_acc, lse_part, rows);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fa::api_fail(FA_B200_ERR_CUDA, cudaGetErrorString(e));
  fa::count_launch(); fa::count_launch();
  return FA_B200_OK;
}

extern "C" int fa_b200_combine_partials(const void* O_parts, const float* lse_parts, int nparts, void* O, float* lse,
                                        int64_t rows, int d, int dtype, void* stream) {
  if (!O_parts || !lse_parts || !O) return fa::api_fail(FA_B200_ERR_NULL, "combine: NULL pointer");
  if (nparts <= 0 || rows <= 0 || rows > 0x7fffffffLL || d <= 0 || d % 8)
    return fa::api_fail(FA_B200_ERR_SHAPE, "combine: nparts > 0, 0 < rows < 2^31 and d % 8 == 0 required");
  if (dtype != FA_B200_FP16 && dtype != FA_B200_BF16) return fa::api_fail(FA_B200_ERR_DTYPE, "combine: bad dtype");
  if ((reinterpret_cast<uintptr_t>(O_parts) | reinterpret_cast<uintptr_t>(O)) & 15)
    return fa::api_fail(FA_B200_ERR_ALIGNMENT, "combine: O_parts and O must be 16-byte aligned");
  int rc = fa::api_check_device();
  if (rc) return rc;
  // one "batch", one "head" of `rows` rows: dense [rows, d] output, dense [rows] statistics
  return fa::launch_split_combine(O_parts, lse_parts, nullptr, O, lse, nullptr, nullptr, nparts, rows, d, 1, (int)rows,
                                  0, 0, d, 0, 0, dtype, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int fa_b200_cast_output(void* O, const float* O_acc, int64_t rows, int d, int dtype, void* stream) {
  if (!O || !O_acc) return fa::api_fail(FA_B200_ERR_NULL, "cast: NULL pointer");
  if (rows <= 0 || d <= 0 || d % 8) return fa::api_fail(FA_B200_ERR_SHAPE, "cast: rows > 0 and d % 8 == 0 required");
  if (dtype != FA_B200_FP16 && dtype != FA_B200_BF16) return fa::api_fail(FA_B200_ERR_DTYPE, "cast: bad dtype");
  if ((reinterpret_cast<uintptr_t>(O_acc) | reinterpret_cast<uintptr_t>(O)) & 15)
    return fa::api_fail(FA_B200_ERR_ALIGNMENT, "cast: O and O_acc must be 16-byte aligned");
  int rc = fa::api_check_device();
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long nvec = rows * (d / 8);
  if (dtype == FA_B200_BF16)
    fa::cast_output_kernel<true><<<fa::grid_for(nvec), 256, 0, s>>>((uint4*)O, (const float4*)O_acc, nvec);
  else
    fa::cast_output_kernel<false><<<fa::grid_for(nvec), 256, 0, s>>>((uint4*)O, (const float4*)O_acc, nvec);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fa::api_fail(FA_B200_ERR_CUDA, cudaGetErrorString(e));
  fa::count_launch();
  return FA_B200_OK;
}
