// Masked mean pool of a PLM's last hidden state into fp32 item embeddings (reference
// data_process/amazon_text_emb.py:91-96).  HBM bound: every hidden element is read once with 128-bit loads, eight rows in
// flight per thread; padded positions are not read.  Small batches (the reference pools ONE sequence at a time) are split
// along the sequence so that the grid still covers the chip; the partial sums are combined in split order (deterministic).
#include <algorithm>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace lcrec {

constexpr int kPoolThreads = 128;
constexpr int kPoolRows = 8;            // rows in flight per thread

template <typename T> struct PoolVec;
template <> struct PoolVec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
  static __device__ __forceinline__ float one(const float* p) { return __ldg(p); }
};
template <> struct PoolVec<__half> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __half* p, float (&v)[8]) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  static __device__ __forceinline__ float one(const __half* p) { return __half2float(*p); }
};
template <> struct PoolVec<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  static __device__ __forceinline__ float one(const __nv_bfloat16* p) { return __bfloat162float(*p); }
};

__device__ __forceinline__ void pool_store(float* out, float sum, int64_t count, int accumulate, float divide_by) {
  float r = __fdiv_rn(sum, (float)count);                  // `/ mask.sum(-1, keepdim=True)`: int64 -> fp32, IEEE divide
  if (accumulate) r = __fadd_rn(*out, r);
  if (divide_by > 0.f) r = __fdiv_rn(r, divide_by);
  *out = r;
}

// grid (column chunks, splits, sequences).  VEC = elements per 16-byte load, or 1 for hidden sizes that are not a multiple.
template <typename T, int VEC>
__global__ void __launch_bounds__(kPoolThreads)
pool_partial_kernel(const T* __restrict__ hidden, const int64_t* __restrict__ mask, int64_t seq_len, int hidden_dim,
                    int rows_per_split, int n_splits, float* __restrict__ out, int64_t out_stride, int accumulate,
                    float divide_by, float* __restrict__ partial, int64_t* __restrict__ partial_count) {
  const int64_t b = blockIdx.z;
  const int s = blockIdx.y;
  const int col = (blockIdx.x * kPoolThreads + threadIdx.x) * VEC;
  const bool live = col < hidden_dim;
  const int64_t t0 = (int64_t)s * rows_per_split;
  const int64_t t1 = t0 + rows_per_split < seq_len ? t0 + rows_per_split : seq_len;
  const T* base = hidden + b * seq_len * hidden_dim + (live ? col : 0);
  const int64_t* mrow = mask + b * seq_len;
  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
  int64_t count = 0;
  for (int64_t t = t0; t < t1; t += kPoolRows) {
    float v[kPoolRows][VEC];
    float m[kPoolRows];
#pragma unroll
    for (int r = 0; r < kPoolRows; ++r) {
      const int64_t mv = t + r < t1 ? __ldg(mrow + t + r) : 0;
      count += mv;
      m[r] = (float)mv;
      if (mv != 0 && live) {
        if constexpr (VEC == 1) v[r][0] = PoolVec<T>::one(base + (t + r) * hidden_dim);
        else PoolVec<T>::load(base + (t + r) * hidden_dim, v[r]);
      } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[r][i] = 0.f;
      }
    }
#pragma unroll
    for (int r = 0; r < kPoolRows; ++r)
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = __fadd_rn(acc[i], __fmul_rn(v[r][i], m[r]));   // h * mask, then the sum
  }
  if (n_splits == 1) {
    if (live) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) pool_store(out + b * out_stride + col + i, acc[i], count, accumulate, divide_by);
    }
    return;
  }
  if (live) {
    float* p = partial + (b * n_splits + s) * hidden_dim + col;
#pragma unroll
    for (int i = 0; i < VEC; ++i) p[i] = acc[i];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) partial_count[b * n_splits + s] = count;
}

__global__ void __launch_bounds__(kPoolThreads)
pool_combine_kernel(const float* __restrict__ partial, const int64_t* __restrict__ partial_count, int hidden_dim,
                    int n_splits, float* __restrict__ out, int64_t out_stride, int accumulate, float divide_by) {
  const int64_t b = blockIdx.y;
  const int col = blockIdx.x * kPoolThreads + threadIdx.x;
  if (col >= hidden_dim) return;
  float sum = 0.f;
  int64_t count = 0;
  for (int s = 0; s < n_splits; ++s) {
    sum = __fadd_rn(sum, partial[(b * n_splits + s) * hidden_dim + col]);
    count += partial_count[b * n_splits + s];
  }
  pool_store(out + b * out_stride + col, sum, count, accumulate, divide_by);
}

static int pool_splits(int64_t n_seq, int64_t seq_len, int col_chunks) {
  const int64_t want = ceil_div((int64_t)4 * num_sms(), n_seq * col_chunks);     // >= 4 CTAs per SM when the input allows
  const int64_t most = ceil_div(seq_len, 2 * kPoolRows);                         // >= 16 rows per split
  return (int)std::max<int64_t>(1, std::min<int64_t>(want, most));
}

template <typename T>
static int pool_launch(const void* hidden, const int64_t* mask, int64_t n_seq, int64_t seq_len, int hidden_dim, float* out,
                       int64_t out_stride, int accumulate, float divide_by, void* ws, int64_t ws_bytes, cudaStream_t st) {
  const bool vec = hidden_dim % PoolVec<T>::N == 0 && ((uintptr_t)hidden % 16) == 0;
  const int per_cta = kPoolThreads * (vec ? PoolVec<T>::N : 1);
  const int col_chunks = (int)ceil_div(hidden_dim, per_cta);
  const int splits = pool_splits(n_seq, seq_len, col_chunks);
  const int rows_per_split = (int)round_up(ceil_div(seq_len, splits), kPoolRows);
  const int n_splits = (int)ceil_div(seq_len, rows_per_split);
  float* partial = nullptr;
  int64_t* pcount = nullptr;
  if (n_splits > 1) {
    Arena a(ws, ws_bytes);
    partial = a.take<float>(n_seq * n_splits * hidden_dim);
    pcount = a.take<int64_t>(n_seq * n_splits);
    if (!a.ok()) { set_error("masked_mean_pool: workspace too small"); return LCREC_ERR_NOMEM; }
  }
  LC_ARG(n_seq <= 65535 && n_splits <= 65535);
  const dim3 grid(col_chunks, n_splits, (unsigned)n_seq);
  if (vec)
    pool_partial_kernel<T, PoolVec<T>::N><<<grid, kPoolThreads, 0, st>>>(
        (const T*)hidden, mask, seq_len, hidden_dim, rows_per_split, n_splits, out, out_stride, accumulate, divide_by,
        partial, pcount);
  else
    pool_partial_kernel<T, 1><<<grid, kPoolThreads, 0, st>>>(
        (const T*)hidden, mask, seq_len, hidden_dim, rows_per_split, n_splits, out, out_stride, accumulate, divide_by,
        partial, pcount);
  LC_LAUNCH_CHECK("pool_partial_kernel");
  if (n_splits > 1) {
    pool_combine_kernel<<<dim3((unsigned)ceil_div(hidden_dim, kPoolThreads), (unsigned)n_seq), kPoolThreads, 0, st>>>(
        partial, pcount, hidden_dim, n_splits, out, out_stride, accumulate, divide_by);
    LC_LAUNCH_CHECK("pool_combine_kernel");
  }
  return LCREC_OK;
}

}  // namespace lcrec

using namespace lcrec;

extern "C" int64_t lcrec_masked_mean_pool_workspace_bytes(int64_t n_seq, int64_t seq_len, int hidden_dim) {
  if (n_seq <= 0 || seq_len <= 0 || hidden_dim <= 0) return 256;
  n_seq = std::min<int64_t>(n_seq, 65535);             // larger batches are processed in slices that reuse it
  const int64_t splits = ceil_div(seq_len, 2 * kPoolRows);            // upper bound of pool_splits()
  return arena_need(n_seq * splits * hidden_dim * (int64_t)sizeof(float)) + arena_need(n_seq * splits * 8);
}

extern "C" int lcrec_masked_mean_pool(const void* hidden, int dtype, const int64_t* mask, int64_t n_seq, int64_t seq_len,
                                      int hidden_dim, float* out, int64_t out_stride, int accumulate, double divide_by,
                                      void* workspace, int64_t workspace_bytes, void* stream) {
  LC_ARG(n_seq >= 0 && seq_len >= 1 && hidden_dim > 0 && out_stride >= hidden_dim);
  LC_ARG(dtype >= 0 && dtype <= 2);
  if (n_seq == 0) return LCREC_OK;
  LC_ARG(out && mask && hidden);
  LC_TRY(lcrec_device_check());
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t esize = dtype == 0 ? 4 : 2;
  for (int64_t b0 = 0; b0 < n_seq; b0 += 65535) {           // grid.z limit
    const int64_t nb = std::min<int64_t>(65535, n_seq - b0);
    const void* h = (const char*)hidden + b0 * seq_len * hidden_dim * esize;
    const int64_t* m = mask + b0 * seq_len;
    float* o = out + b0 * out_stride;
    const float div = (float)divide_by;
    int rc;
    if (dtype == 0) rc = pool_launch<float>(h, m, nb, seq_len, hidden_dim, o, out_stride, accumulate, div, workspace, workspace_bytes, st);
    else if (dtype == 1) rc = pool_launch<__half>(h, m, nb, seq_len, hidden_dim, o, out_stride, accumulate, div, workspace, workspace_bytes, st);
    else rc = pool_launch<__nv_bfloat16>(h, m, nb, seq_len, hidden_dim, o, out_stride, accumulate, div, workspace, workspace_bytes, st);
    LC_TRY(rc);
  }
  return LCREC_OK;
}
