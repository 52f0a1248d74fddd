// Ordered per-code row sums: the CTA walks an assignment vector in item order, compacts the rows assigned to code k into a
// shared-memory list that keeps their order (warp ballots + prefix) and thread d adds x[row][d] to acc[d] in exactly that
// order - the summation order of a sequential CPU scatter-add (torch index_add_), bit-reproducible, no floating-point atomics.
#pragma once
#include <stdint.h>

namespace lcrec {

constexpr int kSegThreads = 128;
constexpr int kSegList = 2048;     // members buffered between flushes

struct SegSumSmem {
  int list[kSegList];
  int warp_total[kSegThreads / 32];
};

// acc: e_dim floats in shared memory, zeroed here.  Returns the number of rows assigned to k (same value in every thread).
// Ends with a __syncthreads(): acc is complete and visible to the whole CTA.  n < 2^31.
__device__ __forceinline__ int64_t ordered_code_sum(const float* __restrict__ x, const int64_t* __restrict__ indices, int64_t n,
                                                    int e_dim, int k, float* acc, SegSumSmem& sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int d = tid; d < e_dim; d += kSegThreads) acc[d] = 0.f;
  int filled = 0;
  int64_t count = 0;
  __syncthreads();

  auto flush = [&]() {
    for (int d = tid; d < e_dim; d += kSegThreads) {
      float a = acc[d];
      int j = 0;
      for (; j + 4 <= filled; j += 4) {          // loads issued together, additions in list order
        const float v0 = x[(int64_t)sm.list[j] * e_dim + d];
        const float v1 = x[(int64_t)sm.list[j + 1] * e_dim + d];
        const float v2 = x[(int64_t)sm.list[j + 2] * e_dim + d];
        const float v3 = x[(int64_t)sm.list[j + 3] * e_dim + d];
        a = __fadd_rn(a, v0); a = __fadd_rn(a, v1); a = __fadd_rn(a, v2); a = __fadd_rn(a, v3);
      }
      for (; j < filled; ++j) a = __fadd_rn(a, x[(int64_t)sm.list[j] * e_dim + d]);
      acc[d] = a;
    }
  };

  for (int64_t base = 0; base < n; base += kSegThreads) {
    const int64_t i = base + tid;
    const bool hit = i < n && indices[i] == (int64_t)k;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) sm.warp_total[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kSegThreads / 32; ++w) {
      const int t = sm.warp_total[w];
      before += w < warp ? t : 0;
      total += t;
    }
    if (hit) sm.list[filled + before + __popc(m & ((1u << lane) - 1u))] = (int)i;
    filled += total;
    count += total;
    __syncthreads();
    if (filled > kSegList - kSegThreads) {
      flush();
      filled = 0;
      __syncthreads();
    }
  }
  flush();
  __syncthreads();
  return count;
}

}  // namespace lcrec
