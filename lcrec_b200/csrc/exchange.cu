// Multi-GPU hand-over of PASS-0 results to the owners of the prefix buckets (lcrec_b200/distributed.py).
//
// After PASS 0 every item's codes (L x 8 B) and the residual entering the last level (D x 4 B) travel ONCE to the rank that
// owns its prefix bucket (hash of the first L-1 codes mod world), the owner resolves its buckets locally, and the last-level
// codes travel back.  Round 1 did the partition with torch ops (owner arithmetic, stable argsort, bincount, two index_selects:
// 3.1 ms per 1 M items per rank) and three all_to_all_single calls with host-read split sizes.  Here:
//
//   lcrec_exchange_pack    owner of every item, STABLE partition by owner (per-tile histograms -> one-CTA scan -> scatter with
//                          ballot ranks; ascending item order inside a destination, so that arrival order = ascending global
//                          item id) straight into ONE send buffer of `world` fixed-size slabs [16 B header: row count][records
//                          of L codes + D residual floats]; slot[i] remembers where item i went.  One equal-split
//                          all_to_all_single moves everything (no split sizes on the host).
//   lcrec_exchange_unpack  the received slabs compacted, source rank by source rank, into the (n x L) code table and the
//                          (n x D) residual table lcrec_indexer_resolve works on.
//   lcrec_exchange_pack_last / lcrec_exchange_scatter_last   the way back: last-level codes in slabs along the same routes,
//                          scattered into this rank's table through slot[].
// All of it is integer / byte shuffling bound by HBM and NVLink; nothing is computed.
#include <algorithm>

#include "common.cuh"

namespace lcrec {

constexpr int kExThreads = 256;
constexpr int kExTile = 1024;              // items per CTA
constexpr int kExMaxWorld = 16;
constexpr int kExHeader = 16;              // bytes in front of every slab: int64 row count, int64 overflow flag

struct ExRadix { long long k[LCREC_MAX_LEVELS]; int L; };

__device__ __forceinline__ int ex_owner(const int64_t* __restrict__ row, const ExRadix& r, int world) {
  unsigned long long prefix = 0;
  for (int l = 0; l < r.L - 1; ++l) prefix = prefix * (unsigned long long)r.k[l] + (unsigned long long)row[l];
  const unsigned long long mixed = ((prefix * 0x9E3779B97F4A7C15ull) & 0x7FFFFFFFFFFFFFFFull) >> 24;     // == distributed.bucket_owner
  return (int)(mixed % (unsigned long long)world);
}

__global__ void __launch_bounds__(kExThreads) ex_hist_kernel(const int64_t* __restrict__ codes, int64_t n, ExRadix radix, int world,
                                                             unsigned char* __restrict__ owner, int* __restrict__ hist /* world x n_tiles */,
                                                             int n_tiles) {
  __shared__ int h[kExMaxWorld];
  if (threadIdx.x < kExMaxWorld) h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kExTile;
  for (int j = threadIdx.x; j < kExTile; j += kExThreads) {
    const int64_t i = base + j;
    if (i < n) {
      const int o = ex_owner(codes + i * radix.L, radix, world);
      owner[i] = (unsigned char)o;
      atomicAdd(&h[o], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x < world) hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan over the tiles of every destination; row counts into the slab headers
__global__ void ex_scan_kernel(int* __restrict__ hist, int n_tiles, int world, unsigned char* __restrict__ send, int64_t slab_bytes,
                               int64_t slab_rows, int64_t* __restrict__ counts_out) {
  const int b = blockIdx.x;                  // one CTA per destination
  __shared__ long long carry;
  __shared__ long long wsum[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  int* hrow = hist + (size_t)b * n_tiles;
  for (int t0 = 0; t0 < n_tiles; t0 += blockDim.x) {
    const int t = t0 + threadIdx.x;
    const long long v = t < n_tiles ? hrow[t] : 0;
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      long long w = threadIdx.x < (blockDim.x >> 5) ? wsum[threadIdx.x] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += y; }
      wsum[threadIdx.x] = w;
    }
    __syncthreads();
    const long long before = carry + (threadIdx.x >= 32 ? wsum[(threadIdx.x >> 5) - 1] : 0) + x - v;
    if (t < n_tiles) hrow[t] = (int)before;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int64_t* hdr = reinterpret_cast<int64_t*>(send + (size_t)b * slab_bytes);
    hdr[0] = carry;
    hdr[1] = carry > slab_rows ? 1 : 0;
    if (counts_out) counts_out[b] = carry;
  }
}

__global__ void __launch_bounds__(kExThreads) ex_scatter_kernel(const int64_t* __restrict__ codes, const float* __restrict__ resid, int64_t n,
                                                                int L, int D, int world, const unsigned char* __restrict__ owner,
                                                                const int* __restrict__ tile_base, int n_tiles, unsigned char* __restrict__ send,
                                                                int64_t slab_bytes, int64_t slab_rows, int rec_bytes, int32_t* __restrict__ slot) {
  __shared__ int wcount[kExThreads / 32][kExMaxWorld];
  __shared__ int run[kExMaxWorld];                     // items of this tile already placed per destination (previous passes)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < kExMaxWorld) run[threadIdx.x] = 0;
  const int64_t base = (int64_t)blockIdx.x * kExTile;
  for (int pass = 0; pass < kExTile / kExThreads; ++pass) {       // item order inside the tile: pass-major, then thread id
    __syncthreads();
    const int64_t i = base + pass * kExThreads + threadIdx.x;
    const int o = i < n ? (int)owner[i] : -1;
    int rank_in_warp = 0;
    for (int b = 0; b < world; ++b) {
      const unsigned m = __ballot_sync(0xffffffffu, o == b);
      if (o == b) rank_in_warp = __popc(m & ((1u << lane) - 1u));
      if (lane == 0) wcount[warp][b] = __popc(m);
    }
    __syncthreads();
    if (o >= 0) {
      int before = run[o];
      for (int w = 0; w < warp; ++w) before += wcount[w][o];
      const int64_t dest = (int64_t)tile_base[(size_t)o * n_tiles + blockIdx.x] + before + rank_in_warp;
      slot[i] = (int32_t)(dest < slab_rows ? (int64_t)o * slab_rows + dest : -1);
      if (dest < slab_rows) {
        unsigned char* rec = send + (size_t)o * slab_bytes + kExHeader + (size_t)dest * rec_bytes;
        int64_t* rc = reinterpret_cast<int64_t*>(rec);
        for (int l = 0; l < L; ++l) rc[l] = codes[i * L + l];
        float* rr = reinterpret_cast<float*>(rec + (((size_t)L * 8 + 15) & ~(size_t)15));
        if ((D & 3) == 0) {
          const float4* src = reinterpret_cast<const float4*>(resid + i * D);
          float4* dst = reinterpret_cast<float4*>(rr);
          for (int d = 0; d < D / 4; ++d) dst[d] = src[d];
        } else {
          for (int d = 0; d < D; ++d) rr[d] = resid[i * D + d];
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < world) {
      int t = 0;
      for (int w = 0; w < kExThreads / 32; ++w) t += wcount[w][threadIdx.x];
      run[threadIdx.x] += t;
    }
  }
}

// recv: world slabs (source rank order).  out rows = concatenation of the sources' records.
__global__ void __launch_bounds__(kExThreads) ex_unpack_kernel(const unsigned char* __restrict__ recv, int world, int64_t slab_bytes, int64_t slab_rows,
                                                               int rec_bytes, int L, int D, int64_t* __restrict__ codes, float* __restrict__ resid,
                                                               int64_t cap_rows) {
  int64_t start[kExMaxWorld + 1];
  start[0] = 0;
  for (int s = 0; s < world; ++s) {
    const int64_t c = min(reinterpret_cast<const int64_t*>(recv + (size_t)s * slab_bytes)[0], slab_rows);
    start[s + 1] = start[s] + c;
  }
  const int64_t total = min(start[world], cap_rows);
  for (int64_t r = (int64_t)blockIdx.x * kExThreads + threadIdx.x; r < total; r += (int64_t)gridDim.x * kExThreads) {
    int s = 0;
    while (r >= start[s + 1]) ++s;
    const unsigned char* rec = recv + (size_t)s * slab_bytes + kExHeader + (size_t)(r - start[s]) * rec_bytes;
    const int64_t* rc = reinterpret_cast<const int64_t*>(rec);
    for (int l = 0; l < L; ++l) codes[r * L + l] = rc[l];
    const float* rr = reinterpret_cast<const float*>(rec + (((size_t)L * 8 + 15) & ~(size_t)15));
    if ((D & 3) == 0) {
      const float4* src = reinterpret_cast<const float4*>(rr);
      float4* dst = reinterpret_cast<float4*>(resid + r * D);
      for (int d = 0; d < D / 4; ++d) dst[d] = src[d];
    } else {
      for (int d = 0; d < D; ++d) resid[r * D + d] = rr[d];
    }
  }
}

// way back: the owner puts the last-level code of its r-th row (source s, position p) into slab s, position p
__global__ void __launch_bounds__(kExThreads) ex_pack_last_kernel(const int64_t* __restrict__ codes, int L, const unsigned char* __restrict__ recv,
                                                                  int world, int64_t slab_bytes, int64_t slab_rows, int64_t* __restrict__ back /* world x slab_rows */) {
  int64_t start[kExMaxWorld + 1];
  start[0] = 0;
  for (int s = 0; s < world; ++s) start[s + 1] = start[s] + min(reinterpret_cast<const int64_t*>(recv + (size_t)s * slab_bytes)[0], slab_rows);
  const int64_t total = start[world];
  for (int64_t r = (int64_t)blockIdx.x * kExThreads + threadIdx.x; r < total; r += (int64_t)gridDim.x * kExThreads) {
    int s = 0;
    while (r >= start[s + 1]) ++s;
    back[(size_t)s * slab_rows + (r - start[s])] = codes[r * L + (L - 1)];
  }
}

__global__ void __launch_bounds__(kExThreads) ex_scatter_last_kernel(const int64_t* __restrict__ back_recv, const int32_t* __restrict__ slot, int64_t n,
                                                                     int L, int64_t* __restrict__ codes) {
  for (int64_t i = (int64_t)blockIdx.x * kExThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kExThreads) {
    const int32_t s = slot[i];
    if (s >= 0) codes[i * L + (L - 1)] = back_recv[s];
  }
}

}  // namespace lcrec

using namespace lcrec;

// codes padded to 16 bytes so that the residual part of every record is float4-aligned (3 levels: 24 -> 32 bytes)
extern "C" int64_t lcrec_exchange_record_bytes(int n_levels, int e_dim) { return round_up((int64_t)n_levels * 8, 16) + round_up((int64_t)e_dim * 4, 16); }
extern "C" int64_t lcrec_exchange_slab_bytes(int64_t slab_rows, int n_levels, int e_dim) {
  return round_up(kExHeader + slab_rows * lcrec_exchange_record_bytes(n_levels, e_dim), 256);
}
extern "C" int64_t lcrec_exchange_workspace_bytes(int64_t n, int world) {
  const int64_t tiles = ceil_div(std::max<int64_t>(n, 1), kExTile);
  return arena_need(std::max<int64_t>(n, 1)) + arena_need(4 * tiles * world) + 1024;
}

// Partition this rank's n items by bucket owner into `send` (world slabs of lcrec_exchange_slab_bytes each).
// slot (n int32, device): where item i went (owner * slab_rows + position; -1 = its slab overflowed, header[1] of that slab is set);
// counts_dev (nullable, world int64): rows per destination.
extern "C" int lcrec_exchange_pack(const int64_t* codes, const float* resid, int64_t n, int n_levels, int e_dim, const int32_t* n_codes,
                                   int world, int64_t slab_rows, void* send, int32_t* slot, int64_t* counts_dev, void* ws, int64_t ws_bytes,
                                   void* stream) {
  LC_ARG(n >= 0 && n_levels >= 1 && n_levels <= LCREC_MAX_LEVELS && e_dim > 0 && n_codes && world >= 1 && world <= kExMaxWorld);
  LC_ARG(slab_rows >= 1 && send && (n == 0 || (codes && resid && slot)) && n < ((int64_t)1 << 31) && world * slab_rows < ((int64_t)1 << 31));
  LC_TRY(lcrec_device_check());
  cudaStream_t st = (cudaStream_t)stream;
  const int n_tiles = (int)ceil_div(std::max<int64_t>(n, 1), kExTile);
  Arena ar(ws, ws_bytes);
  unsigned char* owner = ar.take<unsigned char>(std::max<int64_t>(n, 1));
  int* hist = ar.take<int>((int64_t)n_tiles * world);
  if (!ar.ok()) { set_error("exchange_pack: workspace too small"); return LCREC_ERR_NOMEM; }
  ExRadix radix{};
  radix.L = n_levels;
  for (int l = 0; l < n_levels; ++l) radix.k[l] = n_codes[l];
  const int64_t slab_bytes = lcrec_exchange_slab_bytes(slab_rows, n_levels, e_dim);
  ex_hist_kernel<<<n_tiles, kExThreads, 0, st>>>(codes, n, radix, world, owner, hist, n_tiles);
  LC_LAUNCH_CHECK("ex_hist_kernel");
  ex_scan_kernel<<<world, 256, 0, st>>>(hist, n_tiles, world, (unsigned char*)send, slab_bytes, slab_rows, counts_dev);
  LC_LAUNCH_CHECK("ex_scan_kernel");
  if (n > 0) {
    ex_scatter_kernel<<<n_tiles, kExThreads, 0, st>>>(codes, resid, n, n_levels, e_dim, world, owner, hist, n_tiles, (unsigned char*)send,
                                                      slab_bytes, slab_rows, (int)lcrec_exchange_record_bytes(n_levels, e_dim), slot);
    LC_LAUNCH_CHECK("ex_scatter_kernel");
  }
  return LCREC_OK;
}

// recv = `world` slabs as received (source rank order) -> codes (cap_rows x L), resid (cap_rows x D); the row count is the sum of
// the slab headers (the caller reads them: `world` int64 at recv + s * slab_bytes).
extern "C" int lcrec_exchange_unpack(const void* recv, int world, int64_t slab_rows, int n_levels, int e_dim, int64_t* codes, float* resid,
                                     int64_t cap_rows, void* stream) {
  LC_ARG(recv && world >= 1 && world <= kExMaxWorld && slab_rows >= 1 && n_levels >= 1 && e_dim > 0 && cap_rows >= 0);
  LC_TRY(lcrec_device_check());
  if (cap_rows == 0) return LCREC_OK;
  LC_ARG(codes && resid);
  const int64_t slab_bytes = lcrec_exchange_slab_bytes(slab_rows, n_levels, e_dim);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(cap_rows, kExThreads), (int64_t)num_sms() * 8));
  ex_unpack_kernel<<<grid, kExThreads, 0, (cudaStream_t)stream>>>((const unsigned char*)recv, world, slab_bytes, slab_rows,
                                                                (int)lcrec_exchange_record_bytes(n_levels, e_dim), n_levels, e_dim, codes, resid, cap_rows);
  LC_LAUNCH_CHECK("ex_unpack_kernel");
  return LCREC_OK;
}

// The way back.  pack_last: owner side, codes = the resolved table in unpack order, back = world x slab_rows int64.
// scatter_last: origin side, back_recv = world x slab_rows int64 as received, slot from lcrec_exchange_pack.
extern "C" int lcrec_exchange_pack_last(const int64_t* codes, int n_levels, const void* recv, int world, int64_t slab_rows, int e_dim,
                                        int64_t* back, int64_t n_rows_hint, void* stream) {
  LC_ARG(recv && back && world >= 1 && world <= kExMaxWorld && slab_rows >= 1 && n_levels >= 1 && (n_rows_hint == 0 || codes));
  LC_TRY(lcrec_device_check());
  if (n_rows_hint == 0) return LCREC_OK;
  const int64_t slab_bytes = lcrec_exchange_slab_bytes(slab_rows, n_levels, e_dim);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_rows_hint, kExThreads), (int64_t)num_sms() * 8));
  ex_pack_last_kernel<<<grid, kExThreads, 0, (cudaStream_t)stream>>>(codes, n_levels, (const unsigned char*)recv, world, slab_bytes, slab_rows, back);
  LC_LAUNCH_CHECK("ex_pack_last_kernel");
  return LCREC_OK;
}

extern "C" int lcrec_exchange_scatter_last(const int64_t* back_recv, const int32_t* slot, int64_t n, int n_levels, int64_t* codes, void* stream) {
  LC_ARG(n >= 0 && n_levels >= 1);
  LC_TRY(lcrec_device_check());
  if (n == 0) return LCREC_OK;
  LC_ARG(back_recv && slot && codes);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, kExThreads), (int64_t)num_sms() * 8));
  ex_scatter_last_kernel<<<grid, kExThreads, 0, (cudaStream_t)stream>>>(back_recv, slot, n, n_levels, codes);
  LC_LAUNCH_CHECK("ex_scatter_last_kernel");
  return LCREC_OK;
}
