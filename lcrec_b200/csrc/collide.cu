// Collision bookkeeping on the device: pack code tuples, LSD radix sort, run detection, CSR groups.
//
// Replaces the Python string/dict machinery of the reference: check_collision / get_collision_item /
// get_indices_count (index/generate_indices.py:18-42) and the set-of-strings collision rate of
// Trainer._valid_epoch (index/trainer.py:141-150).  Items whose (c_0..c_{L-1}) tuples are identical
// form a collision group.  The tuple is packed into one u64 key (ceil(log2 K_l) bits per level),
// (key, item) pairs are sorted with a stable 8-bit LSD radix sort (so members of a group come out in
// ascending item order, like the reference's lists) and groups are emitted in CSR form.
// HBM-bound: per pass one histogram read and one scatter read + write of 12 B/item.
#include "common.cuh"

namespace lcrec {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kItemsPerThread = 16;
constexpr int kTile = kSortThreads * kItemsPerThread;   // 4096 items per CTA

struct PackArgs { int n_levels; int shift[LCREC_MAX_LEVELS]; };

__global__ void pack_codes_kernel(const int64_t* __restrict__ codes, int64_t n, PackArgs pa,
                                  uint64_t* __restrict__ keys, uint32_t* __restrict__ items) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t k = 0;
    for (int l = 0; l < pa.n_levels; ++l) k |= (uint64_t)codes[i * pa.n_levels + l] << pa.shift[l];
    keys[i] = k;
    items[i] = (uint32_t)i;
  }
}

// per-CTA digit histogram -> hist[digit * n_tiles + tile]
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint64_t* __restrict__ keys, int64_t n,
                                                                  int shift, uint32_t* __restrict__ hist, int n_tiles) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kTile;
#pragma unroll 4
  for (int j = 0; j < kItemsPerThread; ++j) {
    const int64_t i = base + j * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of a u32 array with one CTA (array length up to a few million)
__global__ void __launch_bounds__(1024) scan_single_kernel(uint32_t* __restrict__ data, int64_t n) {
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < n; base += 4096) {
    const int64_t i0 = base + (int64_t)threadIdx.x * 4;
    uint32_t v[4]; uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[j] = (i0 + j < n) ? data[i0 + j] : 0u; s += v[j]; }
    uint32_t inc = s;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_tot[lane];
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
      warp_tot[lane] = w;
    }
    __syncthreads();
    uint32_t excl = carry + (warp > 0 ? warp_tot[warp - 1] : 0u) + (inc - s);
#pragma unroll
    for (int j = 0; j < 4; ++j) { if (i0 + j < n) data[i0 + j] = excl; excl += v[j]; }
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_tot[31];
    __syncthreads();
  }
}

// stable scatter: warp w of the CTA owns items [w*512, (w+1)*512) of the tile, 16 rounds of 32
// consecutive items; __match_any_sync ranks equal digits inside a round in lane order.
__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(
    const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ items_in, int64_t n, int shift,
    const uint32_t* __restrict__ hist_scanned, int n_tiles, uint64_t* __restrict__ keys_out,
    uint32_t* __restrict__ items_out) {
  __shared__ uint32_t wcount[kSortWarps][256];
  __shared__ uint32_t gbase[256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&wcount[0][0])[i] = 0;
  gbase[threadIdx.x] = hist_scanned[(size_t)threadIdx.x * n_tiles + blockIdx.x];
  __syncthreads();
  const int64_t wbase = (int64_t)blockIdx.x * kTile + (int64_t)warp * (32 * kItemsPerThread);
  uint64_t key[kItemsPerThread]; uint32_t item[kItemsPerThread]; uint32_t rank[kItemsPerThread];
#pragma unroll
  for (int j = 0; j < kItemsPerThread; ++j) {
    const int64_t i = wbase + j * 32 + lane;
    const bool valid = i < n;
    key[j] = valid ? keys_in[i] : 0ull;
    item[j] = valid ? items_in[i] : 0u;
    const uint32_t digit = valid ? (uint32_t)((key[j] >> shift) & 255u) : 256u;
    const uint32_t peers = __match_any_sync(0xffffffffu, digit);
    const uint32_t below = __popc(peers & ((1u << lane) - 1u));
    uint32_t prev = 0;
    if (valid) prev = wcount[warp][digit];
    __syncwarp();
    if (valid && below == 0) wcount[warp][digit] = prev + __popc(peers);
    __syncwarp();
    rank[j] = prev + below;
  }
  __syncthreads();
  // exclusive prefix over warps per digit (thread d handles digit d)
  {
    uint32_t run = gbase[threadIdx.x];
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) { const uint32_t c = wcount[w][threadIdx.x]; wcount[w][threadIdx.x] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kItemsPerThread; ++j) {
    const int64_t i = wbase + j * 32 + lane;
    if (i < n) {
      const uint32_t digit = (uint32_t)((key[j] >> shift) & 255u);
      const uint32_t pos = wcount[warp][digit] + rank[j];
      keys_out[pos] = key[j];
      items_out[pos] = item[j];
    }
  }
}

// flags on the sorted keys: low 32 bits = item belongs to a run of length >= 2, high 32 bits = item
// is the head of such a run.  Packed so that one scan yields member positions and group ids.
__global__ void run_flags_kernel(const uint64_t* __restrict__ keys, int64_t n, uint64_t* __restrict__ flags, int drop_bits) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys[i] >> drop_bits;        // drop_bits > 0: runs of equal PREFIX (low digits ignored)
    const bool eq_prev = i > 0 && (keys[i - 1] >> drop_bits) == k;
    const bool eq_next = i + 1 < n && (keys[i + 1] >> drop_bits) == k;
    flags[i] = (uint64_t)((eq_prev || eq_next) ? 1u : 0u) | ((uint64_t)((!eq_prev && eq_next) ? 1u : 0u) << 32);
  }
}

// exclusive scan of packed u64 (two independent 32-bit counters), 3 kernels
constexpr int kScanTile = 2048;
__global__ void __launch_bounds__(256) scan_reduce_kernel(const uint64_t* __restrict__ in, int64_t n, uint64_t* __restrict__ sums) {
  __shared__ uint64_t red[8];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  uint64_t s = 0;
  for (int j = threadIdx.x; j < kScanTile; j += 256) if (base + j < n) s += in[base + j];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { uint64_t t = 0; for (int w = 0; w < 8; ++w) t += red[w]; sums[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024) scan_sums_kernel(uint64_t* __restrict__ sums, int64_t nb, uint64_t* __restrict__ total) {
  __shared__ uint64_t warp_tot[32];
  __shared__ uint64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < nb; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const uint64_t v = i < nb ? sums[i] : 0ull;
    uint64_t inc = v;
    for (int o = 1; o < 32; o <<= 1) { const uint64_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint64_t w = warp_tot[lane];
      for (int o = 1; o < 32; o <<= 1) { const uint64_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
      warp_tot[lane] = w;
    }
    __syncthreads();
    const uint64_t excl = carry + (warp > 0 ? warp_tot[warp - 1] : 0ull) + (inc - v);
    if (i < nb) sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_tot[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
// applies the scan and emits the CSR: members[pos] = item, offsets[gid] = pos at group heads
__global__ void __launch_bounds__(256) emit_groups_kernel(const uint64_t* __restrict__ flags, const uint32_t* __restrict__ items,
                                                          int64_t n, const uint64_t* __restrict__ sums,
                                                          int64_t* __restrict__ offsets, int64_t* __restrict__ members) {
  __shared__ uint64_t warp_tot[8];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t carry = sums[blockIdx.x];
  for (int r = 0; r < kScanTile / 256; ++r) {
    const int64_t i = base + r * 256 + threadIdx.x;
    const uint64_t v = i < n ? flags[i] : 0ull;
    uint64_t inc = v;
    for (int o = 1; o < 32; o <<= 1) { const uint64_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    uint64_t wpre = 0, tot = 0;
    for (int w = 0; w < 8; ++w) { if (w < warp) wpre += warp_tot[w]; tot += warp_tot[w]; }
    const uint64_t excl = carry + wpre + (inc - v);
    if (i < n && (v & 0xffffffffull)) {
      const int64_t pos = (int64_t)(excl & 0xffffffffull);
      members[pos] = items[i];
      if (v >> 32) offsets[(int64_t)(excl >> 32)] = pos;
    }
    carry += tot;
    __syncthreads();
  }
}
// counts: [n_unique, n_groups, n_colliding_rows, max_multiplicity]; also closes the CSR
__global__ void finish_counts_kernel(const uint64_t* __restrict__ total, int64_t n, int64_t* __restrict__ offsets,
                                     int64_t* __restrict__ counts) {
  const int64_t rows = (int64_t)(*total & 0xffffffffull), groups = (int64_t)(*total >> 32);
  offsets[groups] = rows;
  counts[0] = n - (rows - groups);
  counts[1] = groups;
  counts[2] = rows;
  counts[3] = groups > 0 ? 0 : (n > 0 ? 1 : 0);
}
__global__ void max_mult_kernel(const int64_t* offsets, const int64_t* counts_in, int64_t* counts) {
  const int64_t groups = counts_in[1];
  long long m = 0;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x)
    m = max(m, (long long)(offsets[g + 1] - offsets[g]));
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax((long long*)(counts + 3), m);
}


// ---------------------------------------------------------------------------- collisions inside prefix segments
// While the rounds rewrite the last level only, two items can collide only if they share a prefix segment
// (lcrec_prefix_segments, built once).  Per round: sort the members of every segment by (code of `level`, item);
// runs of equal code with >= 2 members are the collision groups - no global re-sort.
// Group slots and member slots are claimed TOGETHER with one 64-bit atomicAdd (groups in the high, rows in the low
// word), which serialises the claims: the groups of claim k start exactly where the rows of claim k-1 end, so the
// CSR stays consistent (offsets[g + 1] closes group g) although segments finish in arbitrary order (groups are
// independent problems: their order changes no result; members of a group stay in ascending item order).
// ctl: [0] packed (groups << 32 | rows) cursor, [1] max multiplicity, [2] number of big segments, [3] fallback flag.
constexpr int kSegMax = 1024;     // largest segment sorted on chip; anything bigger raises the fallback flag
__global__ void __launch_bounds__(256) segment_collisions_kernel(const int64_t* __restrict__ codes, int n_levels, int level,
                                                                 const int64_t* __restrict__ seg_offsets,
                                                                 const int64_t* __restrict__ seg_members,
                                                                 const int64_t* __restrict__ n_segs_dev,
                                                                 int64_t* __restrict__ offsets, int64_t* __restrict__ members,
                                                                 unsigned long long* __restrict__ ctl, int32_t* __restrict__ big_list,
                                                                 const int32_t* __restrict__ active_in, const int* __restrict__ n_active_in,
                                                                 int32_t* __restrict__ active_out, int* __restrict__ n_active_out) {
  const int lane = threadIdx.x & 31;
  // a segment without a collision this round cannot have one later (nothing in it changes): from the second check
  // on only the segments that collided last time (active_in) are examined, and those that still collide are passed on
  const int64_t n_segs = active_in ? (int64_t)*n_active_in : *n_segs_dev;
  long long local_max = 0;
  for (int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; w < n_segs; w += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const int64_t sg = active_in ? (int64_t)active_in[w] : w;
    const int64_t beg = seg_offsets[sg];
    const int m = (int)min((int64_t)(kSegMax + 1), seg_offsets[sg + 1] - beg);
    if (m > 32) {
      if (lane == 0) {
        if (m > kSegMax) atomicExch(ctl + 3, 1ull);
        else {
          big_list[atomicAdd(ctl + 2, 1ull)] = (int32_t)sg;
          if (active_out) active_out[atomicAdd(n_active_out, 1)] = (int32_t)sg;      // big segments stay on the list
        }
      }
      continue;
    }
    unsigned long long key = ~0ull;
    if (lane < m) {
      const int64_t item = seg_members[beg + lane];
      key = ((unsigned long long)codes[item * n_levels + level] << 32) | (unsigned long long)item;
    }
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1)          // bitonic sort of the 32 keys held by the warp
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, j);
        const bool take_min = ((lane & k) == 0) == ((lane & j) == 0);
        key = take_min ? (key < other ? key : other) : (key > other ? key : other);
      }
    const unsigned code = (unsigned)(key >> 32);
    const unsigned prev = __shfl_up_sync(0xffffffffu, code, 1), next = __shfl_down_sync(0xffffffffu, code, 1);
    const bool valid = lane < m;
    const bool eq_prev = valid && lane > 0 && prev == code;
    const bool eq_next = valid && lane + 1 < m && next == code;
    const unsigned in_run = __ballot_sync(0xffffffffu, eq_prev || eq_next);
    const unsigned heads = __ballot_sync(0xffffffffu, !eq_prev && eq_next);
    if (in_run == 0) continue;
    unsigned long long base = 0;
    if (lane == 0) {
      base = atomicAdd(ctl, ((unsigned long long)__popc(heads) << 32) | (unsigned long long)__popc(in_run));
      if (active_out) active_out[atomicAdd(n_active_out, 1)] = (int32_t)sg;
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    const int64_t rb = (int64_t)(base & 0xffffffffull), gb = (int64_t)(base >> 32);
    const unsigned below = (1u << lane) - 1u;
    if (eq_prev || eq_next) {
      const int64_t pos = rb + __popc(in_run & below);
      members[pos] = (int64_t)(key & 0xffffffffull);
      if (!eq_prev) {
        offsets[gb + __popc(heads & below)] = pos;
        const unsigned after = lane == 31 ? 0u : ((in_run & ~heads) >> (lane + 1));    // followers of this head
        local_max = max(local_max, (long long)__ffs(~after));
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
  if (lane == 0 && local_max > 0) atomicMax((long long*)(ctl + 1), local_max);
}

// segments of 33 .. kSegMax members: one CTA per segment, bitonic sort in shared memory
__global__ void __launch_bounds__(256) segment_collisions_big_kernel(const int64_t* __restrict__ codes, int n_levels, int level,
                                                                     const int64_t* __restrict__ seg_offsets,
                                                                     const int64_t* __restrict__ seg_members,
                                                                     int64_t* __restrict__ offsets, int64_t* __restrict__ members,
                                                                     unsigned long long* __restrict__ ctl, const int32_t* __restrict__ big_list) {
  __shared__ unsigned long long key[kSegMax];
  __shared__ unsigned warp_rows[8], warp_heads[8];
  __shared__ unsigned long long s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_big = (int)ctl[2];
  long long local_max = 0;
  for (int w = blockIdx.x; w < n_big; w += gridDim.x) {
    const int64_t sg = big_list[w];
    const int64_t beg = seg_offsets[sg];
    const int m = (int)(seg_offsets[sg + 1] - beg);
    __syncthreads();
    for (int i = tid; i < kSegMax; i += 256) {
      unsigned long long k = ~0ull;
      if (i < m) { const int64_t item = seg_members[beg + i]; k = ((unsigned long long)codes[item * n_levels + level] << 32) | (unsigned long long)item; }
      key[i] = k;
    }
    __syncthreads();
    for (int k = 2; k <= kSegMax; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < kSegMax; i += 256) {
          const int p = i ^ j;
          if (p > i) {
            const unsigned long long a = key[i], b = key[p];
            if (((i & k) == 0) == (a > b)) { key[i] = b; key[p] = a; }
          }
        }
        __syncthreads();
      }
    // thread t owns positions 4t .. 4t+3
    unsigned rows = 0, heads = 0; bool run[4], head[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = 4 * tid + e;
      const unsigned c = (unsigned)(key[min(i, kSegMax - 1)] >> 32);
      const bool eq_prev = i < m && i > 0 && (unsigned)(key[i - 1] >> 32) == c;
      const bool eq_next = i + 1 < m && (unsigned)(key[i + 1] >> 32) == c;
      run[e] = eq_prev || eq_next; head[e] = !eq_prev && eq_next;
      rows += run[e]; heads += head[e];
      if (head[e]) { int len = 1; while (i + len < m && (unsigned)(key[i + len] >> 32) == c) ++len; local_max = max(local_max, (long long)len); }
    }
    unsigned rinc = rows, hinc = heads;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned a = __shfl_up_sync(0xffffffffu, rinc, o), b = __shfl_up_sync(0xffffffffu, hinc, o);
      if (lane >= o) { rinc += a; hinc += b; }
    }
    if (lane == 31) { warp_rows[warp] = rinc; warp_heads[warp] = hinc; }
    __syncthreads();
    unsigned rpre = 0, hpre = 0, rtot = 0, htot = 0;
    for (int q = 0; q < 8; ++q) { if (q < warp) { rpre += warp_rows[q]; hpre += warp_heads[q]; } rtot += warp_rows[q]; htot += warp_heads[q]; }
    if (tid == 0 && rtot > 0) s_base = atomicAdd(ctl, ((unsigned long long)htot << 32) | (unsigned long long)rtot);
    __syncthreads();
    if (rtot == 0) continue;
    int64_t pos = (int64_t)(s_base & 0xffffffffull) + rpre + (rinc - rows);
    int64_t gid = (int64_t)(s_base >> 32) + hpre + (hinc - heads);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (!run[e]) continue;
      members[pos] = (int64_t)(key[4 * tid + e] & 0xffffffffull);
      if (head[e]) offsets[gid++] = pos;
      ++pos;
    }
  }
  for (int o = 16; o > 0; o >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
  if (lane == 0 && local_max > 0) atomicMax((long long*)(ctl + 1), local_max);
}

// counts: [n_unique, n_groups, rows, max_multiplicity, -, fallback]; closes the CSR
__global__ void segment_finish_kernel(const unsigned long long* __restrict__ ctl, int64_t n, int64_t* __restrict__ offsets,
                                      int64_t* __restrict__ counts) {
  const int64_t rows = (int64_t)(ctl[0] & 0xffffffffull), groups = (int64_t)(ctl[0] >> 32);
  offsets[groups] = rows;
  counts[0] = n - (rows - groups);
  counts[1] = groups;
  counts[2] = rows;
  counts[3] = groups > 0 ? (int64_t)ctl[1] : (n > 0 ? 1 : 0);
  counts[5] = (int64_t)ctl[3];
}

struct SortPlan { PackArgs pa; int total_bits; int n_tiles; };

static int make_plan(int64_t n, int n_levels, const int32_t* n_codes, SortPlan* plan) {
  int bits = 0;
  plan->pa.n_levels = n_levels;
  for (int l = n_levels - 1; l >= 0; --l) {   // level 0 most significant
    int b = 1;
    while ((1 << b) < n_codes[l]) ++b;
    plan->pa.shift[l] = bits;
    bits += b;
  }
  if (bits > 64) { set_error("collisions: packed code needs %d bits (> 64)", bits); return LCREC_ERR_UNSUPPORTED; }
  if (n >= ((int64_t)1 << 32)) { set_error("collisions: more than 2^32-1 items"); return LCREC_ERR_UNSUPPORTED; }
  plan->total_bits = bits;
  plan->n_tiles = (int)ceil_div(n, kTile);
  return LCREC_OK;
}

static int sort_impl(const int64_t* codes, int64_t n, const SortPlan& plan, uint64_t* k0, uint32_t* i0, uint64_t* k1,
                     uint32_t* i1, uint32_t* hist, cudaStream_t st, uint64_t** keys_sorted, uint32_t** items_sorted) {
  const int blocks = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)num_sms() * 8);
  pack_codes_kernel<<<blocks, 256, 0, st>>>(codes, n, plan.pa, k0, i0);
  LC_LAUNCH_CHECK("pack_codes_kernel");
  uint64_t *ka = k0, *kb = k1; uint32_t *ia = i0, *ib = i1;
  for (int shift = 0; shift < plan.total_bits; shift += 8) {
    radix_hist_kernel<<<plan.n_tiles, kSortThreads, 0, st>>>(ka, n, shift, hist, plan.n_tiles);
    LC_LAUNCH_CHECK("radix_hist_kernel");
    scan_single_kernel<<<1, 1024, 0, st>>>(hist, (int64_t)256 * plan.n_tiles);
    LC_LAUNCH_CHECK("scan_single_kernel");
    radix_scatter_kernel<<<plan.n_tiles, kSortThreads, 0, st>>>(ka, ia, n, shift, hist, plan.n_tiles, kb, ib);
    LC_LAUNCH_CHECK("radix_scatter_kernel");
    std::swap(ka, kb); std::swap(ia, ib);
  }
  *keys_sorted = ka; *items_sorted = ia;
  return LCREC_OK;
}

}  // namespace lcrec

using namespace lcrec;

extern "C" int64_t lcrec_collisions_workspace_bytes(int64_t n) {
  const int64_t tiles = ceil_div(std::max<int64_t>(n, 1), kTile);
  const int64_t sblocks = ceil_div(std::max<int64_t>(n, 1), kScanTile);
  return 2 * arena_need(8 * n) + 2 * arena_need(4 * n) + arena_need(4 * 256 * tiles) + arena_need(8 * n) +
         arena_need(8 * (sblocks + 1)) + arena_need(64) + 1024;
}

extern "C" int lcrec_sort_codes(const int64_t* codes, int64_t n, int n_levels, const int32_t* n_codes,
                                uint64_t* keys_out, uint32_t* items_out, void* ws, int64_t ws_bytes, void* stream) {
  LC_ARG(n >= 0 && n_levels >= 1 && n_levels <= LCREC_MAX_LEVELS && n_codes);
  LC_TRY(lcrec_device_check());
  if (n == 0) return LCREC_OK;
  LC_ARG(codes && keys_out && items_out);
  cudaStream_t st = (cudaStream_t)stream;
  SortPlan plan;
  LC_TRY(make_plan(n, n_levels, n_codes, &plan));
  Arena ar(ws, ws_bytes);
  uint64_t* k0 = ar.take<uint64_t>(n); uint64_t* k1 = ar.take<uint64_t>(n);
  uint32_t* i0 = ar.take<uint32_t>(n); uint32_t* i1 = ar.take<uint32_t>(n);
  uint32_t* hist = ar.take<uint32_t>((int64_t)256 * plan.n_tiles);
  if (!ar.ok()) { set_error("sort_codes: workspace too small"); return LCREC_ERR_NOMEM; }
  uint64_t* ks; uint32_t* is;
  LC_TRY(sort_impl(codes, n, plan, k0, i0, k1, i1, hist, st, &ks, &is));
  LC_CUDA(cudaMemcpyAsync(keys_out, ks, 8 * n, cudaMemcpyDeviceToDevice, st));
  LC_CUDA(cudaMemcpyAsync(items_out, is, 4 * n, cudaMemcpyDeviceToDevice, st));
  return LCREC_OK;
}

static int collisions_impl(const int64_t* codes, int64_t n, int n_levels, const int32_t* n_codes, int drop_levels,
                           int64_t* offsets, int64_t* members, int64_t* counts, void* ws, int64_t ws_bytes, void* stream);

extern "C" int lcrec_collisions(const int64_t* codes, int64_t n, int n_levels, const int32_t* n_codes, int64_t* offsets,
                                int64_t* members, int64_t* counts, void* ws, int64_t ws_bytes, void* stream) {
  return collisions_impl(codes, n, n_levels, n_codes, 0, offsets, members, counts, ws, ws_bytes, stream);
}

// Items that share their first n_levels - 1 codes (runs of >= 2): the only items that can ever collide while the
// rounds rewrite the last level only.  Same CSR and counts layout as lcrec_collisions; members of a segment are
// ordered by (last code, item).
extern "C" int lcrec_prefix_segments(const int64_t* codes, int64_t n, int n_levels, const int32_t* n_codes, int64_t* seg_offsets,
                                     int64_t* seg_members, int64_t* counts, void* ws, int64_t ws_bytes, void* stream) {
  LC_ARG(n_levels >= 2);
  return collisions_impl(codes, n, n_levels, n_codes, 1, seg_offsets, seg_members, counts, ws, ws_bytes, stream);
}

static int collisions_impl(const int64_t* codes, int64_t n, int n_levels, const int32_t* n_codes, int drop_levels,
                           int64_t* offsets, int64_t* members, int64_t* counts, void* ws, int64_t ws_bytes, void* stream) {
  LC_ARG(n >= 0 && n_levels >= 1 && n_levels <= LCREC_MAX_LEVELS && n_codes && counts);
  LC_TRY(lcrec_device_check());
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) { LC_CUDA(cudaMemsetAsync(counts, 0, 4 * sizeof(int64_t), st)); return LCREC_OK; }
  LC_ARG(codes && offsets && members);
  SortPlan plan;
  LC_TRY(make_plan(n, n_levels, n_codes, &plan));
  Arena ar(ws, ws_bytes);
  uint64_t* k0 = ar.take<uint64_t>(n); uint64_t* k1 = ar.take<uint64_t>(n);
  uint32_t* i0 = ar.take<uint32_t>(n); uint32_t* i1 = ar.take<uint32_t>(n);
  uint32_t* hist = ar.take<uint32_t>((int64_t)256 * plan.n_tiles);
  uint64_t* flags = ar.take<uint64_t>(n);
  const int64_t sblocks = ceil_div(n, kScanTile);
  uint64_t* sums = ar.take<uint64_t>(sblocks + 1);
  uint64_t* total = ar.take<uint64_t>(8);
  if (!ar.ok()) { set_error("collisions: workspace too small (%lld given, %lld needed)", (long long)ws_bytes, (long long)lcrec_collisions_workspace_bytes(n)); return LCREC_ERR_NOMEM; }
  uint64_t* ks; uint32_t* is;
  LC_TRY(sort_impl(codes, n, plan, k0, i0, k1, i1, hist, st, &ks, &is));
  const int drop_bits = drop_levels > 0 ? plan.pa.shift[n_levels - 1 - drop_levels] : 0;   // level L-1 sits in the low bits
  const int blocks = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)num_sms() * 8);
  run_flags_kernel<<<blocks, 256, 0, st>>>(ks, n, flags, drop_bits);
  LC_LAUNCH_CHECK("run_flags_kernel");
  scan_reduce_kernel<<<(unsigned)sblocks, 256, 0, st>>>(flags, n, sums);
  LC_LAUNCH_CHECK("scan_reduce_kernel");
  scan_sums_kernel<<<1, 1024, 0, st>>>(sums, sblocks, total);
  LC_LAUNCH_CHECK("scan_sums_kernel");
  emit_groups_kernel<<<(unsigned)sblocks, 256, 0, st>>>(flags, is, n, sums, offsets, members);
  LC_LAUNCH_CHECK("emit_groups_kernel");
  finish_counts_kernel<<<1, 1, 0, st>>>(total, n, offsets, counts);
  LC_LAUNCH_CHECK("finish_counts_kernel");
  max_mult_kernel<<<std::max(1, std::min(blocks, 256)), 256, 0, st>>>(offsets, counts, counts);
  LC_LAUNCH_CHECK("max_mult_kernel");
  return LCREC_OK;
}

extern "C" int64_t lcrec_segment_collisions_workspace_bytes(int64_t max_segments) {
  return arena_need(64) + arena_need(4 * (max_segments + 1)) + 1024;
}

// Collision groups of the current codes[:, level] inside the prefix segments (see above).  counts (device, 8 int64):
// [n_unique, n_groups, rows, max_multiplicity, -, fallback]; fallback = 1 means a segment was too large for the
// on-chip sort and the result is incomplete: call lcrec_collisions instead.
extern "C" int lcrec_collisions_in_segments_active(const int64_t* codes, int64_t n, int n_levels, int level,
                                                   const int64_t* seg_offsets, const int64_t* seg_members,
                                                   const int64_t* n_segs_dev, int64_t max_segments, const int32_t* active_in,
                                                   const int32_t* n_active_in, int32_t* active_out, int32_t* n_active_out,
                                                   int64_t* offsets, int64_t* members, int64_t* counts, void* ws,
                                                   int64_t ws_bytes, void* stream);

extern "C" int lcrec_collisions_in_segments(const int64_t* codes, int64_t n, int n_levels, int level, const int64_t* seg_offsets,
                                            const int64_t* seg_members, const int64_t* n_segs_dev, int64_t max_segments,
                                            int64_t* offsets, int64_t* members, int64_t* counts, void* ws, int64_t ws_bytes,
                                            void* stream) {
  return lcrec_collisions_in_segments_active(codes, n, n_levels, level, seg_offsets, seg_members, n_segs_dev, max_segments,
                                             nullptr, nullptr, nullptr, nullptr, offsets, members, counts, ws, ws_bytes, stream);
}

// Same with an explicit active list: active_in / n_active_in (device; NULL = examine every segment) name the segments
// to examine, active_out / n_active_out (device; NULL = not wanted; *n_active_out must be 0 on entry) receive the
// segments that still contain a collision - the input of the next round.
extern "C" int lcrec_collisions_in_segments_active(const int64_t* codes, int64_t n, int n_levels, int level,
                                                   const int64_t* seg_offsets, const int64_t* seg_members,
                                                   const int64_t* n_segs_dev, int64_t max_segments, const int32_t* active_in,
                                                   const int32_t* n_active_in, int32_t* active_out, int32_t* n_active_out,
                                                   int64_t* offsets, int64_t* members, int64_t* counts, void* ws,
                                                   int64_t ws_bytes, void* stream) {
  LC_ARG(n >= 0 && n_levels >= 1 && level >= 0 && level < n_levels && max_segments >= 0 && counts);
  LC_ARG((active_in == nullptr) == (n_active_in == nullptr) && (active_out == nullptr) == (n_active_out == nullptr));
  LC_TRY(lcrec_device_check());
  cudaStream_t st = (cudaStream_t)stream;
  LC_ARG(codes && seg_offsets && seg_members && n_segs_dev && offsets && members);
  Arena ar(ws, ws_bytes);
  unsigned long long* ctl = ar.take<unsigned long long>(8);
  int32_t* big_list = ar.take<int32_t>(max_segments + 1);
  if (!ar.ok()) { set_error("collisions_in_segments: workspace too small"); return LCREC_ERR_NOMEM; }
  LC_CUDA(cudaMemsetAsync(ctl, 0, 64, st));
  if (max_segments > 0) {
    const int64_t blocks = std::max<int64_t>(1, std::min<int64_t>(ceil_div(max_segments, 8), (int64_t)num_sms() * 8));
    segment_collisions_kernel<<<(unsigned)blocks, 256, 0, st>>>(codes, n_levels, level, seg_offsets, seg_members, n_segs_dev,
                                                                 offsets, members, ctl, big_list, active_in, n_active_in,
                                                                 active_out, n_active_out);
    LC_LAUNCH_CHECK("segment_collisions_kernel");
    segment_collisions_big_kernel<<<(unsigned)std::min<int64_t>(max_segments, num_sms()), 256, 0, st>>>(
        codes, n_levels, level, seg_offsets, seg_members, offsets, members, ctl, big_list);
    LC_LAUNCH_CHECK("segment_collisions_big_kernel");
  }
  segment_finish_kernel<<<1, 1, 0, st>>>(ctl, n, offsets, counts);
  LC_LAUNCH_CHECK("segment_finish_kernel");
  return LCREC_OK;
}
