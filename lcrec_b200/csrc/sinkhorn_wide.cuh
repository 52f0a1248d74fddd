// Per-group Sinkhorn for LARGE codebooks (BASELINE configs[4]: 8192 codes, e_dim 256) - included by sinkhorn.cu.
//
// The CTA kernel of sinkhorn.cu serves a group with 256 threads and one WARP per row: at K = 8192 a 3-row group keeps 3 warps
// busy, its n x 8192 fp64 plan (64 KB per row) lives in an L2 slice, and every group streams the whole 8 MB codebook through
// shared memory for its distances.  Round 1 of a 100 000-item run: 134 ms for 51 678 rows (bench.py --config c5, round 2 start).
// Two kernels replace it for groups of up to 24 rows:
//
//  wide_distances_kernel   fp32 distances of ALL colliding rows in one pass, (rows x K) into the workspace: a register-tiled
//                          SIMT kernel (128 x 128 tile, 8 x 8 outputs per thread) whose every output is ONE fma chain in
//                          ascending dimension - bit-identical to the per-group chains of sinkhorn.cu / rq_fused.cu (a
//                          tensor-core GEMM would be ~10x faster but changes the last ulps, and a Sinkhorn pick can flip on one
//                          ulp: DESIGN.md 2.1).  The codebook is read once per 128 rows instead of once per group.
//  sinkhorn_wide_kernel<C> one thread-block CLUSTER of C in {1, 2, 4, 8} CTAs x 1024 threads per group: CTA r owns K / C
//                          columns, every thread <= 8 of them; E = exp(-dc / eps) for the CTA's columns sits in its shared
//                          memory (192 KB / (K / C x 8 B) rows: 3, 6, 12, 24), v in registers.  Row step: thread-local fma
//                          chain -> warp shuffle tree -> shared memory -> (C > 1) one exchange of the n partial row sums
//                          through distributed shared memory + cluster barrier.  Column step: local.  Scaling-vector form
//                          with the literal last column step and the certainty filter of sinkhorn.cu (risky groups are re-run
//                          by the literal kernel); groups of more than 24 rows stay on the CTA kernel.
#pragma once

namespace lcrec {

constexpr int kWdBM = 128, kWdBN = 128, kWdBK = 16, kWdThreads = 256;

// cc[k] = sum_d cb[k][d]^2 as an fma chain in ascending d (same chain as everywhere else)
__global__ void wide_sqnorm_kernel(const float* __restrict__ cb, int K, int D, float* __restrict__ cc) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float s = 0.f;
  for (int d = 0; d < D; ++d) { const float v = cb[(size_t)k * D + d]; s = fmaf(v, v, s); }
  cc[k] = s;
}

// dist[r][k] = (xx_r + cc_k) - 2 dot(r, k) for the CSR rows r < offsets[n_groups] (row r = resid[members[r]]), D % 16 == 0,
// K % 128 == 0.  128 x 128 tile, 16 x 16 threads, 8 x 8 outputs per thread (64 FMAs per four 128-bit shared-memory loads).
__global__ void __launch_bounds__(kWdThreads) wide_distances_kernel(const float* __restrict__ resid, const int64_t* __restrict__ members,
                                                                   const int64_t* __restrict__ offsets, const int64_t* __restrict__ n_groups_dev,
                                                                   const float* __restrict__ cb, const float* __restrict__ cc, int K, int D,
                                                                   float* __restrict__ dist, int64_t rows_cap) {
  __shared__ __align__(16) float As[kWdBK][kWdBM + 4];
  __shared__ __align__(16) float Bs[kWdBK][kWdBN + 4];
  __shared__ float xx_s[kWdBM];
  const int64_t n_rows = min(offsets[*n_groups_dev], rows_cap);
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;            // 16 x 16 threads: rows ty * 8 .. + 7, columns tx * 8 .. + 7
  const int64_t row_tiles = (n_rows + kWdBM - 1) / kWdBM;
  const int col_tiles = K / kWdBN;
  for (int64_t t = blockIdx.x; t < row_tiles * col_tiles; t += gridDim.x) {
    const int64_t r0 = (t / col_tiles) * kWdBM;
    const int c0 = (int)(t % col_tiles) * kWdBN;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float xx = 0.f;                                  // threads 0..127: squared norm of row r0 + tid (chain in ascending d)
    const int64_t my_row = r0 + tid;
    const float* xrow = (tid < kWdBM && my_row < n_rows) ? resid + members[my_row] * D : nullptr;
    // loaders: 128 rows x 16 dims = 512 float4 per operand, two per thread: row = idx / 4, dims (idx % 4) * 4
    const float* a_src[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t a_row = r0 + ((tid + h * kWdThreads) >> 2);
      a_src[h] = a_row < n_rows ? resid + members[a_row] * D : nullptr;
    }
    for (int d0 = 0; d0 < D; d0 += kWdBK) {
      __syncthreads();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int idx = tid + h * kWdThreads;
        const int rr = idx >> 2, dd = (idx & 3) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a_src[h]) v = *reinterpret_cast<const float4*>(a_src[h] + d0 + dd);
        As[dd + 0][rr] = v.x; As[dd + 1][rr] = v.y; As[dd + 2][rr] = v.z; As[dd + 3][rr] = v.w;
        const float4 w = *reinterpret_cast<const float4*>(cb + (size_t)(c0 + rr) * D + d0 + dd);
        Bs[dd + 0][rr] = w.x; Bs[dd + 1][rr] = w.y; Bs[dd + 2][rr] = w.z; Bs[dd + 3][rr] = w.w;
      }
      if (xrow)
#pragma unroll
        for (int dd = 0; dd < kWdBK; ++dd) { const float v = xrow[d0 + dd]; xx = fmaf(v, v, xx); }
      __syncthreads();
#pragma unroll
      for (int dd = 0; dd < kWdBK; ++dd) {             // ascending d: every acc[i][j] is one sequential fma chain
        const float4 a0 = *reinterpret_cast<const float4*>(&As[dd][ty * 8]), a1 = *reinterpret_cast<const float4*>(&As[dd][ty * 8 + 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[dd][tx * 8]), b1 = *reinterpret_cast<const float4*>(&Bs[dd][tx * 8 + 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
    __syncthreads();
    if (tid < kWdBM) xx_s[tid] = xx;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t r = r0 + ty * 8 + i;
      if (r >= n_rows) continue;
      const float x2 = xx_s[ty * 8 + i];
      float out[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) out[j] = (x2 + cc[c0 + tx * 8 + j]) - 2.f * acc[i][j];     // vq.py:71-73 evaluation order
      float4* dst = reinterpret_cast<float4*>(dist + r * K + c0 + tx * 8);
      dst[0] = make_float4(out[0], out[1], out[2], out[3]);
      dst[1] = make_float4(out[4], out[5], out[6], out[7]);
    }
  }
}

// ---- size classes of the wide path: <= 3, <= 6, <= 12, <= 24 rows (cluster of 1, 2, 4, 8 CTAs), larger -> CTA kernel
constexpr int kWideClasses = 5;
struct WideCaps { int rows[4]; };     // largest group served by a cluster of 1, 2, 4, 8 CTAs (0: class not available)
__global__ void __launch_bounds__(256) classify_wide_kernel(const int64_t* __restrict__ offsets, const int64_t* __restrict__ n_groups_dev,
                                                           int part_mod, int part_rem, WideCaps caps, int64_t rows_cap,
                                                           int32_t* __restrict__ lists, int64_t list_stride, int* __restrict__ counts) {
  const int64_t n_groups = *n_groups_dev;
  const int lane = threadIdx.x & 31;
  for (int64_t g0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) - lane; g0 < n_groups; g0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = g0 + lane;
    int cls = -1;
    if (g < n_groups && (part_mod <= 1 || (int)(g % part_mod) == part_rem)) {
      const int64_t n = offsets[g + 1] - offsets[g];
      if (n >= 2) {
        cls = n <= caps.rows[0] ? 0 : (n <= caps.rows[1] ? 1 : (n <= caps.rows[2] ? 2 : (n <= caps.rows[3] ? 3 : 4)));
        if (offsets[g + 1] > rows_cap) cls = 4;        // its distances are not in the (bounded) distance buffer
      }
    }
#pragma unroll
    for (int c = 0; c < kWideClasses; ++c) {
      const unsigned m = __ballot_sync(0xffffffffu, cls == c);
      if (m == 0) continue;
      int base = 0;
      if (lane == __ffs(m) - 1) base = atomicAdd(counts + c, __popc(m));
      base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
      if (cls == c) lists[(int64_t)c * list_stride + base + __popc(m & ((1u << lane) - 1u))] = (int32_t)g;
    }
  }
}

constexpr int kWideThreads = 1024;
constexpr int kWideCpt = 8;              // columns per thread at most: K / C <= 8192
constexpr int kWideMaxRows = 24;

struct SkWideArgs {
  const float* dist; int K; const int64_t* offsets; const int64_t* members;
  const int32_t* list; const int* count;
  double eps; int iters; int64_t* codes; int n_levels; int level; int32_t* flags;
  int32_t* risky_list; int* risky_count;       // null: no certainty filter (mode 1)
  int rows_cap_cta;                            // rows of E that fit this CTA's shared memory
  int rows_lo, rows_hi;                        // groups outside [rows_lo, rows_hi] are skipped
};

template <int C, bool LIT>
__global__ void __launch_bounds__(kWideThreads, 1) sinkhorn_wide_kernel(const SkWideArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char wd_smem[];
  const int K = a.K, Kc = K / C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kWideThreads / 32;
  const unsigned rank = C > 1 ? cluster.block_rank() : 0;
  double* E = reinterpret_cast<double*>(wd_smem);                     // rows_cap_cta x Kc
  double* red = E + (size_t)a.rows_cap_cta * Kc;                      // NW x kWideMaxRows (warp partials / candidates)
  double* xch = red + NW * kWideMaxRows;                              // 2 x kWideMaxRows: this CTA's row partials (double-buffered)
  double* u_s = xch + 2 * kWideMaxRows;                               // kWideMaxRows
  double* best_v = u_s + kWideMaxRows;                                // kWideMaxRows: this CTA's best value per row
  double* row_v = best_v + kWideMaxRows;                              // kWideMaxRows: the group's winning value per row
  int* red_k = reinterpret_cast<int*>(row_v + kWideMaxRows);          // NW x kWideMaxRows
  int* best_k = red_k + NW * kWideMaxRows;                            // kWideMaxRows
  int* row_k = best_k + kWideMaxRows;                                 // kWideMaxRows
  float* mm = reinterpret_cast<float*>(row_k + kWideMaxRows);         // [0] max, [1] min of this CTA; [2] mid, [3] amp
  int* misc = reinterpret_cast<int*>(mm + 4);                         // [0] risky of this CTA
  const int ncols = (Kc + kWideThreads - 1) / kWideThreads;           // columns this thread owns: tid + 1024 c (c < ncols) if < Kc
  const int n_work = *a.count;
  const double Kd = (double)K, invK = 1.0 / Kd;
  const bool kpow2 = (K & (K - 1)) == 0;
  const unsigned n_clusters = gridDim.x / C, cluster_id = blockIdx.x / C;
  bool bad = false;
  unsigned parity = 0;                                                // exchange-buffer parity, advances with every exchange
  for (int w = (int)cluster_id; w < n_work; w += (int)n_clusters) {
    const int64_t g = a.list[w];
    const int64_t beg = a.offsets[g];
    const int n = (int)(a.offsets[g + 1] - beg);
    if (n < a.rows_lo || n > a.rows_hi) continue;        // (a list that mixes sizes, e.g. the literal re-run of risky groups)
    const double Bd = (double)n;
    // ---- distances of this CTA's columns, max / min over the whole group (vq.py:54-55)
    float lmax = -INFINITY, lmin = INFINITY;
    for (int i = 0; i < n; ++i)
#pragma unroll
      for (int c = 0; c < kWideCpt; ++c) {
        const int kl = tid + kWideThreads * c;
        if (c < ncols && kl < Kc) {
          const float d = a.dist[(beg + i) * K + rank * Kc + kl];
          E[(size_t)i * Kc + kl] = (double)d;
          lmax = fmaxf(lmax, d); lmin = fminf(lmin, d);
        }
      }
    lmax = warp_max(lmax); lmin = warp_min(lmin);
    float* fred = reinterpret_cast<float*>(red);
    if (lane == 0) { fred[warp] = lmax; fred[NW + warp] = lmin; }
    __syncthreads();
    if (tid == 0) {
      float mx = fred[0], mn = fred[NW];
      for (int q = 1; q < NW; ++q) { mx = fmaxf(mx, fred[q]); mn = fminf(mn, fred[NW + q]); }
      mm[0] = mx; mm[1] = mn;
    }
    if constexpr (C > 1) cluster.sync(); else __syncthreads();
    if (tid == 0) {
      float mx = mm[0], mn = mm[1];
      if constexpr (C > 1)
        for (unsigned r = 0; r < (unsigned)C; ++r) { const float* o = cluster.map_shared_rank(mm, r); mx = fmaxf(mx, o[0]); mn = fminf(mn, o[1]); }
      const float mid = (mx + mn) / 2.f;                 // vq.py:57
      const float amp = (mx - mid) + 1e-5f;              // vq.py:58
      mm[2] = mid; mm[3] = amp;
      if (!(amp > 0.f) && rank == 0) atomicOr(a.flags, 4);   // vq.py:59
    }
    __syncthreads();
    const float mid = mm[2], amp = mm[3];
    if constexpr (C > 1) cluster.sync();                 // peers have read mm[0..1] before a later group overwrites them
    // ---- E = exp(-dc / eps) (layers.py:87), fp32 centring (vq.py:60)
    for (int i = 0; i < n; ++i)
#pragma unroll
      for (int c = 0; c < kWideCpt; ++c) {
        const int kl = tid + kWideThreads * c;
        if (c < ncols && kl < Kc) {
          const float dc = ((float)E[(size_t)i * Kc + kl] - mid) / amp;
          E[(size_t)i * Kc + kl] = exp(-((double)dc / a.eps));
        }
      }
    // Row totals of all CTAs: thread partials part[i] (i < n) -> warp tree -> shared memory -> (C > 1) DSMEM exchange.
    // RCP: u_s[i] = 1 / (B total_i) (scaling form); otherwise u_s[i] = total_i (literal form).
    auto reduce_rows = [&](const double* part, bool rcp) {
      for (int i = 0; i < n; ++i) {
        const double rs = warp_sum(part[i]);
        if (lane == 0) red[warp * kWideMaxRows + i] = rs;
      }
      __syncthreads();
      if (warp < n) {                                    // warp i adds the 32 warp partials of row i
        double rs = red[lane * kWideMaxRows + warp];
        rs = warp_sum(rs);
        if (lane == 0) {
          if constexpr (C > 1) xch[(parity & 1) * kWideMaxRows + warp] = rs;
          else u_s[warp] = rcp ? fast_rcp(Bd * rs) : rs;
        }
      }
      if constexpr (C > 1) {
        cluster.sync();                                  // every CTA's row partials of this step are visible
        if (tid < n) {
          double rs = 0.0;
          for (unsigned r = 0; r < (unsigned)C; ++r) rs += cluster.map_shared_rank(xch, r)[(parity & 1) * kWideMaxRows + tid];   // rank order
          u_s[tid] = rcp ? fast_rcp(Bd * rs) : rs;
        }
        ++parity;
      }
      __syncthreads();
    };
    double part[kWideMaxRows];
    if constexpr (LIT) {
      // ---- the reference's literal in-place divide sequence (layers.py:93-107) on this cluster's copy of the matrix
      double tot = 0.0;
      for (int i = 0; i < n; ++i)
#pragma unroll
        for (int c = 0; c < kWideCpt; ++c) {
          const int kl = tid + kWideThreads * c;
          if (c < ncols && kl < Kc) tot += E[(size_t)i * Kc + kl];
        }
      part[0] = tot;
      {                                                  // grand total through the same reduction (one "row")
        const double rs = warp_sum(part[0]);
        if (lane == 0) red[warp * kWideMaxRows] = rs;
        __syncthreads();
        if (warp == 0) {
          double t = warp_sum(red[lane * kWideMaxRows]);
          if (lane == 0) { if constexpr (C > 1) xch[(parity & 1) * kWideMaxRows] = t; else u_s[0] = t; }
        }
        if constexpr (C > 1) {
          cluster.sync();
          if (tid == 0) {
            double t = 0.0;
            for (unsigned r = 0; r < (unsigned)C; ++r) t += cluster.map_shared_rank(xch, r)[(parity & 1) * kWideMaxRows];
            u_s[0] = t;
          }
          ++parity;
        }
        __syncthreads();
      }
      const double total = u_s[0];
      __syncthreads();
      for (int i = 0; i < n; ++i)
#pragma unroll
        for (int c = 0; c < kWideCpt; ++c) {
          const int kl = tid + kWideThreads * c;
          if (c < ncols && kl < Kc) E[(size_t)i * Kc + kl] /= total;                        // layers.py:94
        }
      for (int it = 0; it < a.iters; ++it) {
        for (int i = 0; i < n; ++i) {
          double rs = 0.0;
#pragma unroll
          for (int c = 0; c < kWideCpt; ++c) {
            const int kl = tid + kWideThreads * c;
            if (c < ncols && kl < Kc) rs += E[(size_t)i * Kc + kl];
          }
          part[i] = rs;
        }
        reduce_rows(part, false);
#pragma unroll
        for (int c = 0; c < kWideCpt; ++c) {
          const int kl = tid + kWideThreads * c;
          if (c < ncols && kl < Kc) {
            double cs = 0.0;
            for (int i = 0; i < n; ++i) {                                                    // rows: /= rowsum, /= B (layers.py:99-100)
              const double q = (E[(size_t)i * Kc + kl] / u_s[i]) / Bd;
              E[(size_t)i * Kc + kl] = q;
              cs += q;
            }
            for (int i = 0; i < n; ++i) E[(size_t)i * Kc + kl] = (E[(size_t)i * Kc + kl] / cs) / Kd;      // columns (layers.py:103-104)
          }
        }
        __syncthreads();
      }
      for (int i = 0; i < n; ++i)
#pragma unroll
        for (int c = 0; c < kWideCpt; ++c) {
          const int kl = tid + kWideThreads * c;
          if (c < ncols && kl < Kc) {
            const double val = E[(size_t)i * Kc + kl] * Bd;                                  // layers.py:107
            bad = bad || isnan(val) || isinf(val);
            E[(size_t)i * Kc + kl] = val;
          }
        }
    } else {
      // ---- scaling-vector iterations.  R_1 with v = 1, then (iters - 1) FUSED passes [column step with u_k; row partials with
      // the new v_k]: one read of the thread's columns per iteration instead of two (the kernel is shared-memory-bandwidth
      // bound: 8 doubles x n rows per thread per read).  NRR rows of a column are held in registers inside a pass.
      constexpr int NRR = C == 1 ? 3 : (C == 2 ? 6 : 0);
      double v[kWideCpt];
#pragma unroll
      for (int c = 0; c < kWideCpt; ++c) v[c] = 1.0;
      for (int i = 0; i < n; ++i) {
        double rs = 0.0;
#pragma unroll
        for (int c = 0; c < kWideCpt; ++c) {
          const int kl = tid + kWideThreads * c;
          if (c < ncols && kl < Kc) rs += E[(size_t)i * Kc + kl];                            // fma(E, 1.0, rs)
        }
        part[i] = rs;
      }
      reduce_rows(part, true);
      const bool rows_in_regs = NRR > 0 && a.rows_hi <= NRR;      // (smaller codebooks give a CTA more rows than NRR: two-read form)
      for (int it = 1; it < a.iters; ++it) {
        if (rows_in_regs) {
          constexpr int NR = NRR > 0 ? NRR : 1;
          double pr[NR];                                  // row partials of this pass, in registers
#pragma unroll
          for (int i = 0; i < NR; ++i) pr[i] = 0.0;
#pragma unroll
          for (int c = 0; c < kWideCpt; ++c) {
            const int kl = tid + kWideThreads * c;
            if (c < ncols && kl < Kc) {
              double e[NR];
#pragma unroll
              for (int i = 0; i < NR; ++i) e[i] = i < n ? E[(size_t)i * Kc + kl] : 0.0;
              double cs = 0.0;
#pragma unroll
              for (int i = 0; i < NR; ++i) if (i < n) cs = fma(u_s[i], e[i], cs);
              const double vc = fast_rcp(Kd * cs);
              v[c] = vc;
#pragma unroll
              for (int i = 0; i < NR; ++i) if (i < n) pr[i] = fma(e[i], vc, pr[i]);
            }
          }
#pragma unroll
          for (int i = 0; i < NR; ++i) if (i < n) part[i] = pr[i];
        } else {
#pragma unroll
          for (int c = 0; c < kWideCpt; ++c) {            // column step (local), v in registers
            const int kl = tid + kWideThreads * c;
            if (c < ncols && kl < Kc) {
              double cs = 0.0;
              for (int i = 0; i < n; ++i) cs = fma(u_s[i], E[(size_t)i * Kc + kl], cs);
              v[c] = fast_rcp(Kd * cs);
            }
          }
          for (int i = 0; i < n; ++i) {                   // row partials with the new v
            double rs = 0.0;
#pragma unroll
            for (int c = 0; c < kWideCpt; ++c) {
              const int kl = tid + kWideThreads * c;
              if (c < ncols && kl < Kc) rs = fma(E[(size_t)i * Kc + kl], v[c], rs);
            }
            part[i] = rs;
          }
        }
        __syncthreads();                                 // every thread has read u_s before reduce_rows overwrites it
        reduce_rows(part, true);
      }
      // ---- literal last column step on the materialised plan, * B (rounded products, plain adds, IEEE divisions)
#pragma unroll
      for (int c = 0; c < kWideCpt; ++c) {
        const int kl = tid + kWideThreads * c;
        if (c < ncols && kl < Kc) {
          double cs = 0.0;
          for (int i = 0; i < n; ++i) cs = __dadd_rn(cs, __dmul_rn(__dmul_rn(u_s[i], E[(size_t)i * Kc + kl]), v[c]));
          for (int i = 0; i < n; ++i) {
            const double q = __dmul_rn(__dmul_rn(u_s[i], E[(size_t)i * Kc + kl]), v[c]);
            const double val = __dmul_rn(kpow2 ? __dmul_rn(__ddiv_rn(q, cs), invK) : __ddiv_rn(__ddiv_rn(q, cs), Kd), Bd);   // / K is an exact multiply for K = 2^m
            bad = bad || isnan(val) || isinf(val);
            E[(size_t)i * Kc + kl] = val;
          }
        }
      }
    }
    const double scale = Kd / Bd;
    __syncthreads();
    // ---- argmax per row (vq.py:81-83): thread -> warp -> CTA -> cluster, torch.argmax order (NaN first, then value, then index)
    for (int i = 0; i < n; ++i) {
      double bv = 0.0; int bk = 0x7fffffff;
#pragma unroll
      for (int c = 0; c < kWideCpt; ++c) {
        const int kl = tid + kWideThreads * c;
        if (c < ncols && kl < Kc) {
          const double val = E[(size_t)i * Kc + kl];
          const int k = (int)rank * Kc + kl;
          if (bk == 0x7fffffff || arg_better(val, k, bv, bk)) { bv = val; bk = k; }
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, bv, o);
        const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
        if (ok != 0x7fffffff && (bk == 0x7fffffff || arg_better(ob, ok, bv, bk))) { bv = ob; bk = ok; }
      }
      if (lane == 0) { red[warp * kWideMaxRows + i] = bv; red_k[warp * kWideMaxRows + i] = bk; }
    }
    __syncthreads();
    if (warp < n) {
      double bv = red[lane * kWideMaxRows + warp]; int bk = red_k[lane * kWideMaxRows + warp];
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, bv, o);
        const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
        if (ok != 0x7fffffff && (bk == 0x7fffffff || arg_better(ob, ok, bv, bk))) { bv = ob; bk = ok; }
      }
      if (lane == 0) { best_v[warp] = bv; best_k[warp] = bk; }
    }
    if constexpr (C > 1) cluster.sync(); else __syncthreads();
    if (tid < n) {
      double bv = best_v[tid]; int bk = best_k[tid];
      if constexpr (C > 1)
        for (unsigned r = 0; r < (unsigned)C; ++r) {
          const double ob = cluster.map_shared_rank(best_v, r)[tid];
          const int ok = cluster.map_shared_rank(best_k, r)[tid];
          if (ok != 0x7fffffff && (bk == 0x7fffffff || arg_better(ob, ok, bv, bk))) { bv = ob; bk = ok; }
        }
      row_v[tid] = bv;                                   // the row's winning value
      row_k[tid] = bk;                                   // and its column
      if (rank == 0) a.codes[a.members[beg + tid] * a.n_levels + a.level] = bk;
    }
    __syncthreads();
    if (!LIT && a.risky_list != nullptr) {
      // certainty filter of sinkhorn.cu: is the argmax provably the one the literal kernel computes?
      bool risky = false;
      for (int i = 0; i < n; ++i) {
        const double best = row_v[i];
        const int bk = row_k[i];
        const double rowdev = fmax(0.0, 1.0 - best * scale) + 0x1p-50;
        if (!(best == best)) risky = true;
#pragma unroll
        for (int c = 0; c < kWideCpt; ++c) {
          const int kl = tid + kWideThreads * c;
          if (c < ncols && kl < Kc) {
            const int k = (int)rank * Kc + kl;
            if (k == bk) continue;
            const double val = E[(size_t)i * Kc + kl];
            const double dev = fmax(fmax(0.0, 1.0 - val * scale) + 0x1p-50, rowdev);
            if (dev > 0x1p-40 && val >= best - best * (0x1p-51 + 2e-11 * dev)) risky = true;
          }
        }
      }
      if (tid == 0) misc[0] = 0;
      __syncthreads();
      if (__any_sync(0xffffffffu, risky) && lane == 0) atomicOr(&misc[0], 1);
      if constexpr (C > 1) cluster.sync(); else __syncthreads();
      if (rank == 0 && tid == 0) {
        int r = misc[0];
        if constexpr (C > 1)
          for (unsigned q = 1; q < (unsigned)C; ++q) r |= cluster.map_shared_rank(misc, q)[0];
        if (r) a.risky_list[atomicAdd(a.risky_count, 1)] = (int32_t)g;
      }
    }
    if constexpr (C > 1) cluster.sync(); else __syncthreads();      // shared memory is reused by the next group
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(a.flags, 1);
}

}  // namespace lcrec
