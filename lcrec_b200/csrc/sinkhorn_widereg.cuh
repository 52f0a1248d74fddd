// Register-resident variant of the large-codebook cluster kernel (sinkhorn_wide.cuh) - included by sinkhorn.cu after it.
//
// sinkhorn_wide_kernel keeps E = exp(-dc / eps) in shared memory and is bound by the shared-memory reads of the same warps that
// issue the fp64 work (8 doubles x n rows per thread per pass).  Here every thread holds its share of E in REGISTERS:
// 512 threads per CTA, RM rows x CP columns = 48 doubles per thread, with (C, RM, CP) = (1, 3, 16) and (2, 6, 8) for 8192 codes
// and (1, 6, 8) for 4096 codes - the classes that hold ~90 % of the groups.  Shared memory only carries the row-sum reduction
// (warp tree -> 16 warp partials -> (C = 2) one DSMEM exchange per step).  Same arithmetic as the shared-memory kernel: first row
// step with v = 1, then fused [column step; row partials] passes, the reference's last column step evaluated literally, the
// certainty filter; v is not stored (the last pass's v is recomputed from the previous u, bit-identically).
#pragma once

namespace lcrec {

constexpr int kWrThreads = 512;

template <int C, int RM, int CP>
__global__ void __launch_bounds__(kWrThreads, 1) sinkhorn_widereg_kernel(const SkWideArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int NW = kWrThreads / 32;
  __shared__ double red[NW][RM];
  __shared__ double xch[2][RM];
  __shared__ double u_s[2][RM];
  __shared__ double best_v[RM], row_v[RM];
  __shared__ int red_k[NW][RM];
  __shared__ int best_k[RM], row_k[RM];
  __shared__ float fred[2][NW];
  __shared__ float mm[4];
  __shared__ int misc[1];
  const int K = a.K, Kc = K / C;
  if (Kc != CP * kWrThreads) return;                    // the host dispatches this kernel only for K = C x CP x 512: every column slot is live
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned rank = C > 1 ? cluster.block_rank() : 0;
  const int n_work = *a.count;
  const double Kd = (double)K, invK = 1.0 / Kd;
  const bool kpow2 = (K & (K - 1)) == 0;
  const unsigned n_clusters = gridDim.x / C, cluster_id = blockIdx.x / C;
  bool bad = false;
  unsigned parity = 0;
  for (int w = (int)cluster_id; w < n_work; w += (int)n_clusters) {
    const int64_t g = a.list[w];
    const int64_t beg = a.offsets[g];
    const int n = (int)(a.offsets[g + 1] - beg);
    if (n < a.rows_lo || n > a.rows_hi || n > RM) continue;
    const double Bd = (double)n;
    double E[RM][CP];
    float lmax = -INFINITY, lmin = INFINITY;
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int c = 0; c < CP; ++c) {
        const int kl = tid + kWrThreads * c;
        E[i][c] = 0.0;
        if (i < n) {
          const float d = a.dist[(beg + i) * K + rank * Kc + kl];
          E[i][c] = (double)d;
          lmax = fmaxf(lmax, d); lmin = fminf(lmin, d);
        }
      }
    lmax = warp_max(lmax); lmin = warp_min(lmin);
    if (lane == 0) { fred[0][warp] = lmax; fred[1][warp] = lmin; }
    __syncthreads();
    if (tid == 0) {
      float mx = fred[0][0], mn = fred[1][0];
      for (int q = 1; q < NW; ++q) { mx = fmaxf(mx, fred[0][q]); mn = fminf(mn, fred[1][q]); }
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
    if constexpr (C > 1) cluster.sync();                 // peers have read mm[0..1]
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int c = 0; c < CP; ++c) {
        const int kl = tid + kWrThreads * c;
        if (i < n) {
          const float dc = ((float)E[i][c] - mid) / amp;                    // fp32 centring, vq.py:60
          E[i][c] = exp(-((double)dc / a.eps));                             // layers.py:87
        }
      }
    // row totals over all CTAs of the cluster -> u_s[slot][i] = 1 / (B total_i)
    auto reduce_rows = [&](const double (&part)[RM], int slot) {
#pragma unroll
      for (int i = 0; i < RM; ++i)
        if (i < n) {
          const double rs = warp_sum(part[i]);
          if (lane == 0) red[warp][i] = rs;
        }
      __syncthreads();
      if (warp < n) {
        double rs = lane < NW ? red[lane][warp] : 0.0;
        rs = warp_sum(rs);
        if (lane == 0) {
          if constexpr (C > 1) xch[parity & 1][warp] = rs;
          else u_s[slot][warp] = fast_rcp(Bd * rs);
        }
      } else if (warp < RM && lane == 0) {
        u_s[slot][warp] = 0.0;                           // rows beyond the group: E = 0 and u = 0, the hot loops run unpredicated
      }
      if constexpr (C > 1) {
        cluster.sync();
        if (tid < n) {
          double rs = 0.0;
          for (unsigned r = 0; r < (unsigned)C; ++r) rs += cluster.map_shared_rank(&xch[0][0], r)[(parity & 1) * RM + tid];     // rank order
          u_s[slot][tid] = fast_rcp(Bd * rs);
        } else if (tid < RM) {
          u_s[slot][tid] = 0.0;
        }
        ++parity;
      }
      __syncthreads();
    };
    double part[RM];
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      double rs = 0.0;
#pragma unroll
      for (int c = 0; c < CP; ++c) rs += E[i][c];        // E = 0 outside the group / the CTA's columns
      part[i] = rs;
    }
    int cur = 0;
    reduce_rows(part, cur);
    for (int it = 1; it < a.iters; ++it) {               // fused [column step with u_cur; row partials with the new v]
      constexpr int UR = RM <= 6 ? RM : 1;               // u in registers for the small classes, broadcast shared-memory reads beyond
      double u[UR];
#pragma unroll
      for (int i = 0; i < UR; ++i) u[i] = u_s[cur][i];
#pragma unroll
      for (int i = 0; i < RM; ++i) part[i] = 0.0;
      const double* us = u_s[cur];
#pragma unroll
      for (int c = 0; c < CP; ++c) {                     // CP independent chains, no predicates: rows beyond n carry E = 0, u = 0
        double cs = 0.0;
#pragma unroll
        for (int i = 0; i < RM; ++i) cs = fma(RM <= 6 ? u[i < UR ? i : 0] : us[i], E[i][c], cs);
        const double vc = fast_rcp(Kd * cs);
#pragma unroll
        for (int i = 0; i < RM; ++i) part[i] = fma(E[i][c], vc, part[i]);
      }
      reduce_rows(part, cur ^ 1);                        // (its first barrier comes after every thread's reads of u_s[cur])
      cur ^= 1;
    }
    // ---- literal last column step on the materialised plan, * B; v of the last pass recomputed from the previous u
    const double scale = Kd / Bd;
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      {
        double vc = 1.0;
        if (a.iters > 1) {
          double cs0 = 0.0;
#pragma unroll
          for (int i = 0; i < RM; ++i) if (i < n) cs0 = fma(u_s[cur ^ 1][i], E[i][c], cs0);
          vc = fast_rcp(Kd * cs0);
        }
        double cs = 0.0;
#pragma unroll
        for (int i = 0; i < RM; ++i) if (i < n) cs = __dadd_rn(cs, __dmul_rn(__dmul_rn(u_s[cur][i], E[i][c]), vc));
#pragma unroll
        for (int i = 0; i < RM; ++i)
          if (i < n) {
            const double q = __dmul_rn(__dmul_rn(u_s[cur][i], E[i][c]), vc);
            const double val = __dmul_rn(kpow2 ? __dmul_rn(__ddiv_rn(q, cs), invK) : __ddiv_rn(__ddiv_rn(q, cs), Kd), Bd);   // / K is an exact multiply for K = 2^m
            bad = bad || isnan(val) || isinf(val);
            E[i][c] = val;
          }
      }
    }
    // ---- argmax per row (torch.argmax order)
#pragma unroll
    for (int i = 0; i < RM; ++i)
      if (i < n) {
        double bv = 0.0; int bk = 0x7fffffff;
#pragma unroll
        for (int c = 0; c < CP; ++c) {
          const int k = (int)rank * Kc + tid + kWrThreads * c;
          if (bk == 0x7fffffff || arg_better(E[i][c], k, bv, bk)) { bv = E[i][c]; bk = k; }
        }
        for (int o = 16; o > 0; o >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, bv, o);
          const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
          if (ok != 0x7fffffff && (bk == 0x7fffffff || arg_better(ob, ok, bv, bk))) { bv = ob; bk = ok; }
        }
        if (lane == 0) { red[warp][i] = bv; red_k[warp][i] = bk; }
      }
    __syncthreads();
    if (warp < n) {
      double bv = lane < NW ? red[lane][warp] : 0.0; int bk = lane < NW ? red_k[lane][warp] : 0x7fffffff;
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
      row_v[tid] = bv; row_k[tid] = bk;
      if (rank == 0) a.codes[a.members[beg + tid] * a.n_levels + a.level] = bk;
    }
    __syncthreads();
    if (a.risky_list != nullptr) {                       // certainty filter of sinkhorn.cu
      bool risky = false;
#pragma unroll
      for (int i = 0; i < RM; ++i)
        if (i < n) {
          const double best = row_v[i];
          const int bk = row_k[i];
          const double rowdev = fmax(0.0, 1.0 - best * scale) + 0x1p-50;
          if (!(best == best)) risky = true;
#pragma unroll
          for (int c = 0; c < CP; ++c) {
            if ((int)rank * Kc + tid + kWrThreads * c != bk) {
              const double val = E[i][c];
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
    if constexpr (C > 1) cluster.sync(); else __syncthreads();
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(a.flags, 1);
}

}  // namespace lcrec
