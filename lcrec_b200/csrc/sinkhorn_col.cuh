// Per-group Sinkhorn with one thread per code at small codebooks (K <= 256 codes) - included by sinkhorn.cu.  Serves the groups of
// 9..32 rows (RM = 16 / 32) of every round and, with RM = 4 / 8, ALL groups of a late collision round that holds only a few
// hundred of them: a late round is bound by the latency of one group, and a group spread over K threads runs ~50 dependent
// instructions per iteration where the warp kernel (one warp per group, 8 columns per lane) runs 113.
//
// The CTA kernel (sinkhorn_groups_kernel) keeps the n x K kernel matrix in shared memory and every FMA of the 50 iterations
// pays an 8-byte shared-memory read: 6.5 % of the fp64 pipe (ncu, profiles/r2_sk256_*).  Here the CTA has one THREAD PER
// COLUMN (blockDim = K): a thread holds its column of E - RM = 16 or 32 doubles - in registers, v_k is a private register,
// the column step is thread-local, and only the row sums cross threads (transposed warp reduction -> K / 32 warp partials
// in shared memory -> u).  Same arithmetic as the warp kernels: fp32 distances by the same fma chains, scaling-vector form,
// the reference's last column step evaluated literally (IEEE quotient by the shared-reciprocal sequence), certainty filter.
#pragma once

namespace lcrec {

template <int RM> struct SkColShape { static constexpr int MINB = RM <= 16 ? 3 : 2; };

template <int RM, bool FILTER>
__global__ void __launch_bounds__(256, SkColShape<RM>::MINB) sinkhorn_groups_col_kernel(const SkGroupArgs a) {
  extern __shared__ __align__(16) unsigned char sk_smem[];
  constexpr int LOGRM = RM == 4 ? 2 : (RM == 8 ? 3 : (RM == 16 ? 4 : 5));
  static_assert(RM == 4 || RM == 8 || RM == 16 || RM == 32, "row classes of the column kernel");
  constexpr int kTail = 32 >> LOGRM;                     // lanes that still hold partials of the same row after the transposed stages
  const int T = blockDim.x, NW = T >> 5;                 // T == K
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = a.K, D = a.D;
  // shared memory: codebook transposed [d][k] | rows transposed [d][RM] | exp scratch [RM][T] doubles | reductions
  float* cb_s = reinterpret_cast<float*>(sk_smem);
  float* rows_t = cb_s + (size_t)D * K;
  double* scr = reinterpret_cast<double*>(rows_t + (size_t)D * RM);
  double* red = scr + (size_t)RM * T;                    // NW x RM
  double* u_s = red + (size_t)8 * RM;                    // RM
  double* row_v = u_s + RM;                              // RM
  int* red_k = reinterpret_cast<int*>(row_v + RM);       // NW x RM
  int* row_k = red_k + 8 * RM;                           // RM
  float* fred = reinterpret_cast<float*>(row_k + RM);    // 2 x 8 + 2
  __shared__ int s_base;
  const int n_work = *a.work_count;
  if ((int)blockIdx.x >= n_work) return;
  float cc = 0.f;
  {
    const float* src = a.cb + (size_t)tid * D;
    for (int d = 0; d < D; ++d) {
      const float v = __ldg(src + d);
      cb_s[d * K + tid] = v;
      cc = fmaf(v, v, cc);                               // same chain as the other kernels (d ascending)
    }
  }
  const double Kd = (double)K, invK = 1.0 / Kd;          // K is a power of two (host check): x / K == x * invK bit for bit
  bool bad = false;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_base = atomicAdd(a.work_cursor, 1);
    __syncthreads();
    const int w = s_base;
    if (w >= n_work) break;
    const int64_t g = a.work_list[w];
    const int64_t beg = a.offsets[g];
    const int n = (int)(a.offsets[g + 1] - beg);
    if (n > RM) continue;                                // (not reached: the classification bounds the list)
    for (int idx = tid; idx < n * D; idx += T) {
      const int i = idx / D, d = idx - i * D;
      rows_t[d * RM + i] = a.resid[a.members[beg + i] * D + d];
    }
    __syncthreads();
    // ---- fp32 distances of the thread's code to every row, four rows per pass over d
    float dist[RM];
    float lmax = -INFINITY, lmin = INFINITY;
#pragma unroll
    for (int i0 = 0; i0 < RM; i0 += 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) dist[i0 + j] = 0.f;
      if (i0 < n) {
        float dot[4] = {0.f, 0.f, 0.f, 0.f}, xx[4] = {0.f, 0.f, 0.f, 0.f};
        for (int d = 0; d < D; ++d) {
          const float c = cb_s[d * K + tid];
          const float4 r = *reinterpret_cast<const float4*>(rows_t + d * RM + i0);
          dot[0] = fmaf(r.x, c, dot[0]); xx[0] = fmaf(r.x, r.x, xx[0]);
          dot[1] = fmaf(r.y, c, dot[1]); xx[1] = fmaf(r.y, r.y, xx[1]);
          dot[2] = fmaf(r.z, c, dot[2]); xx[2] = fmaf(r.z, r.z, xx[2]);
          dot[3] = fmaf(r.w, c, dot[3]); xx[3] = fmaf(r.w, r.w, xx[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (i0 + j < n) {
            const float dd = (xx[j] + cc) - 2.f * dot[j];
            dist[i0 + j] = dd;
            lmax = fmaxf(lmax, dd); lmin = fminf(lmin, dd);
          }
      }
    }
    lmax = warp_max(lmax); lmin = warp_min(lmin);
    if (lane == 0) { fred[warp] = lmax; fred[8 + warp] = lmin; }
    __syncthreads();
    if (tid == 0) {
      float mx = fred[0], mn = fred[8];
      for (int q = 1; q < NW; ++q) { mx = fmaxf(mx, fred[q]); mn = fminf(mn, fred[8 + q]); }
      const float mid = (mx + mn) / 2.f;                 // vq.py:57
      const float amp = (mx - mid) + 1e-5f;              // vq.py:58
      fred[16] = mid; fred[17] = amp;
      if (!(amp > 0.f)) atomicOr(a.flags, 4);            // vq.py:59
    }
    __syncthreads();
    const float mid = fred[16], amp = fred[17];
    // ---- E = exp(-dc / eps) (layers.py:87): rolled loop over a thread-private shared-memory column (the inlined exp is ~100
    // instructions; RM copies of it would be most of the kernel)
    double* my = scr + tid;
#pragma unroll
    for (int i = 0; i < RM; ++i) my[(size_t)i * T] = (double)((dist[i] - mid) / amp);      // fp32 centring, vq.py:60
#pragma unroll 2
    for (int i = 0; i < n; ++i) my[(size_t)i * T] = exp(-(my[(size_t)i * T] / a.eps));
    double E[RM];
#pragma unroll
    for (int i = 0; i < RM; ++i) E[i] = i < n ? my[(size_t)i * T] * Kd : 0.0;      // K E: the column step needs no multiply
    const double Bd = (double)n, BdK = Bd * invK;
    // ---- iterations, scaling-vector form
    double v = 1.0;
    for (int it = 0; it < a.iters; ++it) {
      double cur[RM];
#pragma unroll
      for (int i = 0; i < RM; ++i) cur[i] = E[i] * v;    // E = 0 beyond n
      // transposed warp reduction: each exchange halves the values a lane carries; row i ends in the lanes whose top LOGRM bits spell i
#pragma unroll
      for (int st = 0; st < LOGRM; ++st) {
        const int o = 16 >> st;
        const int cnt = RM >> (st + 1);
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int j = 0; j < cnt; ++j) {
          const double keep = upper ? cur[j + cnt] : cur[j];
          const double send = upper ? cur[j] : cur[j + cnt];
          cur[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
#pragma unroll
      for (int o = kTail >> 1; o > 0; o >>= 1) cur[0] += __shfl_xor_sync(0xffffffffu, cur[0], o);
      if ((lane & (kTail - 1)) == 0) red[warp * RM + (lane >> (5 - LOGRM))] = cur[0];
      __syncthreads();
      if (tid < RM) {
        double rs = 0.0;
        for (int q = 0; q < NW; ++q) rs += red[q * RM + tid];
        u_s[tid] = tid < n ? fast_rcp(BdK * rs) : 0.0;   // cur = K rs: B rs = (B / K) cur, exact scaling
      }
      __syncthreads();
      if (it == a.iters - 1) break;
      double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;     // four partial chains (the scaling form is not order-bound)
#pragma unroll
      for (int i = 0; i < RM; i += 4) {
        c0 = fma(u_s[i], E[i], c0); c1 = fma(u_s[i + 1], E[i + 1], c1);
        c2 = fma(u_s[i + 2], E[i + 2], c2); c3 = fma(u_s[i + 3], E[i + 3], c3);
      }
      v = fast_rcp((c0 + c1) + (c2 + c3));               // K x column sum
    }
    // ---- literal last column step on the materialised plan, * B (see the warp kernel: q' = K q, q' / cs' == q / cs)
    double cs = 0.0;
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      const double q = __dmul_rn(__dmul_rn(u_s[i], E[i]), v);      // rows beyond n: +0
      E[i] = q;
      cs = __dadd_rn(cs, q);
    }
    const double rc = div_rcp(cs);
    const bool inexact = !div_den_ok(cs);
    unsigned loose = 0;                                  // FILTER: rows whose quotient is only approximate (see the warp kernel)
    double chk = 0.0;
    int kk[RM];
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      kk[i] = tid;
      if (i < n) {
        bool ok;
        double sh = div_by_rcp(E[i], cs, rc, ok);
        if constexpr (FILTER) loose |= ok ? 0u : (1u << i);
        else if (!ok) sh = __ddiv_rn(E[i], cs);
        const double val = __dmul_rn(__dmul_rn(sh, invK), Bd);
        E[i] = val;
        if constexpr (FILTER) my[(size_t)i * T] = val;   // the filter reads it back: the reduction below consumes E in place
        chk = fma(val, 0.0, chk);
      } else {
        E[i] = 0.0;
      }
    }
    bad = bad || (chk != chk);
    // ---- argmax per row (torch.argmax order): transposed reduction of (value, code) pairs
    double (&bv)[RM] = E;
#pragma unroll
    for (int st = 0; st < LOGRM; ++st) {
      const int o = 16 >> st;
      const int cnt = RM >> (st + 1);
      const bool upper = (lane & o) != 0;
#pragma unroll
      for (int j = 0; j < cnt; ++j) {
        const double keep_v = upper ? bv[j + cnt] : bv[j];
        const int keep_k = upper ? kk[j + cnt] : kk[j];
        const double send_v = upper ? bv[j] : bv[j + cnt];
        const int send_k = upper ? kk[j] : kk[j + cnt];
        const double ov = __shfl_xor_sync(0xffffffffu, send_v, o);
        const int ok = __shfl_xor_sync(0xffffffffu, send_k, o);
        const bool take = arg_better(ov, ok, keep_v, keep_k);
        bv[j] = take ? ov : keep_v;
        kk[j] = take ? ok : keep_k;
      }
    }
#pragma unroll
    for (int o = kTail >> 1; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv[0], o);
      const int ok = __shfl_xor_sync(0xffffffffu, kk[0], o);
      if (arg_better(ov, ok, bv[0], kk[0])) { bv[0] = ov; kk[0] = ok; }
    }
    if ((lane & (kTail - 1)) == 0) {
      red[warp * RM + (lane >> (5 - LOGRM))] = bv[0];
      red_k[warp * RM + (lane >> (5 - LOGRM))] = kk[0];
    }
    __syncthreads();
    if (tid < n) {
      double best = red[tid]; int bk = red_k[tid];
      for (int q = 1; q < NW; ++q) {
        const double ov = red[q * RM + tid]; const int ok = red_k[q * RM + tid];
        if (arg_better(ov, ok, best, bk)) { best = ov; bk = ok; }
      }
      row_v[tid] = best; row_k[tid] = bk;
      a.codes[a.members[beg + tid] * a.n_levels + a.level] = bk;
    }
    __syncthreads();
    if constexpr (FILTER) {                              // certainty filter of the warp kernel
      bool risky = (chk != chk) || inexact;
      const double scale = Kd / Bd;
#pragma unroll
      for (int i = 0; i < RM; ++i)
        if (i < n) {
          const double best = row_v[i];
          const double val = my[(size_t)i * T];
          if (val >= best - best * 2.1e-11 && ((loose >> i) & 1u)) risky = true;      // an approximate quotient competes
          if (val >= best - best * 2.1e-11 && tid != row_k[i]) {
            const double rowdev = fmax(0.0, 1.0 - best * scale) + 0x1p-50;
            const double dev = fmax(fmax(0.0, 1.0 - val * scale) + 0x1p-50, rowdev);
            if (dev > 0x1p-40 && val >= best - best * (0x1p-51 + 2e-11 * dev)) risky = true;
          }
          if (!(best == best)) risky = true;
        }
      const int any = __syncthreads_or(risky ? 1 : 0);
      if (any && tid == 0) a.risky_list[atomicAdd(a.risky_count, 1)] = (int32_t)g;
    }
  }
  if (__syncthreads_or(bad ? 1 : 0) && tid == 0) atomicOr(a.flags, 1);
}

}  // namespace lcrec
