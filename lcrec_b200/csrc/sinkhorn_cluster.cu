// Dense Sinkhorn of the TRAINING forward (one (batch x K) problem per step, index/models/vq.py:76-83 inside
// index/trainer.py:114) as ONE thread-block cluster: the fp64 kernel exp(-d / eps) lives in the shared memory of up to 16 CTAs
// (64 rows x 256 columns x 8 B = 128 KB each at batch 1024), the row step is local, and the K column marginals of every
// iteration are exchanged through DISTRIBUTED SHARED MEMORY with one hardware cluster barrier - no grid-wide barrier, no
// global-memory round trip.  The cooperative-grid kernel it replaces on this path (sinkhorn_dense_kernel: in place in L2,
// four IEEE divides per element per iteration, one grid.sync per iteration) took 0.93 ms of the 4.1 ms of GPU time of a
// batch-1024 training step (ncu launch list, profiles/r2_train_launches_before.csv).
//
// Arithmetic = the scaling-vector form of the per-group kernels (sinkhorn.cu): Q = diag(u) E diag(v), u_i = 1 / (B sum_j E_ij v_j),
// v_j = 1 / (K sum_i u_i E_ij), 2 FMAs per element per iteration, and the reference's LAST column step evaluated literally on the
// materialised plan - ((q / colsum) / K) * B with rounded products, plain adds and IEEE divisions - so that the exact ties
// Q_ij == B / K of layers.py:85-108 come out as in the literal form.  Only the argmax leaves the kernel (what
// vq.py:83 consumes); `sinkhorn_algorithm()` as an API (the plan itself) stays on the literal kernel.
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace lcrec {

constexpr int kDcThreads = 256;
constexpr int kDcMaxCluster = 16;
constexpr int64_t kDcSmemBudget = 200 * 1024;

__device__ __forceinline__ double dc_rcp(double x) {        // same reciprocal as sinkhorn.cu::fast_rcp
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  e = fma(e, e, e);
  return fma(r, e, r);
}
__device__ __forceinline__ bool dc_arg_better(double a, int ia, double b, int ib) {     // torch.argmax order, NaN first
  const bool na = isnan(a), nb = isnan(b);
  if (na != nb) return na;
  if (na) return ia < ib;
  if (a != b) return a > b;
  return ia < ib;
}

struct SkClusterArgs {
  const double* dist; int64_t B; int K; double eps; int iters; int64_t* argmax; int32_t* flags; int rows_per_cta;
};

__global__ void __launch_bounds__(kDcThreads, 1) sinkhorn_dense_cluster_kernel(const SkClusterArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char dc_smem[];
  const int K = a.K, R = a.rows_per_cta;
  double* E = reinterpret_cast<double*>(dc_smem);          // R x K
  double* v = E + (size_t)R * K;                           // K
  double* part = v + K;                                    // 2 x K: this CTA's column partials, double-buffered by iteration parity
  double* u = part + 2 * K;                                // R
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kDcThreads / 32;
  const unsigned G = cluster.num_blocks(), rank = cluster.block_rank();
  const int64_t r0 = (int64_t)rank * R;
  const int nrows = (int)max((int64_t)0, min((int64_t)R, a.B - r0));
  const double Bd = (double)a.B, Kd = (double)K;

  for (int i = warp; i < nrows; i += nwarps)
    for (int k = lane; k < K; k += 32) E[(size_t)i * K + k] = exp(-(a.dist[(r0 + i) * K + k] / a.eps));      // layers.py:87
  for (int k = tid; k < K; k += kDcThreads) v[k] = 1.0;
  __syncthreads();

  for (int it = 0; it < a.iters; ++it) {
    for (int i = warp; i < nrows; i += nwarps) {                       // row step (local)
      double rs = 0.0;
      for (int k = lane; k < K; k += 32) rs = fma(E[(size_t)i * K + k], v[k], rs);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
      if (lane == 0) u[i] = dc_rcp(Bd * rs);
    }
    __syncthreads();
    if (it == a.iters - 1) break;
    double* mine = part + (size_t)(it & 1) * K;
    for (int k = tid; k < K; k += kDcThreads) {                        // this CTA's share of the column marginals
      double cs = 0.0;
      for (int i = 0; i < nrows; ++i) cs = fma(u[i], E[(size_t)i * K + k], cs);
      mine[k] = cs;
    }
    cluster.sync();                                                    // every CTA's partials of this iteration are in place
    for (int k = tid; k < K; k += kDcThreads) {
      double cs = 0.0;
      for (unsigned c = 0; c < G; ++c) cs += cluster.map_shared_rank(part, c)[(size_t)(it & 1) * K + k];     // rank order: same sum everywhere
      v[k] = dc_rcp(Kd * cs);
    }
    __syncthreads();
  }

  // literal last column step on the materialised plan: q = (u E) v with rounded products, column sums with plain adds
  double* mine = part + (size_t)(a.iters & 1) * K;
  for (int k = tid; k < K; k += kDcThreads) {
    const double vk = v[k];
    double cs = 0.0;
    for (int i = 0; i < nrows; ++i) cs = __dadd_rn(cs, __dmul_rn(__dmul_rn(u[i], E[(size_t)i * K + k]), vk));
    mine[k] = cs;
  }
  cluster.sync();
  double* colsum = part + (size_t)((a.iters + 1) & 1) * K;             // the other buffer: free since the last exchange
  for (int k = tid; k < K; k += kDcThreads) {
    double cs = 0.0;
    for (unsigned c = 0; c < G; ++c) cs = __dadd_rn(cs, cluster.map_shared_rank(part, c)[(size_t)(a.iters & 1) * K + k]);
    colsum[k] = cs;
  }
  __syncthreads();
  bool bad = false;
  for (int i = warp; i < nrows; i += nwarps) {
    double best = 0.0; int best_k = 0x7fffffff;
    const double ui = u[i];
    for (int k = lane; k < K; k += 32) {
      const double q = __dmul_rn(__dmul_rn(ui, E[(size_t)i * K + k]), v[k]);
      const double val = __dmul_rn(__ddiv_rn(__ddiv_rn(q, colsum[k]), Kd), Bd);      // ((q / colsum) / K) * B, layers.py:104-107
      bad = bad || isnan(val) || isinf(val);
      if (best_k == 0x7fffffff || dc_arg_better(val, k, best, best_k)) { best = val; best_k = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
      if (ok != 0x7fffffff && (best_k == 0x7fffffff || dc_arg_better(ob, ok, best, best_k))) { best = ob; best_k = ok; }
    }
    if (lane == 0) a.argmax[r0 + i] = best_k;
  }
  if (a.flags && __any_sync(0xffffffffu, bad) && lane == 0) atomicOr(a.flags, 1);
  cluster.sync();                                                      // no CTA may exit while a peer still reads its shared memory
}

static int g_dense_cluster = 1;

}  // namespace lcrec

using namespace lcrec;

// 1 (default): lcrec_sinkhorn_dense_argmax uses the cluster kernel when the problem fits; 0: always the literal kernel
extern "C" int lcrec_sinkhorn_set_dense_cluster(int on) { g_dense_cluster = on ? 1 : 0; return LCREC_OK; }

extern "C" int64_t lcrec_sinkhorn_dense_argmax_workspace_bytes(int64_t n_rows, int n_codes) {
  return lcrec_sinkhorn_workspace_bytes(n_rows, n_codes) + arena_need((int64_t)sizeof(double) * std::max<int64_t>(n_rows, 1) * n_codes);
}

// argmax_j of sinkhorn_algorithm(distances, epsilon, iters) per row (what vq.py:83 consumes), distances (n_rows x n_codes) fp64
// centred.  One cluster of <= 16 CTAs when n_rows / 16 x n_codes doubles fit shared memory (batch 1024 x 256 codes: 128 KB per
// CTA), otherwise the literal cooperative kernel with a scratch plan in the workspace.
extern "C" int lcrec_sinkhorn_dense_argmax(const double* distances, int64_t n_rows, int n_codes, double epsilon, int iters,
                                           int64_t* argmax, int32_t* flags, void* ws, int64_t ws_bytes, void* stream) {
  LC_ARG(n_rows >= 0 && n_codes > 0 && iters >= 0 && epsilon != 0.0);
  LC_TRY(lcrec_device_check());
  if (n_rows == 0) return LCREC_OK;
  LC_ARG(distances && argmax);
  cudaStream_t st = (cudaStream_t)stream;
  int G = 1;
  while (G < kDcMaxCluster && ceil_div(n_rows, G) > 32) G *= 2;
  const int R = (int)ceil_div(n_rows, G);
  const int64_t smem = (int64_t)sizeof(double) * ((int64_t)R * n_codes + 3 * (int64_t)n_codes + R) + 16;
  static int cluster_ok = -1;          // -1 unknown, 0 the device refused the launch once, 1 works
  if (g_dense_cluster && iters >= 1 && smem <= kDcSmemBudget && cluster_ok != 0) {
    static bool attr_set = false;
    if (!attr_set) {
      LC_CUDA(cudaFuncSetAttribute(sinkhorn_dense_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDcSmemBudget));
      LC_CUDA(cudaFuncSetAttribute(sinkhorn_dense_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      attr_set = true;
    }
    if (flags) LC_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t), st));
    SkClusterArgs a{distances, n_rows, n_codes, epsilon, iters, argmax, flags, R};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)G); cfg.blockDim = dim3(kDcThreads); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, sinkhorn_dense_cluster_kernel, a);
    if (e == cudaSuccess) { cluster_ok = 1; count_launch(); return LCREC_OK; }
    (void)cudaGetLastError();
    if (cluster_ok == 1) { set_error("sinkhorn_dense_cluster_kernel launch failed: %s", cudaGetErrorString(e)); return LCREC_ERR_CUDA; }
    cluster_ok = 0;                    // e.g. a partition without 16 co-schedulable SMs: use the literal kernel from now on
  }
  Arena ar(ws, ws_bytes);
  double* q = ar.take<double>(n_rows * n_codes);
  if (!ar.ok()) { set_error("sinkhorn_dense_argmax: workspace too small"); return LCREC_ERR_NOMEM; }
  char* rest = (char*)q + round_up((int64_t)sizeof(double) * n_rows * n_codes, 256);
  return lcrec_sinkhorn_dense(distances, n_rows, n_codes, epsilon, iters, q, argmax, flags, rest, ws_bytes - (rest - (char*)ws), stream);
}

// ---- fp64 FMA peak of this GPU, measured: the denominator of the per-group Sinkhorn's pipe fraction (MEASURED_PEAKS.json has no fp64
// figure).  Every thread runs 8 independent DFMA chains; 148 x 8 CTAs x 256 threads.  Returns FLOP/s through *flops_out (host).
namespace lcrec {
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* __restrict__ out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
}  // namespace lcrec

extern "C" int lcrec_fp64_peak_probe(double* flops_out, void* ws, int64_t ws_bytes, void* stream) {
  LC_ARG(flops_out != nullptr);
  LC_TRY(lcrec_device_check());
  cudaStream_t st = (cudaStream_t)stream;
  const int ctas = num_sms() * 8, iters = 1 << 14;
  LC_ARG(ws != nullptr && ws_bytes >= (int64_t)sizeof(double) * ctas * 256);
  cudaEvent_t e0, e1;
  LC_CUDA(cudaEventCreate(&e0)); LC_CUDA(cudaEventCreate(&e1));
  dfma_peak_kernel<<<ctas, 256, 0, st>>>((double*)ws, iters, 0.999999, 1e-9);      // warm-up
  LC_LAUNCH_CHECK("dfma_peak_kernel");
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    LC_CUDA(cudaEventRecord(e0, st));
    dfma_peak_kernel<<<ctas, 256, 0, st>>>((double*)ws, iters, 0.999999, 1e-9);
    LC_LAUNCH_CHECK("dfma_peak_kernel");
    LC_CUDA(cudaEventRecord(e1, st));
    LC_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    LC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    best = ms < best ? ms : best;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *flops_out = 2.0 * 8.0 * (double)iters * (double)ctas * 256.0 / ((double)best * 1e-3);
  return LCREC_OK;
}
