// Device Lloyd iterations for the k-means codebook initialisation (reference index/models/layers.py:69-82
// -> scikit-learn KMeans, pinned 1.9.0 in this image: sklearn/cluster/_kmeans.py `_kmeans_single_lloyd`).
//
// The seeding (k-means++) stays with scikit-learn on the host so that it consumes numpy's global RNG exactly like the
// reference; everything after it runs here with sklearn's structure: data centred by the column means, E-step = nearest
// centre (the residual quantiser's argmin kernel, lowest index on ties), M-step = per-cluster sums in item order
// (segsum.cuh: deterministic, no floating-point atomics) times the reciprocal count, empty clusters take the points
// farthest from their centres, stop on unchanged labels (strict) or sum of squared centre shifts <= tol, final E-step
// when the stop was not strict.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "segsum.cuh"

namespace lcrec {

struct KmeansStatus {
  long long changed;      // labels that differ from the previous iteration
  long long n_empty;      // clusters without a point
  double shift_tot;       // sum over clusters of |new - old|^2
};

// ---- column statistics (fp64 partials in a fixed order), centring ---------------------------------------------------------
constexpr int kStatRows = 8;
__global__ void __launch_bounds__(32 * kStatRows)
col_partial_kernel(const float* __restrict__ x, int64_t n, int d, int rows_per_split, double* __restrict__ psum,
                   double* __restrict__ psq) {
  __shared__ double s1[kStatRows][33], s2[kStatRows][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int r = threadIdx.x >> 5;
  const int64_t t0 = (int64_t)blockIdx.y * rows_per_split;
  const int64_t t1 = t0 + rows_per_split < n ? t0 + rows_per_split : n;
  double a = 0.0, b = 0.0;
  if (c < d)
    for (int64_t t = t0 + r; t < t1; t += kStatRows) {
      const double v = (double)x[t * d + c];
      a += v;
      b += v * v;
    }
  s1[r][threadIdx.x & 31] = a;
  s2[r][threadIdx.x & 31] = b;
  __syncthreads();
  if (r == 0 && c < d) {
    for (int i = 1; i < kStatRows; ++i) { a += s1[i][threadIdx.x]; b += s2[i][threadIdx.x]; }
    psum[(int64_t)blockIdx.y * d + c] = a;
    psq[(int64_t)blockIdx.y * d + c] = b;
  }
}

__global__ void col_final_kernel(const double* __restrict__ psum, const double* __restrict__ psq, int n_splits, int64_t n,
                                 int d, float* __restrict__ mean, double* __restrict__ var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  double a = 0.0, b = 0.0;
  for (int s = 0; s < n_splits; ++s) { a += psum[(int64_t)s * d + c]; b += psq[(int64_t)s * d + c]; }
  const double m = a / (double)n;
  mean[c] = (float)m;
  var[c] = fmax(b / (double)n - m * m, 0.0);
}

__global__ void center_rows_kernel(const float* __restrict__ x, const float* __restrict__ mean, int64_t total, int d,
                                   float* __restrict__ xc) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    xc[i] = __fsub_rn(x[i], mean[i % d]);
}

// ---- M-step: one CTA per cluster --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSegThreads)
kmeans_update_kernel(const float* __restrict__ x, const int64_t* __restrict__ labels, int64_t n, int d,
                     const float* __restrict__ centers_old, float* __restrict__ centers_new, long long* __restrict__ counts,
                     float* __restrict__ shift2) {
  extern __shared__ float acc[];
  __shared__ SegSumSmem sm;
  __shared__ float red[kSegThreads];
  const int k = blockIdx.x, tid = threadIdx.x;
  const int64_t count = ordered_code_sum(x, labels, n, d, k, acc, sm);
  // sklearn `_average_centers`: centers[j] *= 1.0 / weight_in_clusters[j]; an empty cluster keeps a zero sum until relocated
  const float alpha = count > 0 ? __fdiv_rn(1.0f, (float)count) : 0.f;
  float part = 0.f;
  for (int c = tid; c < d; c += kSegThreads) {
    const float v = count > 0 ? __fmul_rn(acc[c], alpha) : centers_old[(int64_t)k * d + c];
    centers_new[(int64_t)k * d + c] = v;
    const float diff = __fsub_rn(v, centers_old[(int64_t)k * d + c]);
    part = fmaf(diff, diff, part);
  }
  red[tid] = part;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int i = 0; i < kSegThreads; ++i) s += red[i];
    shift2[k] = s;
    counts[k] = count;
  }
}

// labels_old <- labels, number of changed labels; block 0 also totals the shifts and counts the empty clusters
__global__ void __launch_bounds__(256)
kmeans_status_kernel(const int64_t* __restrict__ labels, int64_t* __restrict__ labels_old, int64_t n,
                     const long long* __restrict__ counts, const float* __restrict__ shift2, int n_codes,
                     KmeansStatus* __restrict__ st) {
  __shared__ double sred[256];
  __shared__ int ered[256];
  int changed = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t l = labels[i];
    changed += l != labels_old[i];
    labels_old[i] = l;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) changed += __shfl_xor_sync(0xffffffffu, changed, o);
  if ((threadIdx.x & 31) == 0 && changed) atomicAdd((unsigned long long*)&st->changed, (unsigned long long)changed);
  if (blockIdx.x == 0) {
    double s = 0.0;
    int e = 0;
    for (int k = threadIdx.x; k < n_codes; k += 256) { s += (double)shift2[k]; e += counts[k] == 0; }
    sred[threadIdx.x] = s;
    ered[threadIdx.x] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      int te = 0;
      for (int i = 0; i < 256; ++i) { t += sred[i]; te += ered[i]; }
      st->shift_tot = t;
      st->n_empty = te;
    }
  }
}

// ---- empty clusters (sklearn `_relocate_empty_clusters_dense`) ---------------------------------------------------------------
__global__ void point_dist_kernel(const float* __restrict__ x, const int64_t* __restrict__ labels, int64_t n, int d,
                                  const float* __restrict__ centers, float* __restrict__ dist) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const float* c = centers + labels[i] * d;
  float s = 0.f;
  for (int j = lane; j < d; j += 32) { const float t = __fsub_rn(x[i * d + j], c[j]); s = fmaf(t, t, s); }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) dist[i] = s;
}

// one CTA: the j-th empty cluster (ascending id) takes the j-th farthest point (ties: lowest item); labels_tmp gets the move
__global__ void __launch_bounds__(1024)
relocate_kernel(float* __restrict__ dist, int64_t n, const long long* __restrict__ counts, int n_codes,
                int64_t* __restrict__ labels_tmp) {
  __shared__ float bv[32];
  __shared__ long long bi[32];
  const int tid = threadIdx.x;
  for (int k = 0; k < n_codes; ++k) {
    if (counts[k] != 0) continue;                              // uniform across the CTA
    float v = -1.f;
    long long idx = -1;
    for (int64_t i = tid; i < n; i += 1024) {
      const float q = dist[i];
      if (q > v) { v = q; idx = i; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi >= 0 && (idx < 0 || oi < idx))) { v = ov; idx = oi; }
    }
    if ((tid & 31) == 0) { bv[tid >> 5] = v; bi[tid >> 5] = idx; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 32; ++w)
        if (bv[w] > v || (bv[w] == v && bi[w] >= 0 && (idx < 0 || bi[w] < idx))) { v = bv[w]; idx = bi[w]; }
      if (idx >= 0) { labels_tmp[idx] = k; dist[idx] = -2.f; }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024)
inertia_kernel(const float* __restrict__ dist, int64_t n, double* __restrict__ out) {
  __shared__ double red[1024];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) s += (double)dist[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}

__global__ void add_mean_kernel(float* __restrict__ centers, const float* __restrict__ mean, int64_t total, int d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) centers[i] = __fadd_rn(centers[i], mean[i % d]);
}

static int stat_splits(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(64, ceil_div(n, 256))); }

}  // namespace lcrec

using namespace lcrec;

extern "C" int64_t lcrec_kmeans_workspace_bytes(int64_t n, int e_dim, int n_codes) {
  if (n < 0 || e_dim <= 0 || n_codes <= 0) return 256;
  int64_t b = 0;
  b += 3 * arena_need(n * 8);                               // labels, previous labels, labels after relocation
  b += arena_need((int64_t)n_codes * e_dim * 4);            // centres of the next iteration
  b += arena_need((int64_t)n_codes * 8) + arena_need((int64_t)n_codes * 4);
  b += arena_need(n * 4);                                   // point-to-centre distances
  b += 2 * arena_need(64 * (int64_t)e_dim * 8) + arena_need((int64_t)e_dim * 8);   // column statistics
  b += 4 * arena_need(64);
  return b;
}

extern "C" int lcrec_kmeans_center(const float* x, int64_t n, int e_dim, float* xc, float* mean, double* mean_variance_host,
                                   void* workspace, int64_t workspace_bytes, void* stream) {
  LC_ARG(x && xc && mean && n > 0 && e_dim > 0);
  LC_TRY(lcrec_device_check());
  cudaStream_t st = (cudaStream_t)stream;
  Arena a(workspace, workspace_bytes);
  const int splits = stat_splits(n);
  double* psum = a.take<double>((int64_t)splits * e_dim);
  double* psq = a.take<double>((int64_t)splits * e_dim);
  double* var = a.take<double>(e_dim);
  if (!a.ok()) { set_error("kmeans_center: workspace too small"); return LCREC_ERR_NOMEM; }
  const int rows_per_split = (int)ceil_div(n, splits);
  col_partial_kernel<<<dim3((unsigned)ceil_div(e_dim, 32), splits), 32 * kStatRows, 0, st>>>(x, n, e_dim, rows_per_split, psum, psq);
  LC_LAUNCH_CHECK("col_partial_kernel");
  col_final_kernel<<<(unsigned)ceil_div(e_dim, 128), 128, 0, st>>>(psum, psq, splits, n, e_dim, mean, var);
  LC_LAUNCH_CHECK("col_final_kernel");
  const int64_t total = n * e_dim;
  center_rows_kernel<<<(unsigned)std::min<int64_t>(ceil_div(total, 256), 148 * 16), 256, 0, st>>>(x, mean, total, e_dim, xc);
  LC_LAUNCH_CHECK("center_rows_kernel");
  if (mean_variance_host) {                                   // sklearn `_tolerance`: mean of the column variances
    std::vector<double> h(e_dim);
    LC_CUDA(cudaMemcpyAsync(h.data(), var, sizeof(double) * e_dim, cudaMemcpyDeviceToHost, st));
    LC_CUDA(cudaStreamSynchronize(st));
    double s = 0.0;
    for (double v : h) s += v;
    *mean_variance_host = s / e_dim;
  }
  return LCREC_OK;
}

extern "C" int lcrec_kmeans_lloyd(const float* xc, int64_t n, int e_dim, float* centers, int n_codes, int max_iter,
                                  double tol, const float* add_mean, int64_t* labels_out, double* inertia_host,
                                  int* n_iter_host, void* workspace, int64_t workspace_bytes, void* stream) {
  LC_ARG(xc && centers && n > 0 && n < ((int64_t)1 << 31) && e_dim > 0 && e_dim <= 8192 && n_codes > 0 && max_iter >= 1);
  LC_TRY(lcrec_device_check());
  cudaStream_t st = (cudaStream_t)stream;
  Arena a(workspace, workspace_bytes);
  int64_t* labels = a.take<int64_t>(n);
  int64_t* labels_old = a.take<int64_t>(n);
  int64_t* labels_tmp = a.take<int64_t>(n);
  float* cnew = a.take<float>((int64_t)n_codes * e_dim);
  long long* counts = a.take<long long>(n_codes);
  float* shift2 = a.take<float>(n_codes);
  float* dist = a.take<float>(n);
  KmeansStatus* status = a.take<KmeansStatus>(1);
  double* inertia = a.take<double>(1);
  if (!a.ok()) { set_error("kmeans_lloyd: workspace too small"); return LCREC_ERR_NOMEM; }

  float* cur = centers;          // centres the E-step uses
  float* nxt = cnew;
  const int32_t ks[1] = {n_codes};
  auto e_step = [&](const float* c) -> int {
    const float* cbs[1] = {c};
    return lcrec_rq_quantize(xc, n, e_dim, 1, cbs, ks, 1, -1, labels, nullptr, nullptr, nullptr, stream);
  };
  const size_t smem = (size_t)e_dim * sizeof(float);
  const unsigned status_grid = (unsigned)std::min<int64_t>(ceil_div(n, 256), 148 * 8);
  LC_CUDA(cudaMemsetAsync(labels_old, 0xff, sizeof(int64_t) * n, st));            // -1: sklearn's initial labels
  bool strict = false;
  int it = 0;
  for (; it < max_iter; ++it) {
    LC_TRY(e_step(cur));
    kmeans_update_kernel<<<n_codes, kSegThreads, smem, st>>>(xc, labels, n, e_dim, cur, nxt, counts, shift2);
    LC_LAUNCH_CHECK("kmeans_update_kernel");
    LC_CUDA(cudaMemsetAsync(status, 0, sizeof(KmeansStatus), st));
    kmeans_status_kernel<<<status_grid, 256, 0, st>>>(labels, labels_old, n, counts, shift2, n_codes, status);
    LC_LAUNCH_CHECK("kmeans_status_kernel");
    KmeansStatus h;
    LC_CUDA(cudaMemcpyAsync(&h, status, sizeof(h), cudaMemcpyDeviceToHost, st));
    LC_CUDA(cudaStreamSynchronize(st));
    if (h.n_empty > 0) {
      LC_CUDA(cudaMemcpyAsync(labels_tmp, labels, sizeof(int64_t) * n, cudaMemcpyDeviceToDevice, st));
      point_dist_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, st>>>(xc, labels, n, e_dim, cur, dist);
      LC_LAUNCH_CHECK("point_dist_kernel");
      relocate_kernel<<<1, 1024, 0, st>>>(dist, n, counts, n_codes, labels_tmp);
      LC_LAUNCH_CHECK("relocate_kernel");
      kmeans_update_kernel<<<n_codes, kSegThreads, smem, st>>>(xc, labels_tmp, n, e_dim, cur, nxt, counts, shift2);
      LC_LAUNCH_CHECK("kmeans_update_kernel");
      LC_CUDA(cudaMemsetAsync(status, 0, sizeof(KmeansStatus), st));
      kmeans_status_kernel<<<status_grid, 256, 0, st>>>(labels, labels_old, n, counts, shift2, n_codes, status);
      LC_LAUNCH_CHECK("kmeans_status_kernel");
      const long long changed = h.changed;                     // labels_old already equals labels: keep the first count
      LC_CUDA(cudaMemcpyAsync(&h, status, sizeof(h), cudaMemcpyDeviceToHost, st));
      LC_CUDA(cudaStreamSynchronize(st));
      h.changed = changed;
    }
    std::swap(cur, nxt);                                       // sklearn: centers, centers_new = centers_new, centers
    if (h.changed == 0) { strict = true; ++it; break; }
    if (h.shift_tot <= tol) { ++it; break; }
  }
  if (!strict) LC_TRY(e_step(cur));                            // labels consistent with the final centres
  if (inertia_host) {
    point_dist_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, st>>>(xc, labels, n, e_dim, cur, dist);
    LC_LAUNCH_CHECK("point_dist_kernel");
    inertia_kernel<<<1, 1024, 0, st>>>(dist, n, inertia);
    LC_LAUNCH_CHECK("inertia_kernel");
    LC_CUDA(cudaMemcpyAsync(inertia_host, inertia, sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  if (cur != centers)
    LC_CUDA(cudaMemcpyAsync(centers, cur, sizeof(float) * (size_t)n_codes * e_dim, cudaMemcpyDeviceToDevice, st));
  if (add_mean) {
    const int64_t total = (int64_t)n_codes * e_dim;
    add_mean_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(centers, add_mean, total, e_dim);
    LC_LAUNCH_CHECK("add_mean_kernel");
  }
  if (labels_out) LC_CUDA(cudaMemcpyAsync(labels_out, labels, sizeof(int64_t) * n, cudaMemcpyDeviceToDevice, st));
  LC_CUDA(cudaStreamSynchronize(st));
  if (n_iter_host) *n_iter_host = std::min(it, max_iter);
  return LCREC_OK;
}
