// k-means++ seeding on the device with scikit-learn's arithmetic (sklearn 1.9.0 `_kmeans_plusplus`; reference
// index/models/layers.py:69-82 -> KMeans(init="k-means++")), driven by random numbers the host drew up front from numpy's
// global RNG in sklearn's order (oracle.kmeanspp_draws: their count does not depend on the data).  Restated and pinned on the
// CPU by oracle.kmeanspp_predrawn (identical seeds, identical RNG state afterwards).
//
// Per further centre, three launches enqueued without any host read (everything data dependent stays on the device):
//   1. seed_candidates_kernel: lane t walks the SEQUENTIAL fp32 prefix sum of the closest squared distances (numpy's
//      `np.cumsum` on fp32) and records the first position whose prefix reaches u_t * potential (`np.searchsorted`, 'left');
//   2. seed_distances_kernel: squared distance of every row to every candidate in fp64 from the up-cast rows
//      (`-2 x.y + |x|^2 + |y|^2`, rounded to fp32, clipped at 0: `_euclidean_distances_upcast`), min with the closest distance;
//   3. seed_choose_kernel: new potential per candidate, the smallest wins (first on ties), closest distances / potential /
//      chosen row updated in place.
// The potentials are fp32 sums (sklearn: BLAS sdot / sgemv, summation order unspecified): a different order moves
// u * potential by ~1e-7 relative, which changes a candidate only if a prefix value lies inside that sliver.
#include "common.cuh"

namespace lcrec {

constexpr int kSeedMaxTrials = 32;

struct SeedState {
  float pot;                      // current potential
  int pad;
  long long cand[kSeedMaxTrials];
};

// closest[i] = fl32(max(|x_i - x_first|^2, 0)) in fp64 arithmetic, one warp per row; also used for the candidates
__device__ __forceinline__ float seed_sqdist(const float* __restrict__ x, const double* __restrict__ norms, int d, int64_t i,
                                             int64_t c, int lane) {
  double dot = 0.0;
  for (int j = lane; j < d; j += 32) dot += (double)x[i * d + j] * (double)x[c * d + j];
#pragma unroll
  for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  double v = -2.0 * dot;
  v += norms[c];                  // sklearn: d += XX (candidate norms), d += YY (row norms)
  v += norms[i];
  const float f = (float)v;
  return f > 0.f ? f : 0.f;
}

__global__ void seed_norms_kernel(const float* __restrict__ x, int64_t n, int d, double* __restrict__ norms) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  double s = 0.0;
  for (int j = lane; j < d; j += 32) { const double v = (double)x[i * d + j]; s += v * v; }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) norms[i] = s;
}

__global__ void seed_first_kernel(const float* __restrict__ x, const double* __restrict__ norms, int64_t n, int d, int64_t first,
                                  float* __restrict__ closest) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const float v = seed_sqdist(x, norms, d, i, first, lane);
  if (lane == 0) closest[i] = v;
}

// one CTA: potential = sum of closest (fp64 accumulate, rounded once to fp32)
__global__ void __launch_bounds__(1024)
seed_potential_kernel(const float* __restrict__ closest, int64_t n, SeedState* __restrict__ st) {
  __shared__ double red[1024];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) s += (double)closest[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) st->pot = (float)red[0];
}

// one warp: lane t < trials owns draw t
__global__ void __launch_bounds__(32)
seed_candidates_kernel(const float* __restrict__ closest, int64_t n, const double* __restrict__ draws, int trials,
                       SeedState* __restrict__ st) {
  const int t = threadIdx.x;
  const double target = t < trials ? draws[t] * (double)st->pot : 0.0;
  long long found = -1;
  float run = 0.f;
  for (int64_t i = 0; i < n; ++i) {
    run = __fadd_rn(run, closest[i]);             // np.cumsum on fp32: sequential, one rounding per element
    if (found < 0 && (double)run >= target) found = i;
  }
  if (t < trials) st->cand[t] = found < 0 ? n - 1 : found;      // searchsorted past the end, clipped like np.clip
}

// grid: (rows / 8, trials); one warp per (row, candidate)
__global__ void __launch_bounds__(256)
seed_distances_kernel(const float* __restrict__ x, const double* __restrict__ norms, int64_t n, int d,
                      const float* __restrict__ closest, const SeedState* __restrict__ st, float* __restrict__ dc) {
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31, t = blockIdx.y;
  if (i >= n) return;
  const float v = seed_sqdist(x, norms, d, i, st->cand[t], lane);
  if (lane == 0) dc[(int64_t)t * n + i] = fminf(closest[i], v);
}

// one CTA: potentials of the candidates, argmin (first on ties), commit the winner
__global__ void __launch_bounds__(1024)
seed_choose_kernel(const float* __restrict__ dc, int64_t n, int trials, float* __restrict__ closest, SeedState* __restrict__ st,
                   int64_t* __restrict__ indices, int c) {
  __shared__ double red[1024];
  __shared__ float pots[kSeedMaxTrials];
  __shared__ int best_s;
  for (int t = 0; t < trials; ++t) {
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) s += (double)dc[(int64_t)t * n + i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) pots[t] = (float)red[0];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int best = 0;
    for (int t = 1; t < trials; ++t)
      if (pots[t] < pots[best]) best = t;
    best_s = best;
    st->pot = pots[best];
    indices[c] = st->cand[best];
  }
  __syncthreads();
  const int best = best_s;
  for (int64_t i = threadIdx.x; i < n; i += 1024) closest[i] = dc[(int64_t)best * n + i];
}

__global__ void seed_gather_kernel(const float* __restrict__ x, const int64_t* __restrict__ indices, int k, int d,
                                   float* __restrict__ centers) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < (int64_t)k * d) centers[e] = x[indices[e / d] * d + e % d];
}

}  // namespace lcrec

using namespace lcrec;

extern "C" int64_t lcrec_kmeanspp_workspace_bytes(int64_t n, int n_trials) {
  if (n <= 0 || n_trials <= 0) return 256;
  return arena_need(n * 8) + arena_need(n * 4) + arena_need((int64_t)n_trials * n * 4) + arena_need(sizeof(SeedState)) + 256;
}

extern "C" int lcrec_kmeanspp_seed(const float* xc, int64_t n, int e_dim, int n_clusters, int64_t first_index,
                                   const double* draws, int n_trials, int64_t* indices, float* centers, void* workspace,
                                   int64_t workspace_bytes, void* stream) {
  LC_ARG(xc && indices && n > 0 && e_dim > 0 && n_clusters >= 1 && first_index >= 0 && first_index < n);
  LC_ARG(n_trials >= 1 && n_trials <= kSeedMaxTrials && (n_clusters == 1 || draws));
  LC_TRY(lcrec_device_check());
  cudaStream_t st = (cudaStream_t)stream;
  Arena a(workspace, workspace_bytes);
  double* norms = a.take<double>(n);
  float* closest = a.take<float>(n);
  float* dc = a.take<float>((int64_t)n_trials * n);
  SeedState* state = a.take<SeedState>(1);
  if (!a.ok()) { set_error("kmeanspp_seed: workspace too small"); return LCREC_ERR_NOMEM; }
  const unsigned row_grid = (unsigned)ceil_div(n * 32, 256);
  seed_norms_kernel<<<row_grid, 256, 0, st>>>(xc, n, e_dim, norms);
  LC_LAUNCH_CHECK("seed_norms_kernel");
  LC_CUDA(cudaMemcpyAsync(indices, &first_index, sizeof(int64_t), cudaMemcpyHostToDevice, st));
  seed_first_kernel<<<row_grid, 256, 0, st>>>(xc, norms, n, e_dim, first_index, closest);
  LC_LAUNCH_CHECK("seed_first_kernel");
  seed_potential_kernel<<<1, 1024, 0, st>>>(closest, n, state);
  LC_LAUNCH_CHECK("seed_potential_kernel");
  for (int c = 1; c < n_clusters; ++c) {
    seed_candidates_kernel<<<1, 32, 0, st>>>(closest, n, draws + (int64_t)(c - 1) * n_trials, n_trials, state);
    LC_LAUNCH_CHECK("seed_candidates_kernel");
    seed_distances_kernel<<<dim3((unsigned)ceil_div(n, 8), n_trials), 256, 0, st>>>(xc, norms, n, e_dim, closest, state, dc);
    LC_LAUNCH_CHECK("seed_distances_kernel");
    seed_choose_kernel<<<1, 1024, 0, st>>>(dc, n, n_trials, closest, state, indices, c);
    LC_LAUNCH_CHECK("seed_choose_kernel");
  }
  if (centers) {
    const int64_t total = (int64_t)n_clusters * e_dim;
    seed_gather_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(xc, indices, n_clusters, e_dim, centers);
    LC_LAUNCH_CHECK("seed_gather_kernel");
  }
  return LCREC_OK;
}
