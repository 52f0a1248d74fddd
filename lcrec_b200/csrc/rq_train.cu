// Training-side residual quantiser: forward values and analytic backward of ResidualVectorQuantizer.forward
// (reference index/models/rq.py:39-56 over vq.py:87-99) for GIVEN codes, as two native calls instead of ~40 autograd
// nodes per step.
//
// Forward (rq_chain_kernel, one warp per row; the chain is independent per dimension): r_0 = z;  per level
//   q = E_l[code_l], diff_l = q - r_l, x_res = r_l + diff_l (vq.py:95), r_{l+1} = r_l - x_res (rq.py:47), x_q += x_res,
//   sq_err_l += |diff_l|^2  (loss_l = mse + beta * mse, vq.py:90-92).
// Backward: the straight-through estimator makes x_res = r + sg(q - r), hence d r_{l+1} / d r_l = I - I = 0 and
//   d x_q / d z = I:  g_z = g_xq - (g_loss / L) * beta_0 * 2 / (n D) * diff_0   (commitment term of level 0 only),
//   g_E_l[k] = (g_loss / L) * 2 / (n D) * sum_{i: code_l(i) = k} diff_l(i)      (codebook term), summed in item order
//   (segsum.cuh: deterministic, no floating-point atomics - torch's embedding backward uses atomics).
#include <algorithm>

#include "common.cuh"
#include "segsum.cuh"

namespace lcrec {

struct RqTrainLevels {
  const float* codebook[LCREC_MAX_LEVELS];
  float* grad[LCREC_MAX_LEVELS];
  int first_block[LCREC_MAX_LEVELS + 1];       // backward grid: blocks [first_block[l], first_block[l+1]) own level l
};

__global__ void __launch_bounds__(256)
rq_chain_kernel(const float* __restrict__ z, const int64_t* __restrict__ codes, int64_t n, int d, int n_levels,
                RqTrainLevels lv, float* __restrict__ xq, float* __restrict__ diffs, int64_t* __restrict__ codes_t,
                double* __restrict__ sq_err) {
  __shared__ double red[LCREC_MAX_LEVELS][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 8 + warp;
  double err[LCREC_MAX_LEVELS];
#pragma unroll
  for (int l = 0; l < LCREC_MAX_LEVELS; ++l) err[l] = 0.0;
  if (i < n) {
    int64_t code[LCREC_MAX_LEVELS];
#pragma unroll
    for (int l = 0; l < LCREC_MAX_LEVELS; ++l)
      if (l < n_levels) {
        code[l] = codes[i * n_levels + l];
        if (lane == 0) codes_t[(int64_t)l * n + i] = code[l];
      }
    for (int j = lane; j < d; j += 32) {
      float r = z[i * d + j];
      float acc = 0.f;
#pragma unroll
      for (int l = 0; l < LCREC_MAX_LEVELS; ++l)
        if (l < n_levels) {
          const float q = lv.codebook[l][code[l] * d + j];
          const float diff = __fsub_rn(q, r);
          diffs[((int64_t)l * n + i) * d + j] = diff;
          err[l] += (double)diff * (double)diff;
          const float x_res = __fadd_rn(r, diff);
          r = __fsub_rn(r, x_res);
          acc = l == 0 ? x_res : __fadd_rn(acc, x_res);
        }
      xq[i * d + j] = acc;
    }
  }
#pragma unroll
  for (int l = 0; l < LCREC_MAX_LEVELS; ++l)
    if (l < n_levels) {
      double e = err[l];
#pragma unroll
      for (int o = 16; o; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
      if (lane == 0) red[l][warp] = e;
    }
  __syncthreads();
  if (threadIdx.x < n_levels) {
    double e = 0.0;
    for (int w = 0; w < 8; ++w) e += red[threadIdx.x][w];
    atomicAdd(sq_err + threadIdx.x, e);
  }
}

__global__ void __launch_bounds__(kSegThreads)
rq_codebook_grad_kernel(const float* __restrict__ diffs, const int64_t* __restrict__ codes_t, int64_t n, int d, int n_levels,
                        RqTrainLevels lv, const float* __restrict__ g_loss, float factor) {
  extern __shared__ float acc[];
  __shared__ SegSumSmem sm;
  int l = 0;
  while (l + 1 < n_levels && (int)blockIdx.x >= lv.first_block[l + 1]) ++l;
  const int k = blockIdx.x - lv.first_block[l];
  ordered_code_sum(diffs + (int64_t)l * n * d, codes_t + (int64_t)l * n, n, d, k, acc, sm);
  const float scale = __fmul_rn(g_loss ? *g_loss : 0.f, factor);
  for (int c = threadIdx.x; c < d; c += kSegThreads) lv.grad[l][(int64_t)k * d + c] = __fmul_rn(acc[c], scale);
}

__global__ void rq_latent_grad_kernel(const float* __restrict__ g_xq, const float* __restrict__ diff0,
                                      const float* __restrict__ g_loss, float factor, int64_t total, float* __restrict__ g_z) {
  const float scale = __fmul_rn(g_loss ? *g_loss : 0.f, factor);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    g_z[i] = fmaf(diff0[i], scale, g_xq ? g_xq[i] : 0.f);
}

}  // namespace lcrec

using namespace lcrec;

extern "C" int lcrec_rq_train_forward(const float* z, const int64_t* codes, int64_t n, int e_dim, int n_levels,
                                      const float* const* codebooks, float* xq, float* diffs, int64_t* codes_t,
                                      double* sq_err, void* stream) {
  LC_ARG(n >= 0 && e_dim > 0 && n_levels >= 1 && n_levels <= LCREC_MAX_LEVELS && codebooks && sq_err);
  LC_TRY(lcrec_device_check());
  cudaStream_t st = (cudaStream_t)stream;
  LC_CUDA(cudaMemsetAsync(sq_err, 0, sizeof(double) * n_levels, st));
  if (n == 0) return LCREC_OK;
  LC_ARG(z && codes && xq && diffs && codes_t);
  RqTrainLevels lv = {};
  for (int l = 0; l < n_levels; ++l) { LC_ARG(codebooks[l]); lv.codebook[l] = codebooks[l]; }
  rq_chain_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(z, codes, n, e_dim, n_levels, lv, xq, diffs, codes_t, sq_err);
  LC_LAUNCH_CHECK("rq_chain_kernel");
  return LCREC_OK;
}

extern "C" int lcrec_rq_train_backward(const float* diffs, const int64_t* codes_t, int64_t n, int e_dim, int n_levels,
                                       const int32_t* n_codes, const float* g_xq, const float* g_loss, double beta0,
                                       float* g_z, float* const* g_codebooks, void* stream) {
  LC_ARG(n > 0 && n < ((int64_t)1 << 31) && e_dim > 0 && e_dim <= 8192 && n_levels >= 1 && n_levels <= LCREC_MAX_LEVELS);
  LC_ARG(diffs && codes_t && n_codes);
  LC_TRY(lcrec_device_check());
  cudaStream_t st = (cudaStream_t)stream;
  const double per_elem = 2.0 / ((double)n * (double)e_dim) / (double)n_levels;   // d mse / d q, and the mean over levels
  if (g_codebooks) {
    RqTrainLevels lv = {};
    int blocks = 0;
    for (int l = 0; l < n_levels; ++l) {
      LC_ARG(g_codebooks[l] && n_codes[l] > 0);
      lv.grad[l] = g_codebooks[l];
      lv.first_block[l] = blocks;
      blocks += n_codes[l];
    }
    lv.first_block[n_levels] = blocks;
    rq_codebook_grad_kernel<<<blocks, kSegThreads, (size_t)e_dim * sizeof(float), st>>>(diffs, codes_t, n, e_dim, n_levels, lv,
                                                                                       g_loss, (float)per_elem);
    LC_LAUNCH_CHECK("rq_codebook_grad_kernel");
  }
  if (g_z) {
    const int64_t total = n * e_dim;
    rq_latent_grad_kernel<<<(unsigned)std::min<int64_t>(ceil_div(total, 256), 148 * 8), 256, 0, st>>>(
        g_xq, diffs, g_loss, (float)(-beta0 * per_elem), total, g_z);
    LC_LAUNCH_CHECK("rq_latent_grad_kernel");
  }
  return LCREC_OK;
}
