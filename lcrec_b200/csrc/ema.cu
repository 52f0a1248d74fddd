// EMA codebook update, usage statistics (reference index_improve/models/vq.py:146-187, 205-217).
//
// One CTA owns one code k: ordered_code_sum (segsum.cuh) adds the rows assigned to k in ascending item order - the
// summation order of torch's CPU index_add_ (vq.py:166-167), so the sums are bit-identical to the reference and
// independent of the grid (deterministic, no floating-point atomics).  The same CTA
// then applies the smoothing and the convex codebook update with the reference's roundings: `x.mul_(decay)` is one
// rounding, `.add_(t, alpha=a)` is a fused multiply-add, `c * (1 - r) + n * r` is three roundings.
#include "common.cuh"
#include "segsum.cuh"

namespace lcrec {

constexpr int kEmaThreads = kSegThreads;

__global__ void __launch_bounds__(kEmaThreads)
ema_update_kernel(const float* __restrict__ latent, const int64_t* __restrict__ indices, int64_t n, int e_dim,
                  float decay, float alpha, float eps, float keep, float rate,
                  float* __restrict__ cluster_size, float* __restrict__ ema_w, float* __restrict__ codebook) {
  extern __shared__ float acc[];                 // e_dim running sums
  __shared__ SegSumSmem sm;
  const int k = blockIdx.x;
  const int tid = threadIdx.x;
  const int64_t count = ordered_code_sum(latent, indices, n, e_dim, k, acc, sm);

  // smoothing + codebook step (vq.py:155-184)
  const float cs = fmaf((float)count, alpha, __fmul_rn(cluster_size[k], decay));
  const bool used = cs > eps;
  const float denom = __fadd_rn(cs, eps);
  for (int d = tid; d < e_dim; d += kEmaThreads) {
    const int64_t o = (int64_t)k * e_dim + d;
    const float w = fmaf(acc[d], alpha, __fmul_rn(ema_w[o], decay));
    ema_w[o] = w;
    if (used) codebook[o] = __fadd_rn(__fmul_rn(codebook[o], keep), __fmul_rn(__fdiv_rn(w, denom), rate));
  }
  __syncthreads();                               // every thread has read cluster_size[k]
  if (tid == 0) cluster_size[k] = cs;
}

// usage = cs / (sum(cs) + eps); used = usage > threshold, unused = usage < threshold (vq.py:83-87, 208-211).
// The sum is taken in fp64 and rounded once (torch's fp32 tree sums differ between its own back ends by an ulp, which
// only matters for a code sitting exactly on the threshold).
__global__ void __launch_bounds__(256)
codebook_usage_kernel(const float* __restrict__ cluster_size, int n_codes, float eps, float threshold,
                      int64_t* __restrict__ used_codes, uint8_t* __restrict__ unused_mask) {
  __shared__ double part[8];
  __shared__ int used_part[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double s = 0.0;
  for (int i = tid; i < n_codes; i += 256) s += (double)cluster_size[i];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) part[warp] = s;
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += part[w];
  const float total = __fadd_rn((float)tot, eps);
  int used = 0;
  for (int i = tid; i < n_codes; i += 256) {
    const float u = __fdiv_rn(cluster_size[i], total);
    used += u > threshold;
    if (unused_mask) unused_mask[i] = u < threshold;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) used += __shfl_xor_sync(0xffffffffu, used, o);
  if (lane == 0) used_part[warp] = used;
  __syncthreads();
  if (tid == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += used_part[w];
    *used_codes = t;
  }
}

}  // namespace lcrec

using namespace lcrec;

extern "C" int lcrec_ema_update(const float* latent, const int64_t* indices, int64_t n, int n_codes, int e_dim,
                                double ema_decay, double epsilon, float* cluster_size, float* ema_w, float* codebook,
                                void* stream) {
  LC_ARG(n >= 0 && n < (int64_t)1 << 31);
  LC_ARG(n_codes > 0 && e_dim > 0 && e_dim <= 8192);
  LC_ARG(cluster_size && ema_w && codebook);
  LC_ARG(n == 0 || (latent && indices));
  LC_TRY(lcrec_device_check());
  cudaStream_t st = (cudaStream_t)stream;
  // Python doubles meet fp32 tensors as fp32 scalars: decay, 1 - decay, and for the convex step 1 - (1 - decay)
  const double update_rate = 1.0 - ema_decay;
  ema_update_kernel<<<n_codes, kEmaThreads, (size_t)e_dim * sizeof(float), st>>>(
      latent, indices, n, e_dim, (float)ema_decay, (float)(1.0 - ema_decay), (float)epsilon,
      (float)(1.0 - update_rate), (float)update_rate, cluster_size, ema_w, codebook);
  LC_LAUNCH_CHECK("ema_update_kernel");
  return LCREC_OK;
}

extern "C" int lcrec_codebook_usage(const float* cluster_size, int n_codes, double epsilon, double reset_threshold,
                                    int64_t* used_codes, uint8_t* unused_mask, void* stream) {
  LC_ARG(cluster_size && used_codes && n_codes > 0);
  LC_TRY(lcrec_device_check());
  codebook_usage_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(cluster_size, n_codes, (float)epsilon,
                                                            (float)reset_threshold, used_codes, unused_mask);
  LC_LAUNCH_CHECK("codebook_usage_kernel");
  return LCREC_OK;
}
