// Gradient clipping + Adam / AdamW update of ALL parameters in two launches (reference index/trainer.py:117-119:
// clip_grad_norm_(model.parameters(), 1.0); optimizer.step() with torch.optim.AdamW / Adam, :49-81).
//
// HBM bound: pass 1 reads every gradient once (sum of squares, per-CTA partials in a fixed order); pass 2 re-derives the
// clip coefficient from the partials in every CTA (deterministic, no atomics, no host read), then reads g, p, m, v and writes
// p, m, v (and the clipped g, which clip_grad_norm_ leaves in .grad): 32 B per parameter against ~100 B for the foreach
// sequence of torch (norms, scale, lerp, mul, addcmul, sqrt, div, add, addcdiv).  The tensor table travels in the kernel
// arguments (gradient buffers move between steps, nothing is uploaded).
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace lcrec {

constexpr int kOptTensors = 40;                 // tensors per launch (kernel-argument table)
constexpr int kOptThreads = 256;
constexpr int kOptChunk = kOptThreads * 16;     // elements per CTA

struct OptTable {
  float* p[kOptTensors];
  float* g[kOptTensors];
  float* m[kOptTensors];
  float* v[kOptTensors];
  int first_block[kOptTensors + 1];
  long long numel[kOptTensors];
  int n;
};

__device__ __forceinline__ int opt_find(const OptTable& t, int block) {
  int i = 0;
  while (i + 1 < t.n && block >= t.first_block[i + 1]) ++i;
  return i;
}

__global__ void __launch_bounds__(kOptThreads)
grad_sqsum_kernel(OptTable t, double* __restrict__ partial) {
  __shared__ double red[kOptThreads / 32];
  const int ti = opt_find(t, blockIdx.x);
  const long long base = (long long)(blockIdx.x - t.first_block[ti]) * kOptChunk;
  const long long end = min(base + (long long)kOptChunk, t.numel[ti]);
  const float* __restrict__ g = t.g[ti];
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  long long i = base + threadIdx.x;
  for (; i + 3 * kOptThreads < end; i += 4 * kOptThreads) {
    const float a = g[i], b = g[i + kOptThreads], c = g[i + 2 * kOptThreads], d = g[i + 3 * kOptThreads];
    s0 = fmaf(a, a, s0); s1 = fmaf(b, b, s1); s2 = fmaf(c, c, s2); s3 = fmaf(d, d, s3);
  }
  for (; i < end; i += kOptThreads) { const float a = g[i]; s0 = fmaf(a, a, s0); }
  double s = ((double)s0 + (double)s1) + ((double)s2 + (double)s3);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < kOptThreads / 32; ++w) tot += red[w];
    partial[blockIdx.x] = tot;
  }
}

struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay;
  float one_minus_b1, one_minus_b2, decay;   // 1 - beta1, 1 - beta2, 1 - lr * weight_decay (Python doubles rounded once)
  float step_size;        // lr / (1 - beta1^t)
  float bc2_sqrt;         // sqrt(1 - beta2^t)
  float max_norm;         // <= 0: no clipping
  int decoupled;          // 1 = AdamW (p *= 1 - lr wd), 0 = Adam (g += wd p)
  int write_grad;         // store the clipped gradient back
  const float* hyper;     // nullable, device: {decay, step_size, bc2_sqrt} of THIS step; overrides the three fields above so that
                          // a captured CUDA graph of the training step can be replayed with a new learning rate / step count
};

__global__ void __launch_bounds__(kOptThreads)
adam_step_kernel(OptTable t, AdamArgs a, const double* __restrict__ partial, int n_partial, int block_offset,
                 float* __restrict__ total_norm_out) {
  __shared__ double red[kOptThreads / 32];
  __shared__ float coef_s;
  float coef = 1.f;
  if (a.max_norm > 0.f) {                        // clip_grad_norm_: coef = min(1, max_norm / (||g|| + 1e-6))
    double s = 0.0;
    for (int i = threadIdx.x; i < n_partial; i += kOptThreads) s += partial[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < kOptThreads / 32; ++w) tot += red[w];
      const float norm = (float)sqrt(tot);
      const float c = __fdiv_rn(a.max_norm, __fadd_rn(norm, 1e-6f));
      coef_s = c < 1.f ? c : 1.f;                // NaN propagates like torch.clamp(max=1)
      if (c != c) coef_s = c;
      if (blockIdx.x == 0 && block_offset == 0 && total_norm_out) *total_norm_out = norm;
    }
    __syncthreads();
    coef = coef_s;
  }
  const int ti = opt_find(t, blockIdx.x);
  const long long base = (long long)(blockIdx.x - t.first_block[ti]) * kOptChunk;
  const long long end = min(base + (long long)kOptChunk, t.numel[ti]);
  float* __restrict__ p = t.p[ti];
  float* __restrict__ g = t.g[ti];
  float* __restrict__ m = t.m[ti];
  float* __restrict__ v = t.v[ti];
  const float one_minus_b1 = a.one_minus_b1, one_minus_b2 = a.one_minus_b2;
  const float decay = a.hyper ? a.hyper[0] : a.decay;
  const float step_size = a.hyper ? a.hyper[1] : a.step_size;
  const float bc2_sqrt = a.hyper ? a.hyper[2] : a.bc2_sqrt;
  for (long long i = base + threadIdx.x; i < end; i += kOptThreads) {
    float gi = g[i];
    float pi = p[i];
    if (a.max_norm > 0.f) {
      gi = __fmul_rn(gi, coef);
      if (a.write_grad) g[i] = gi;
    }
    if (a.decoupled) pi = __fmul_rn(pi, decay);                    // param.mul_(1 - lr * weight_decay)
    else if (a.weight_decay != 0.f) gi = fmaf(pi, a.weight_decay, gi);   // grad.add(param, alpha=weight_decay)
    const float mi = fmaf(__fsub_rn(gi, m[i]), one_minus_b1, m[i]);       // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(__fmul_rn(gi, gi), one_minus_b2, __fmul_rn(v[i], a.beta2));   // mul_(beta2).addcmul_(g, g, 1 - beta2)
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), bc2_sqrt), a.eps);
    pi = fmaf(-step_size, __fdiv_rn(mi, denom), pi);                    // addcdiv_(exp_avg, denom, value=-step_size)
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
  }
}

// SGD / Adagrad / RMSprop of index/trainer.py:62-75 (torch.optim defaults: no momentum, lr_decay 0, alpha 0.99, not centred) with the
// same fused clipping prologue.  `s` = the optimiser's one state tensor per parameter: Adagrad's `sum`, RMSprop's `square_avg`
// (SGD: none).  torch semantics, fp32:  g *= coef;  g += wd p;
//   SGD:     p -= lr g
//   Adagrad: sum += g g;  p -= lr g / (sqrt(sum) + eps)
//   RMSprop: sq = alpha sq + (1 - alpha) g g;  p -= lr g / (sqrt(sq) + eps)
struct SimpleOptArgs { float lr, weight_decay, alpha, one_minus_alpha, eps, max_norm; int kind, write_grad; };

__global__ void __launch_bounds__(kOptThreads)
simple_opt_step_kernel(OptTable t, SimpleOptArgs a, const double* __restrict__ partial, int n_partial, int block_offset,
                       float* __restrict__ total_norm_out) {
  __shared__ double red[kOptThreads / 32];
  __shared__ float coef_s;
  float coef = 1.f;
  if (a.max_norm > 0.f) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n_partial; i += kOptThreads) s += partial[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < kOptThreads / 32; ++w) tot += red[w];
      const float norm = (float)sqrt(tot);
      const float c = __fdiv_rn(a.max_norm, __fadd_rn(norm, 1e-6f));
      coef_s = c < 1.f ? c : 1.f;
      if (c != c) coef_s = c;
      if (blockIdx.x == 0 && block_offset == 0 && total_norm_out) *total_norm_out = norm;
    }
    __syncthreads();
    coef = coef_s;
  }
  const int ti = opt_find(t, blockIdx.x);
  const long long base = (long long)(blockIdx.x - t.first_block[ti]) * kOptChunk;
  const long long end = min(base + (long long)kOptChunk, t.numel[ti]);
  float* __restrict__ p = t.p[ti];
  float* __restrict__ g = t.g[ti];
  float* __restrict__ st = t.m[ti];
  for (long long i = base + threadIdx.x; i < end; i += kOptThreads) {
    float gi = g[i];
    float pi = p[i];
    if (a.max_norm > 0.f) {
      gi = __fmul_rn(gi, coef);
      if (a.write_grad) g[i] = gi;
    }
    if (a.weight_decay != 0.f) gi = fmaf(pi, a.weight_decay, gi);            // grad.add(param, alpha=weight_decay)
    if (a.kind == 1) {
      pi = fmaf(-a.lr, gi, pi);                                               // param.add_(grad, alpha=-lr)
    } else if (a.kind == 2) {
      const float si = fmaf(gi, gi, st[i]);                                   // state_sum.addcmul_(grad, grad, value=1)
      st[i] = si;
      pi = fmaf(-a.lr, __fdiv_rn(gi, __fadd_rn(__fsqrt_rn(si), a.eps)), pi);  // addcdiv_(grad, std, value=-clr)
    } else {
      const float si = fmaf(__fmul_rn(gi, gi), a.one_minus_alpha, __fmul_rn(st[i], a.alpha));   // mul_(alpha).addcmul_(g, g, 1 - alpha)
      st[i] = si;
      pi = fmaf(-a.lr, __fdiv_rn(gi, __fadd_rn(__fsqrt_rn(si), a.eps)), pi);  // addcdiv_(grad, avg, value=-lr)
    }
    p[i] = pi;
  }
}

}  // namespace lcrec

using namespace lcrec;

// kind: 1 = SGD, 2 = Adagrad (state = sum), 3 = RMSprop (state = square_avg).  Same calling convention as lcrec_adam_clip_step;
// `state` may be NULL for SGD.  Workspace: lcrec_adam_workspace_bytes.
extern "C" int lcrec_simple_opt_clip_step(int kind, int n_tensors, float* const* params, float* const* grads, float* const* state,
                                          const int64_t* numel, double lr, double weight_decay, double alpha, double eps, double max_norm,
                                          int write_clipped_grads, float* total_norm_out, void* workspace, int64_t workspace_bytes,
                                          void* stream) {
  LC_ARG(kind >= 1 && kind <= 3 && n_tensors >= 0);
  if (n_tensors == 0) return LCREC_OK;
  LC_ARG(params && grads && numel && (kind == 1 || state));
  LC_TRY(lcrec_device_check());
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks_total = 0;
  for (int i = 0; i < n_tensors; ++i) {
    LC_ARG(numel[i] >= 0 && (numel[i] == 0 || (params[i] && grads[i] && (kind == 1 || state[i]))));
    blocks_total += ceil_div(numel[i], kOptChunk);
  }
  if (blocks_total == 0) return LCREC_OK;
  LC_ARG(blocks_total < ((int64_t)1 << 30));
  Arena ar(workspace, workspace_bytes);
  double* partial = ar.take<double>(blocks_total);
  if (!ar.ok()) { set_error("simple_opt_clip_step: workspace too small"); return LCREC_ERR_NOMEM; }
  SimpleOptArgs a;
  a.lr = (float)lr; a.weight_decay = (float)weight_decay; a.alpha = (float)alpha; a.one_minus_alpha = (float)(1.0 - alpha);
  a.eps = (float)eps; a.max_norm = (float)max_norm; a.kind = kind; a.write_grad = write_clipped_grads;
  for (int pass = max_norm > 0 ? 0 : 1; pass < 2; ++pass) {
    int64_t block_offset = 0;
    for (int t0 = 0; t0 < n_tensors;) {
      OptTable tab = {};
      int blocks = 0, k = 0;
      for (; t0 + k < n_tensors && k < kOptTensors; ++k) {
        tab.p[k] = params[t0 + k]; tab.g[k] = grads[t0 + k]; tab.m[k] = state ? state[t0 + k] : nullptr; tab.v[k] = nullptr;
        tab.numel[k] = numel[t0 + k];
        tab.first_block[k] = blocks;
        blocks += (int)ceil_div(numel[t0 + k], kOptChunk);
      }
      tab.first_block[k] = blocks;
      tab.n = k;
      if (blocks > 0) {
        if (pass == 0) {
          grad_sqsum_kernel<<<blocks, kOptThreads, 0, st>>>(tab, partial + block_offset);
          LC_LAUNCH_CHECK("grad_sqsum_kernel");
        } else {
          simple_opt_step_kernel<<<blocks, kOptThreads, 0, st>>>(tab, a, partial, (int)blocks_total, (int)block_offset, total_norm_out);
          LC_LAUNCH_CHECK("simple_opt_step_kernel");
        }
      }
      block_offset += blocks;
      t0 += k;
    }
  }
  return LCREC_OK;
}

extern "C" int64_t lcrec_adam_workspace_bytes(int n_tensors, const int64_t* numel) {
  int64_t blocks = 0;
  for (int i = 0; i < n_tensors; ++i) blocks += ceil_div(std::max<int64_t>(numel ? numel[i] : 0, 1), kOptChunk);
  return arena_need(blocks * 8) + 256;
}

// The three per-step scalars of the update as the kernel consumes them (fp32): {1 - lr * wd, lr / (1 - beta1^t), sqrt(1 - beta2^t)}
extern "C" int lcrec_adam_hyper(double lr, double beta1, double beta2, double weight_decay, int64_t step, float* out3_host) {
  LC_ARG(out3_host && step >= 1);
  out3_host[0] = (float)(1.0 - lr * weight_decay);
  out3_host[1] = (float)(lr / (1.0 - pow(beta1, (double)step)));
  out3_host[2] = (float)sqrt(1.0 - pow(beta2, (double)step));
  return LCREC_OK;
}

static int adam_clip_step_impl(int n_tensors, float* const* params, float* const* grads, float* const* exp_avg,
                               float* const* exp_avg_sq, const int64_t* numel, double lr, double beta1, double beta2,
                               double eps, double weight_decay, int decoupled, int64_t step, double max_norm,
                               int write_clipped_grads, float* total_norm_out, const float* hyper_dev, void* workspace,
                               int64_t workspace_bytes, void* stream) {
  LC_ARG(n_tensors >= 0 && step >= 1);
  if (n_tensors == 0) return LCREC_OK;
  LC_ARG(params && grads && exp_avg && exp_avg_sq && numel);
  LC_TRY(lcrec_device_check());
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks_total = 0;
  for (int i = 0; i < n_tensors; ++i) {
    LC_ARG(numel[i] >= 0 && (numel[i] == 0 || (params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i])));
    blocks_total += ceil_div(numel[i], kOptChunk);
  }
  if (blocks_total == 0) return LCREC_OK;
  LC_ARG(blocks_total < ((int64_t)1 << 30));
  Arena ar(workspace, workspace_bytes);
  double* partial = ar.take<double>(blocks_total);
  if (!ar.ok()) { set_error("adam_clip_step: workspace too small"); return LCREC_ERR_NOMEM; }

  AdamArgs a;
  a.lr = (float)lr; a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps; a.weight_decay = (float)weight_decay;
  a.one_minus_b1 = (float)(1.0 - beta1); a.one_minus_b2 = (float)(1.0 - beta2); a.decay = (float)(1.0 - lr * weight_decay);
  a.step_size = (float)(lr / (1.0 - pow(beta1, (double)step)));
  a.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
  a.max_norm = (float)max_norm;
  a.decoupled = decoupled;
  a.write_grad = write_clipped_grads;
  a.hyper = hyper_dev;

  // two sweeps over the tensor list in groups that fit the kernel-argument table: all partial sums first, then the updates
  for (int pass = max_norm > 0 ? 0 : 1; pass < 2; ++pass) {
    int64_t block_offset = 0;
    for (int t0 = 0; t0 < n_tensors;) {
      OptTable tab = {};
      int blocks = 0;
      int k = 0;
      for (; t0 + k < n_tensors && k < kOptTensors; ++k) {
        tab.p[k] = params[t0 + k]; tab.g[k] = grads[t0 + k]; tab.m[k] = exp_avg[t0 + k]; tab.v[k] = exp_avg_sq[t0 + k];
        tab.numel[k] = numel[t0 + k];
        tab.first_block[k] = blocks;
        blocks += (int)ceil_div(numel[t0 + k], kOptChunk);
      }
      tab.first_block[k] = blocks;
      tab.n = k;
      if (blocks > 0) {
        if (pass == 0) {
          grad_sqsum_kernel<<<blocks, kOptThreads, 0, st>>>(tab, partial + block_offset);
          LC_LAUNCH_CHECK("grad_sqsum_kernel");
        } else {
          adam_step_kernel<<<blocks, kOptThreads, 0, st>>>(tab, a, partial, (int)blocks_total, (int)block_offset, total_norm_out);
          LC_LAUNCH_CHECK("adam_step_kernel");
        }
      }
      block_offset += blocks;
      t0 += k;
    }
  }
  return LCREC_OK;
}

extern "C" int lcrec_adam_clip_step(int n_tensors, float* const* params, float* const* grads, float* const* exp_avg,
                                    float* const* exp_avg_sq, const int64_t* numel, double lr, double beta1, double beta2,
                                    double eps, double weight_decay, int decoupled, int64_t step, double max_norm,
                                    int write_clipped_grads, float* total_norm_out, void* workspace,
                                    int64_t workspace_bytes, void* stream) {
  return adam_clip_step_impl(n_tensors, params, grads, exp_avg, exp_avg_sq, numel, lr, beta1, beta2, eps, weight_decay, decoupled, step,
                             max_norm, write_clipped_grads, total_norm_out, nullptr, workspace, workspace_bytes, stream);
}

// Same update with the per-step scalars read from DEVICE memory (hyper_dev: 3 fp32 as lcrec_adam_hyper computes them): the
// launches can be captured once in a CUDA graph and replayed while the learning-rate schedule and the step count advance -
// the caller refreshes hyper_dev (one 12-byte copy on the same stream) before each replay.
extern "C" int lcrec_adam_clip_step_dev(int n_tensors, float* const* params, float* const* grads, float* const* exp_avg,
                                        float* const* exp_avg_sq, const int64_t* numel, const float* hyper_dev, double beta1,
                                        double beta2, double eps, double weight_decay, int decoupled, double max_norm,
                                        int write_clipped_grads, float* total_norm_out, void* workspace, int64_t workspace_bytes,
                                        void* stream) {
  LC_ARG(hyper_dev != nullptr);
  return adam_clip_step_impl(n_tensors, params, grads, exp_avg, exp_avg_sq, numel, 1e-3, beta1, beta2, eps, weight_decay, decoupled, 1,
                             max_norm, write_clipped_grads, total_norm_out, hyper_dev, workspace, workspace_bytes, stream);
}
