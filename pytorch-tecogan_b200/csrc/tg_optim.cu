// Fused multi-tensor Adam on the flat parameter / gradient buckets (the optimizer of the reference's
// training driver, main.py:239-243: torch.optim.Adam(lr, betas=(beta, 0.999), eps), stepped under a GradScaler,
// code/train.py:335-342).
//
// Every parameter of a network is a view into ONE flat f32 buffer laid out like the flat gradient bucket the wgrad
// kernels add into (tecogan_b200.parallel), so the whole optimizer step is three launches per network and step:
//   tg_grad_check_finite  found_inf = any(!isfinite(grad))             (GradScaler.unscale_'s check, on the device)
//   tg_adam_step          p, exp_avg, exp_avg_sq updated in place; reads lr / step / 1/loss-scale / found_inf from
//                         device scalars, so nothing in the step depends on host values and the launch sequence can be
//                         captured in a CUDA graph and replayed while StepLR changes the learning rate
//   tg_scaler_update      GradScaler.update()'s growth / back-off of the loss scale + the optimizer's step counter
// followed by the batched bf16 re-pack of the updated weights (tg_gen_pack / tg_disc_pack read the same flat buffer,
// no gather copy).  All HBM-bound streaming kernels: 16-byte accesses, grid = a few waves of the SMs.
#include <math.h>

#include "tg_common.cuh"

namespace tg {

__global__ void __launch_bounds__(256)
grad_check_kernel(const float* __restrict__ g, long long n, float* __restrict__ found_inf) {
  bool bad = false;
  const long long n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(g4 + i);
    // x - x is 0 for every finite x and NaN for +-inf / NaN
    const float s = (v.x - v.x) + (v.y - v.y) + (v.z - v.z) + (v.w - v.w);
    bad |= !(s == 0.f);
  }
  if (blockIdx.x == 0 && threadIdx.x < static_cast<int>(n - n4 * 4)) {
    const float v = g[n4 * 4 + threadIdx.x];
    bad |= !((v - v) == 0.f);
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) *found_inf = 1.f;
}

// torch.optim.Adam, single-tensor formulation (amsgrad=False, maximize=False, weight_decay=0):
//   exp_avg = b1*exp_avg + (1-b1)*g ; exp_avg_sq = b2*exp_avg_sq + (1-b2)*g*g
//   p -= (lr / (1 - b1^t)) * exp_avg / (sqrt(exp_avg_sq) / sqrt(1 - b2^t) + eps)        t = step + 1
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
            const float* __restrict__ lr_p, float b1, float b2, float eps, const float* __restrict__ step_p,
            const float* __restrict__ inv_scale_p, const float* __restrict__ found_inf_p) {
  if (found_inf_p && *found_inf_p != 0.f) return;             // GradScaler.step skips the optimizer on inf / NaN gradients
  const float lr = *lr_p;
  const float t = *step_p + 1.f;
  const float inv_scale = inv_scale_p ? *inv_scale_p : 1.f;
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float step_size = lr / bc1, bc2_sqrt = sqrtf(bc2);
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
    const float gg[4] = {g4.x * inv_scale, g4.y * inv_scale, g4.z * inv_scale, g4.w * inv_scale};
    float pp[4] = {p4.x, p4.y, p4.z, p4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      mm[e] = b1 * mm[e] + (1.f - b1) * gg[e];
      vv[e] = b2 * vv[e] + (1.f - b2) * gg[e] * gg[e];
      pp[e] -= step_size * (mm[e] / (sqrtf(vv[e]) / bc2_sqrt + eps));
    }
    if (inv_scale_p) reinterpret_cast<float4*>(g)[i] = make_float4(gg[0], gg[1], gg[2], gg[3]);   // GradScaler.unscale_ is in place
    reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x < static_cast<int>(n - n4 * 4)) {
    const long long i = n4 * 4 + threadIdx.x;
    const float ge = g[i] * inv_scale;
    if (inv_scale_p) g[i] = ge;
    const float me = b1 * m[i] + (1.f - b1) * ge, ve = b2 * v[i] + (1.f - b2) * ge * ge;
    m[i] = me; v[i] = ve;
    p[i] -= step_size * (me / (sqrtf(ve) / bc2_sqrt + eps));
  }
}

// torch.amp.GradScaler.update() (_amp_update_scale_): back off on inf / NaN, grow after `interval` clean steps; and the
// optimizer's step counter, which GradScaler.step leaves untouched when it skips the step.
__global__ void scaler_update_kernel(float* scale, int* growth_tracker, const float* found_inf, float growth, float backoff,
                                     int interval, float* step) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const bool inf = found_inf && *found_inf != 0.f;
  if (step && !inf) *step += 1.f;
  if (!scale) return;
  if (inf) {
    *scale *= backoff;
    *growth_tracker = 0;
  } else {
    const int t = *growth_tracker + 1;
    if (t == interval) {
      const float ns = *scale * growth;
      if (isfinite(ns)) *scale = ns;                          // (torch keeps the old scale if growing would overflow)
      *growth_tracker = 0;
    } else {
      *growth_tracker = t;
    }
  }
}

static int stream_grid(long long n4) {
  long long blocks = (n4 + 255) / 256;
  const long long cap = static_cast<long long>(tg_num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks < 1 ? 1 : blocks);
}

}  // namespace tg

using namespace tg;

extern "C" int tg_grad_check_finite(const float* grads, long long n, float* found_inf, void* stream) {
  TG_CHECK_ARG(grads && found_inf && n >= 1, "grad_check_finite: bad arguments");
  TG_CHECK_ARG((reinterpret_cast<uintptr_t>(grads) & 15) == 0, "grad_check_finite: grads must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TG_CUDA(cudaMemsetAsync(found_inf, 0, sizeof(float), st));
  tg_prof_pre(TG_K_GLUE, 4.0 * n, st);
  grad_check_kernel<<<stream_grid(n / 4), 256, 0, st>>>(grads, n, found_inf);
  tg_prof_post(st);
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

extern "C" int tg_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                            const float* lr, float beta1, float beta2, float eps, const float* step,
                            const float* inv_scale, const float* found_inf, void* stream) {
  TG_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && lr && step && n >= 1, "adam_step: null pointer / empty");
  TG_CHECK_ARG(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
                 reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0, "adam_step: buffers must be 16-byte aligned");
  TG_CHECK_ARG(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, "adam_step: bad hyper-parameters");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  tg_prof_pre(TG_K_GLUE, 28.0 * n, st);                      // 4 f32 streams read, 3 written
  adam_kernel<<<stream_grid(n / 4), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step,
                                                  inv_scale, found_inf);
  tg_prof_post(st);
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

extern "C" int tg_scaler_update(float* scale, int* growth_tracker, const float* found_inf, float growth_factor,
                                float backoff_factor, int growth_interval, float* step, void* stream) {
  TG_CHECK_ARG((scale == nullptr) == (growth_tracker == nullptr), "scaler_update: scale and growth_tracker go together");
  TG_CHECK_ARG(scale || step, "scaler_update: nothing to update");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  tg_prof_pre(TG_K_GLUE, 16.0, st);
  scaler_update_kernel<<<1, 32, 0, st>>>(scale, growth_tracker, found_inf, growth_factor, backoff_factor, growth_interval, step);
  tg_prof_post(st);
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}
