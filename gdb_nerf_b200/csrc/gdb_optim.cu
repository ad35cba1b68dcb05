// Training step tail on ONE flat fp32 buffer: gradient averaging over ranks, value
// clipping and the Adam update in a single pass.
//
// Reference: train/trainers/trainer.py:63-65 (`clip_grad_value_(parameters, 40)`; `optimizer.step()`), the optimiser built by
// train/optimizer.py:13-29 (torch.optim.Adam, one group per parameter, all with the same lr / weight_decay / eps), gradients
// averaged by DistributedDataParallel (trainer.py:16-22).  Same arithmetic and operation order as torch.optim.Adam's
// single-tensor path (weight decay folded into the gradient; exp_avg.lerp; denom = sqrt(exp_avg_sq) / sqrt(bias2) + eps;
// param -= lr / bias1 * exp_avg / denom).
//
// The step counter lives on the device (the whole training step is captured into a CUDA graph, so nothing the host changes
// between replays may enter a kernel argument): `state[0]` = number of completed steps, read by every thread and advanced
// by gdb_adam_advance after the update.
#include <algorithm>

#include "gdb_common.cuh"

namespace gdb {

__global__ void __launch_bounds__(256) adam_clip_kernel(float4* __restrict__ param, const float4* __restrict__ grad, float4* __restrict__ m,
                                                        float4* __restrict__ v, const float* __restrict__ state, int64_t n4, float lr,
                                                        float beta1, float beta2, float eps, float weight_decay, float clip,
                                                        float grad_scale) {
  const float t = state[0] + 1.f;
  const float bias1 = 1.f - powf(beta1, t), bias2 = 1.f - powf(beta2, t);
  const float step_size = lr / bias1;
  const float sb2 = sqrtf(bias2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 p = param[i], g = grad[i], mm = m[i], vv = v[i];
    float* pp = &p.x;
    float* gg = &g.x;
    float* pm = &mm.x;
    float* pv = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gk = gg[k] * grad_scale;                               // sum over ranks -> mean (DDP)
      gk = fminf(fmaxf(gk, -clip), clip);                          // clip_grad_value_
      gk = fmaf(weight_decay, pp[k], gk);
      pm[k] = fmaf(1.f - beta1, gk - pm[k], pm[k]);                // exp_avg.lerp_(grad, 1 - beta1)
      pv[k] = fmaf(1.f - beta2, gk * gk, beta2 * pv[k]);           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      const float denom = sqrtf(pv[k]) / sb2 + eps;
      pp[k] -= step_size * (pm[k] / denom);
    }
    param[i] = p; m[i] = mm; v[i] = vv;
  }
}

__global__ void adam_advance_kernel(float* state) { state[0] += 1.f; }

}  // namespace gdb

using namespace gdb;

extern "C" int gdb_adam_clip_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const float* state, int64_t n,
                                  float lr, float beta1, float beta2, float eps, float weight_decay, float clip_value, float grad_scale,
                                  void* stream) {
  GDB_REQUIRE(param && grad && exp_avg && exp_avg_sq && state && n > 0 && n % 4 == 0, GDB_E_BADARG,
              "gdb_adam_clip_step: null pointer or n (%lld) not a positive multiple of 4", (long long)n);
  GDB_REQUIRE(aligned16(param) && aligned16(grad) && aligned16(exp_avg) && aligned16(exp_avg_sq), GDB_E_ALIGN,
              "gdb_adam_clip_step: buffers must be 16-byte aligned");
  GDB_REQUIRE(lr > 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps > 0.f && clip_value > 0.f, GDB_E_BADARG,
              "gdb_adam_clip_step: bad hyper-parameter");
  const int64_t n4 = n / 4;
  const int blocks = (int)std::min<int64_t>((n4 + 255) / 256, (int64_t)sm_count() * 8);
  adam_clip_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<float4*>(param), reinterpret_cast<const float4*>(grad),
                                                         reinterpret_cast<float4*>(exp_avg), reinterpret_cast<float4*>(exp_avg_sq), state, n4,
                                                         lr, beta1, beta2, eps, weight_decay, clip_value, grad_scale);
  return cuda_check("gdb_adam_clip_step");
}

extern "C" int gdb_adam_advance(float* state, void* stream) {
  GDB_REQUIRE(state, GDB_E_BADARG, "gdb_adam_advance: null pointer");
  adam_advance_kernel<<<1, 1, 0, as_stream(stream)>>>(state);
  return cuda_check("gdb_adam_advance");
}
