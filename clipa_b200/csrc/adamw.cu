// Fused AdamW step over a FLAT fp32 parameter segment (the optimizer of the reference recipe:
// torch.optim.AdamW, clipa_torch/training/main.py:318-326; decoupled weight decay, bias correction).
// One pass over HBM does everything that otherwise takes four: the update itself, the refresh of the
// bf16 shadow weights the GEMMs read, and zeroing the gradient buffer for the next step.
//   bytes per parameter: read p,g,m,v (16) + write p,m,v (12) + bf16 shadow (2) + zeroed g (4) = 34
#include "host_common.h"
#include "ptx.cuh"

namespace clipa {

__global__ void __launch_bounds__(256)
adamw_kernel(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
             uint2* __restrict__ p_bf16, long long n4, float lr, float beta1, float beta2, float eps,
             float decay_factor, float inv_bc1, float inv_sqrt_bc2, float grad_scale,
             const float* __restrict__ grad_scale_dev, int zero_grad) {
  if (grad_scale_dev) grad_scale *= __ldg(grad_scale_dev);   // e.g. the gradient-clipping factor, computed on the device
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = p[i], gv = g[i], mv = m[i], vv = v[i];
    float pa[4] = {pv.x, pv.y, pv.z, pv.w}, ga[4] = {gv.x, gv.y, gv.z, gv.w};
    float ma[4] = {mv.x, mv.y, mv.z, mv.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = ga[k] * grad_scale;
      pa[k] *= decay_factor;                              // p <- p * (1 - lr * wd)
      ma[k] = fmaf(beta1, ma[k], (1.0f - beta1) * gk);    // m <- b1 m + (1-b1) g
      va[k] = fmaf(beta2, va[k], (1.0f - beta2) * gk * gk);
      const float denom = fmaf(sqrtf(va[k]), inv_sqrt_bc2, eps);
      pa[k] -= lr * inv_bc1 * ma[k] / denom;
    }
    p[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    m[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    v[i] = make_float4(va[0], va[1], va[2], va[3]);
    if (p_bf16) p_bf16[i] = make_uint2(pack_bf16x2(pa[0], pa[1]), pack_bf16x2(pa[2], pa[3]));
    if (zero_grad) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

}  // namespace clipa

using namespace clipa;

extern "C" int clipa_adamw_step(void* param, void* grad, void* exp_avg, void* exp_avg_sq, void* param_bf16,
                                int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                                int32_t step, float grad_scale, const float* grad_scale_dev, int32_t zero_grad,
                                void* stream) {
  CLIPA_REQUIRE(param && grad && exp_avg && exp_avg_sq, CLIPA_ERR_BAD_ARG, "adamw_step: null pointer");
  CLIPA_REQUIRE(n > 0 && n % 4 == 0, CLIPA_ERR_BAD_ARG, "adamw_step: n must be a positive multiple of 4 (got %lld)",
                (long long)n);
  CLIPA_REQUIRE(step >= 1, CLIPA_ERR_BAD_ARG, "adamw_step: step counts from 1");
  const uintptr_t align = reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                          reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq);
  CLIPA_REQUIRE((align & 15) == 0 && (reinterpret_cast<uintptr_t>(param_bf16) & 7) == 0, CLIPA_ERR_BAD_ARG,
                "adamw_step: buffers must be 16-byte aligned (bf16 shadow 8-byte)");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  adamw_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<float4*>(param), static_cast<float4*>(grad), static_cast<float4*>(exp_avg),
      static_cast<float4*>(exp_avg_sq), static_cast<uint2*>(param_bf16), n4, lr, beta1, beta2, eps,
      1.0f - lr * weight_decay, (float)(1.0 / bc1), (float)(1.0 / sqrt(bc2)), grad_scale, grad_scale_dev, zero_grad);
  CLIPA_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return CLIPA_OK;
}
