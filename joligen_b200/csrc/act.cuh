// Activation helpers shared by the memory-bound normalisation kernels (norm.cu) and the fused GroupNorm-backward
// sums in the implicit-GEMM epilogue (conv_common.cuh).
#pragma once
#include <cuda_runtime.h>

#include "../../include/jg_b200.h"

namespace jg {

// sigmoid with ONE MUFU op (ex2) and 7 FMA-pipe + 1 ALU-pipe instructions per element.  These kernels run at HBM speed
// only if they stay near ~10 instructions per element (ncu: the first version, ~22 instructions/element, was issue-bound
// at 76% issue utilisation and 50% DRAM utilisation), and two MUFU ops per element (ex2 + rcp) would load the 16-lane
// special function unit to ~75%.  The reciprocal of x = 1 + e is an integer-subtract seed (5% error) refined by one
// cubically convergent step r*(1 + eps + eps^2), eps = 1 - x*r: relative error < 1.3e-4, 30x below a bf16 ulp.
// The saturating FMA (free modifier) keeps r in [0, 1] and flushes the NaN/inf that the seed produces for
// e >= 2^126 (u < -87) to 0, which is the correct limit: no clamp instruction is needed.
__device__ __forceinline__ float fast_sigmoid(float u) {
  float e;
  const float t = -1.4426950408889634f * u;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  const float x = 1.f + e;
  const float r = __int_as_float(0x7EF311C7 - __float_as_int(x));
  const float eps = fmaf(-x, r, 1.f);
  const float w = fmaf(eps, eps, eps);
  return __saturatef(fmaf(r, w, r));
}
// activation applied after the affine: none / SiLU (UNet) / ReLU, LeakyReLU(0.2) (GAN generator / discriminator);
// a template parameter so that the per-element code has no activation dispatch in it
template <int ACT>
__device__ __forceinline__ float act_f(float u) {
  if (ACT == JG_ACT_SILU) return u * fast_sigmoid(u);
  if (ACT == JG_ACT_RELU) return fmaxf(u, 0.f);
  if (ACT == JG_ACT_LRELU02) return u > 0.f ? u : 0.2f * u;
  return u;
}
template <int ACT>
__device__ __forceinline__ float act_grad(float u) {
  if (ACT == JG_ACT_SILU) {
    const float s = fast_sigmoid(u);
    return s * (1.f + fmaf(-u, s, u));  // s * (1 + u * (1 - s))
  }
  if (ACT == JG_ACT_RELU) return u > 0.f ? 1.f : 0.f;
  if (ACT == JG_ACT_LRELU02) return u > 0.f ? 1.f : 0.2f;
  return 1.f;
}
}  // namespace jg
