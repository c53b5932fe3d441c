// fib_math.cuh -- the transcendental / division layer of the ionic kernels.
//
// BR and Courtemanche are instruction-issue bound, not HBM bound: with
// CUDA's IEEE division (~10 instr + slow path) and libm-grade expf/expm1f/logf/tanhf (12-35
// instr) the BR cell costs ~700 instructions.  This layer replaces them by few-ulp versions built
// on the SFU approximations (MUFU.EX2 / LG2 / RCP), each <= ~3 ulp -- the same error class as
// swapping one fp32 libm for another, which is exactly what the parity tolerance is calibrated
// against (oracle/tfshim.ALT_LIBM, tests/test_gpu_parity.py).
// Build with -DFIB_ACCURATE_MATH=1 to get IEEE division and CUDA's libm instead (A/B checks).
#pragma once
#include <cuda_runtime.h>

#ifndef FIB_ACCURATE_MATH
#define FIB_ACCURATE_MATH 0
#endif

namespace fib {

__device__ __forceinline__ float sfu_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sfu_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sfu_lg2(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sfu_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

#if FIB_ACCURATE_MATH
__device__ __forceinline__ float m_rcp(float x) { return 1.0f / x; }
__device__ __forceinline__ float m_div(float a, float b) { return a / b; }
__device__ __forceinline__ float m_exp(float x) { return expf(x); }
__device__ __forceinline__ float m_expm1(float x) { return expm1f(x); }
__device__ __forceinline__ float m_expm1_neg(float x) { return expm1f(x); }
__device__ __forceinline__ float m_exp_affine(float x, float c, float add) { return c * expf(x) + add; }
__device__ __forceinline__ float m_log(float x) { return logf(x); }
__device__ __forceinline__ float m_sqrt(float x) { return sqrtf(x); }
// 1 / (1 + e^{-2z}) == 0.5 (1 + tanh z)
__device__ __forceinline__ float m_half_1p_tanh(float z) { return 0.5f * (1.f + tanhf(z)); }
#else
__device__ __forceinline__ float m_rcp(float x) { return sfu_rcp(x); }                 // 1 ulp
__device__ __forceinline__ float m_div(float a, float b) { return a * sfu_rcp(b); }    // 2 ulp

// e^x = 2^t * 2^r, t = fl(x*log2e), r = the rounding error of that product + x*lo(log2e),
// 2^r ~ 1 + r ln2.  ~2 ulp over the whole range; +inf / 0 on overflow / underflow like expf.
__device__ __forceinline__ float m_exp(float x) {
  const float L2E_HI = 1.44269502162933349609375f, L2E_LO = 1.92596299112661746e-8f;
  const float t = x * L2E_HI;
  float r = fmaf(x, L2E_HI, -t);
  r = fmaf(x, L2E_LO, r);
  return sfu_ex2(t) * fmaf(r, 0.693147182464599609375f, 1.0f);
}

// c e^x + add with the scale folded into the compensation factor and the add into an FMA: one or
// two instructions fewer than c * m_exp(x) + add (c and add are compile-time literals at the call
// sites; add == 0 folds to a multiply).
__device__ __forceinline__ float m_exp_affine(float x, float c, float add) {
  const float L2E_HI = 1.44269502162933349609375f, L2E_LO = 1.92596299112661746e-8f;
  const float t = x * L2E_HI;
  float r = fmaf(x, L2E_HI, -t);
  r = fmaf(x, L2E_LO, r);
  const float f = fmaf(r, c * 0.693147182464599609375f, c);
  return add == 0.f ? sfu_ex2(t) * f : fmaf(sfu_ex2(t), f, add);
}

// expm1: degree-6 Taylor polynomial for |x| < 0.125 (truncation x^6/7! < 1e-9 relative),
// exp(x)-1 beyond.  The polynomial side is the one that matters for accuracy: |x| = dt/tau is small
// exactly for the slow gates, whose per-step error would otherwise accumulate over hundreds of
// steps; for |x| >= 0.125 a gate relaxes within ~10 steps and exp(x)-1 (|result| >= 0.117, so
// < 8 ulp) is plenty.  Both sides are computed and selected: no divergence.
__device__ __forceinline__ float m_expm1(float x) {
  float p = 1.38888888888889e-3f;                 // 1/6!
  p = fmaf(p, x, 8.33333333333333e-3f);           // 1/5!
  p = fmaf(p, x, 4.16666666666667e-2f);           // 1/4!
  p = fmaf(p, x, 1.66666666666667e-1f);           // 1/3!
  p = fmaf(p, x, 0.5f);
  p = fmaf(p * x, x, x);                          // x + x^2 (1/2 + x/3! + ...)
  const float e = m_exp(x) - 1.0f;
  return fabsf(x) < 0.125f ? p : e;
}

// expm1 for the Rush-Larsen factor, x = -dt/tau <= 0.  On the far side the exponential needs no
// argument compensation: the uncompensated 2^(x log2e) is off by e^x (2 ulp + |x| 2^-24), i.e. an
// ABSOLUTE error <= 1.6e-7 (|x| e^x <= 1/e), against a result of magnitude >= 0.117 -- the same
// < 8 ulp as m_expm1, three instructions cheaper per gate.
__device__ __forceinline__ float m_expm1_neg(float x) {
  float p = 1.38888888888889e-3f;
  p = fmaf(p, x, 8.33333333333333e-3f);
  p = fmaf(p, x, 4.16666666666667e-2f);
  p = fmaf(p, x, 1.66666666666667e-1f);
  p = fmaf(p, x, 0.5f);
  p = fmaf(p * x, x, x);
  const float e = sfu_ex2(x * 1.44269502162933349609375f) - 1.0f;
  return fabsf(x) < 0.125f ? p : e;
}

__device__ __forceinline__ float m_log(float x) { return sfu_lg2(x) * 0.693147182464599609375f; }
__device__ __forceinline__ float m_sqrt(float x) { return sfu_sqrt(x); }
// 1 + e^{-2z} as ONE explicit FMA: left to the compiler, ptxas decides per kernel whether the
// multiply and the add fuse, and two kernels that must agree bit for bit (the one- and the
// two-steps-per-launch 4v kernels) would differ in the last place.
__device__ __forceinline__ float m_half_1p_tanh(float z) {
  return sfu_rcp(m_exp_affine(-2.0f * z, 1.0f, 1.0f));
}
#endif

}  // namespace fib
