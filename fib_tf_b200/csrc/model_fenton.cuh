// model_fenton.cuh -- Cherry-Ehrlich-Nattel-Fenton 4-variable model, pointwise part.
// Restates fenton.py:46-92 (differentiate) and fenton.py:103-106 (explicit Euler).
#pragma once
#include "fib_kernels.cuh"

#ifndef FIB_4V_MINB
#define FIB_4V_MINB 8
#endif

namespace fib {

struct Fenton4v {
  static constexpr int NS = 3;            // V, W, S  (U is the diffusing variable)
  static constexpr int VEC = 4;
  static constexpr int VEC_SMALL = 4;   // cells per thread on small grids (kSmallGridCells)
  static constexpr int BY = 4;
  static constexpr int MAX_R = 4;
  static constexpr int AUTO_R = 4;   // marching depth picked by launch_step (measured best)
  static constexpr int MIN_BLOCKS = FIB_4V_MINB;
  static __host__ __device__ constexpr int min_blocks(int /*cells per thread*/) { return MIN_BLOCKS; }
  static constexpr bool PACKED = false;   // HBM-bound already: scalar cells, four per thread
  static constexpr bool PREFETCH = true;
  static constexpr bool NEED_RAW = true;  // reaction sees the raw U (fenton.py:101), SURVEY fact 3
  static constexpr bool NEED_LAP = true;
  static constexpr bool STORE_X = true;
  static __host__ __device__ constexpr bool stores(int) { return true; }
  static size_t smem_bytes() { return 0; }
  static const char* name() { return "Fenton4v"; }
  struct Params {
    float dt;    // fp32(dt)
    float ddt;   // fp32(diff * dt), folded in double like the reference (fenton.py:103)
  };
  static __device__ __forceinline__ void prologue(const StepArgs<Fenton4v>&) {}

  // Divisions by literals are multiplications by the correctly rounded reciprocal (<= 1 ulp from
  // the reference's fp32 division; budgeted in SURVEY.md Appendix B.2).
  static __device__ __forceinline__ void cell(const StepArgs<Fenton4v>& a, float U, float U0,
                                              float lap, float (&s)[NS], float& Unew) {
    constexpr float tau_vp = 3.33f, tau_vn = 19.2f, tau_wp = 160.0f, tau_wn = 75.0f;
    constexpr float tau_d = 0.065f, tau_si = 31.8364f, tau_so = 31.8364f, tau_a = 0.009f;
    constexpr float u_c = 0.23f, u_m = 1.0f, u_csi = 0.8f, u_so = 0.3f;
    constexpr float r_sn = 1.2f, k_ = 3.0f, b_so = 0.84f, c_so = 0.02f;
    constexpr float c_so_half = (float)(0.5 * (0.115 - 0.009));   // 0.5*(a_so - tau_a), fenton.py:83
    constexpr float r_diff = (float)(0.02 - 1.2);                 // (r_sp - r_sn),      fenton.py:89
    const float dt = a.p.dt;
    float V = s[0], W = s[1], S = s[2];

    // H(x) = (1+sign x)/2, G(x) = (1-sign x)/2 (fenton.py:73-79): 0.5 at x == 0
    const float Hc = U > u_c ? 1.f : (U < u_c ? 0.f : 0.5f);
    const float Hso = U > u_so ? 1.f : (U < u_so ? 0.f : 0.5f);
    const float Gso = 1.f - Hso;

    // Every multiply-add below is spelled out (fmaf / __f*_rn): whether ptxas fuses a free-standing
    // multiply and add depends on the surrounding kernel, and the two-steps-per-launch kernel
    // (fib_fused.cuh) must reproduce this one bit for bit.
    const float X_fi = __fmul_rn(__fmul_rn(__fmul_rn(-V, Hc), U - u_c), u_m - U);
    const float I_si = __fmul_rn(__fmul_rn(-W, S), 1.0f / tau_si);
    // 0.5 (a_so - tau_a) (1 + tanh z) = (a_so - tau_a) * [0.5 (1 + tanh z)]
    const float T_so = m_half_1p_tanh(__fmul_rn(U - b_so, 1.0f / c_so));
    const float I_so = fmaf(2.0f * c_so_half, T_so,
                            fmaf(__fmul_rn(U, Gso), 1.0f / tau_so, __fmul_rn(Hso, tau_a)));
    // dU = -(I_fi + I_si + I_so), I_fi = X_fi / tau_d
    const float dU = -__fadd_rn(fmaf(X_fi, 1.0f / tau_d, I_si), I_so);
    const float dV = U > u_c ? __fmul_rn(-V, 1.0f / tau_vp) : __fmul_rn(1.f - V, 1.0f / tau_vn);
    // tau_wn1 == tau_wn2 == 75 (fenton.py:53-54): the inner tf.where is an identity
    const float dW = U > u_c ? __fmul_rn(-W, 1.0f / tau_wp) : __fmul_rn(1.f - W, 1.0f / tau_wn);
    const float r_s = fmaf(r_diff, Hc, r_sn);
    const float dS = __fmul_rn(r_s, m_half_1p_tanh(__fmul_rn(U - u_csi, k_)) - S);

    // (U0 + dt*dU) + ddt*lap with the reference's rounding sequence (fenton.py:103): near U ~ 0 an
    // FMA's missing rounding would show up as a 1-ulp(|U0|) absolute difference
    Unew = __fadd_rn(__fadd_rn(U0, __fmul_rn(dt, dU)), __fmul_rn(a.p.ddt, lap));
    s[0] = fmaf(dt, dV, V);
    s[1] = fmaf(dt, dW, W);
    s[2] = fmaf(dt, dS, S);
  }
};

}  // namespace fib
