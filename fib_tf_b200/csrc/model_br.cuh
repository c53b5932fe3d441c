// model_br.cuh -- modified 8-variable Beeler-Reuter model, pointwise part.
// Restates br.py:125-173 (solve), :175-205 (exact gates), :207-252 (Chebyshev gates),
// :255-273 (alpha/beta), :289-331 (scaled-monomial expansion), coefficients br.py:49-62.
#pragma once
#include "fib_kernels.cuh"

namespace fib {

// (c0 e^{c1(v+c2)} + c3 (v+c4)) / (e^{c5(v+c2)} + c6)  -- br.py:255-264.  Called with literal
// coefficients only, so every `== 0` test below folds at compile time (exp(0) == 1 exactly).
__device__ __forceinline__ float br_rate(float v, float c0, float c1, float c2, float c3, float c4,
                                         float c5, float c6) {
  const float e_num = (c1 == 0.f) ? 1.f : expf(c1 * (v + c2));
  float num = (c0 == 0.f) ? 0.f : c0 * e_num;
  if (c3 != 0.f) num = (c0 == 0.f) ? c3 * (v + c4) : num + c3 * (v + c4);
  const float e_den = (c5 == 0.f) ? 1.f : expf(c5 * (v + c2));
  return num / (e_den + c6);
}

// gate order: 0 xi, 1 m, 2 h, 3 j, 4 d, 5 f  (br.py:285-286); d/f rates doubled (br.py:46-48)
template <int G>
__device__ __forceinline__ void br_inf_tau_exact(float v, float& inf, float& tau) {
  float a, b;
  if (G == 0) { a = br_rate(v, 0.0005f, 0.083f, 50.f, 0.f, 0.f, 0.057f, 1.f);
                b = br_rate(v, 0.0013f, -0.06f, 20.f, 0.f, 0.f, -0.04f, 1.f); }
  if (G == 1) { a = br_rate(v, 0.f, 0.f, 47.f, -1.f, 47.f, -0.1f, -1.f);
                b = br_rate(v, 40.f, -0.056f, 72.f, 0.f, 0.f, 0.f, 0.f); }
  if (G == 2) { a = br_rate(v, 0.126f, -.25f, 77.f, 0.f, 0.f, 0.f, 0.f);
                b = br_rate(v, 1.7f, 0.f, 22.5f, 0.f, 0.f, -0.082f, 1.f); }
  if (G == 3) { a = br_rate(v, 0.055f, -.25f, 78.f, 0.f, 0.f, -0.2f, 1.f);
                b = br_rate(v, 0.3f, 0.f, 32.f, 0.f, 0.f, -0.1f, 1.f); }
  if (G == 4) { a = br_rate(v, (float)(2 * 0.095), -0.01f, -5.f, 0.f, 0.f, -0.072f, 1.f);
                b = br_rate(v, (float)(2 * 0.07), -0.017f, 44.f, 0.f, 0.f, 0.05f, 1.f); }
  if (G == 5) { a = br_rate(v, (float)(2 * 0.012), -0.008f, 28.f, 0.f, 0.f, 0.15f, 1.f);
                b = br_rate(v, (float)(2 * 0.0065), -0.02f, 30.f, 0.f, 0.f, -0.2f, 1.f); }
  const float ab = a + b;
  inf = a / ab;          // br.py:273
  tau = 1.0f / ab;
}

// r = d0 + d1 S1 + ... + d8 S8, left to right, each product and sum rounded (br.py:327-331).
// Uncontracted on purpose: the scaled-monomial basis is ill-conditioned, so an FMA's missing
// rounding is amplified well above 1 ulp; the reference's fp32 NumPy/TF evaluation is the target.
__device__ __forceinline__ float br_cheby_eval(const float* __restrict__ d, const float (&S)[9]) {
  float r = __fadd_rn(d[0], __fmul_rn(d[1], S[1]));
#pragma unroll
  for (int i = 2; i < 9; ++i) r = __fadd_rn(r, __fmul_rn(d[i], S[i]));
  return r;
}

template <bool CHEBY, bool SLOW>
struct BeelerReuter {
  static constexpr int NS = 7;            // C, M, H, J, D, F, XI  (V is the diffusing variable)
  static constexpr int VEC = 4;
  static constexpr int BY = 4;
  static constexpr int MAX_R = 4;
  static constexpr bool NEED_RAW = false; // everything sees V0 = enforce_boundary(V) (br.py:128)
  static constexpr bool NEED_LAP = true;
  static constexpr bool STORE_X = true;
  // slow gates J, D, F, XI are frozen when n == 0 (br.py:199-203): not even written back
  static __host__ __device__ constexpr bool stores(int k) { return k <= 2 || SLOW; }
  static size_t smem_bytes() { return 0; }
  struct Params {
    float dt;            // fp32(dt)
    float neg_dt;        // fp32(-dt)                  m, h
    float neg_dt_slow;   // fp32(-(dt*n))              xi, j, d, f when n > 0 (br.py:197-200)
    float ddt;           // fp32(diff*dt)
    float cheb[12][9];   // FIB_TABLE_BR_CHEBY (only read when CHEBY)
  };
  static __device__ __forceinline__ void prologue(const StepArgs<BeelerReuter>&) {}

  static __device__ __forceinline__ void cell(const StepArgs<BeelerReuter>& a, float /*raw*/,
                                              float V0, float lap, float (&s)[NS], float& Vnew) {
    const Params& p = a.p;
    const float C = s[0], M = s[1], H = s[2], J = s[3], D = s[4], F = s[5], XI = s[6];

    float inf[6], tau[6];
    if (CHEBY) {
      // x = (V0 - 0.5(max+min)) / (0.5(max-min)) = (V0 + 30)/60  (br.py:215), S_i = 2x S_{i-1}
      const float x = __fdiv_rn(V0 + 30.0f, 60.0f);
      const float x2 = 2.0f * x;
      float S[9];
      S[0] = 1.f; S[1] = x;
#pragma unroll
      for (int i = 2; i < 9; ++i) S[i] = __fmul_rn(x2, S[i - 1]);
#pragma unroll
      for (int g = 0; g < 6; ++g) {
        if (g == 1 || g == 2 || SLOW) {
          inf[g] = br_cheby_eval(p.cheb[2 * g], S);
          tau[g] = br_cheby_eval(p.cheb[2 * g + 1], S);
        }
      }
    } else {
      br_inf_tau_exact<1>(V0, inf[1], tau[1]);
      br_inf_tau_exact<2>(V0, inf[2], tau[2]);
      if (SLOW) {
        br_inf_tau_exact<0>(V0, inf[0], tau[0]);
        br_inf_tau_exact<3>(V0, inf[3], tau[3]);
        br_inf_tau_exact<4>(V0, inf[4], tau[4]);
        br_inf_tau_exact<5>(V0, inf[5], tau[5]);
      }
    }
    s[1] = rush_larsen(M, inf[1], tau[1], p.neg_dt);
    s[2] = rush_larsen(H, inf[2], tau[2], p.neg_dt);
    if (SLOW) {
      s[6] = rush_larsen(XI, inf[0], tau[0], p.neg_dt_slow);
      s[3] = rush_larsen(J, inf[3], tau[3], p.neg_dt_slow);
      s[4] = rush_larsen(D, inf[4], tau[4], p.neg_dt_slow);
      s[5] = rush_larsen(F, inf[5], tau[5], p.neg_dt_slow);
    }

    // currents from V0 and the OLD gates (br.py:150-165)
    const float iK1 = 0.35f * (4.f * (expf(0.04f * (V0 + 85.f)) - 1.f) /
                                   (expf(0.08f * (V0 + 53.f)) + expf(0.04f * (V0 + 53.f))) +
                               0.2f * ((V0 + 23.0f) / (1.0f - expf(-0.04f * (V0 + 23.f)))));
    const float ix1 = XI * 0.8f * (expf(0.04f * (V0 + 77.f)) - 1.f) / expf(0.04f * (V0 + 35.f));
    const float iNa = (4.0f * M * M * M * H * J + 0.005f) * (V0 - 50.0f);
    const float ECa = -82.3f - 13.0278f * logf(C);
    const float iCa = 0.09f * D * F * (V0 - ECa);
    const float I_sum = iK1 + ix1 + iNa + iCa;
    Vnew = clip_nan(fmaf(p.ddt, lap, V0) - p.dt * I_sum, -85.0f, 25.0f);
    const float dC = -1.0e-7f * iCa + 0.07f * (1.0e-7f - C);
    s[0] = fmaf(p.dt, dC, C);
  }
};

}  // namespace fib
