/*
 * TEST INFRASTRUCTURE (see oracle/README.md): plain-C + OpenMP port of the oracle
 * (oracle/monodomain_np.py), used ONLY as the timed CPU baseline of bench.py
 * (cpu_baseline.kind = "port", and `bench.py --impl reference`) and checked against the golden
 * fixtures by tests/test_oracle_cport.py.  Never linked into the product.
 *
 * Each function cites the reference lines it restates (relative to /root/reference).
 * Build: gcc -O3 -march=x86-64-v3 -fopenmp -ffp-contract=off (no FMA contraction, so the
 * arithmetic follows the reference's NumPy/TF rounding; libm's expf/tanhf differ from NumPy's
 * SIMD kernels by <= 1-2 ulp).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int reflecti(int v, int n) { return v < 0 ? -v : (v >= n ? 2 * n - 2 - v : v); }

int fib_cpu_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ionic.py:44-60,70-81,107-113: boundary-enforced, REFLECT-padded 9-point Laplacian (+ phase term)
 * of cell (r,c) read straight from the raw plane X through clamped indices. */
static inline float lap_at(const float* X, const float* P, int H, int W, int r, int c) {
  const int rn = clampi(r - 1, 1, H - 2), rc = clampi(r, 1, H - 2), rs = clampi(r + 1, 1, H - 2);
  const int cw = clampi(c - 1, 1, W - 2), cc = clampi(c, 1, W - 2), ce = clampi(c + 1, 1, W - 2);
  const float N = X[(size_t)rn * W + cc], S = X[(size_t)rs * W + cc];
  const float Wv = X[(size_t)rc * W + cw], E = X[(size_t)rc * W + ce];
  const float NW = X[(size_t)rn * W + cw], SW = X[(size_t)rs * W + cw];
  const float NE = X[(size_t)rn * W + ce], SE = X[(size_t)rs * W + ce];
  const float C = X[(size_t)rc * W + cc];
  float lap = (((N + S) + Wv) + E) + 0.5f * (((NW + SW) + NE) + SE) - 6.0f * C;
  if (P) {
    const float pN = P[(size_t)reflecti(r - 1, H) * W + c], pS = P[(size_t)reflecti(r + 1, H) * W + c];
    const float pW = P[(size_t)r * W + reflecti(c - 1, W)], pE = P[(size_t)r * W + reflecti(c + 1, W)];
    lap += ((S - N) * (pS - pN) + (E - Wv) * (pE - pW)) / (4.0f * P[(size_t)r * W + c]);
  }
  return lap;
}

static inline float x0_at(const float* X, int H, int W, int r, int c) {
  return X[(size_t)clampi(r, 1, H - 2) * W + clampi(c, 1, W - 2)];
}

/* tf.clip_by_value on the reference's GPU target: fmax(fmin(x, hi), lo), the non-NaN operand wins */
static inline float clipf(float x, float lo, float hi) { return fmaxf(fminf(x, hi), lo); }

/* ionic.py:115-123 */
static inline float rush_larsen(float g, float g_inf, float tau, float neg_dt) {
  return clipf(g + (g - g_inf) * expm1f(neg_dt / tau), 0.00001f, 0.99999f);
}

/* ---- Fenton 4v: fenton.py:46-108.  One explicit Euler step; Uout must not alias U. -------- */
void fib_cpu_fenton_step(int H, int W, const float* U, float* Uout, float* V, float* Wg, float* S,
                         const float* phase, double dt_d, double diff_d) {
  const float dt = (float)dt_d, ddt = (float)(diff_d * dt_d);
  const float tau_vp = 3.33f, tau_vn = 19.2f, tau_wp = 160.0f, tau_wn1 = 75.0f, tau_wn2 = 75.0f;
  const float tau_d = 0.065f, tau_si = 31.8364f, tau_so = 31.8364f, tau_a = 0.009f;
  const float u_c = 0.23f, u_w = 0.146f, u_0 = 0.0f, u_m = 1.0f, u_csi = 0.8f, u_so = 0.3f;
  const float r_sn = 1.2f, k_ = 3.0f, b_so = 0.84f, c_so = 0.02f;
  const float c1 = (float)(0.5 * (0.115 - 0.009)), rdiff = (float)(0.02 - 1.2);
#pragma omp parallel for schedule(static)
  for (int r = 0; r < H; ++r)
    for (int c = 0; c < W; ++c) {
      const size_t i = (size_t)r * W + c;
      const float u = U[i], v = V[i], w = Wg[i], s = S[i];
      const float Hc = u > u_c ? 1.f : (u < u_c ? 0.f : 0.5f);
      const float Hso = u > u_so ? 1.f : (u < u_so ? 0.f : 0.5f);
      const float Gso = 1.f - Hso;
      const float I_fi = -v * Hc * (u - u_c) * (u_m - u) / tau_d;
      const float I_si = -w * s / tau_si;
      const float I_so = c1 * (1.f + tanhf((u - b_so) / c_so)) + (u - u_0) * Gso / tau_so + Hso * tau_a;
      const float dU = -(I_fi + I_si + I_so);
      const float dV = u > u_c ? -v / tau_vp : (1.f - v) / tau_vn;
      const float dW = u > u_c ? -w / tau_wp : (u > u_w ? (1.f - w) / tau_wn2 : (1.f - w) / tau_wn1);
      const float r_s = rdiff * Hc + r_sn;
      const float dS = r_s * (0.5f * (1.f + tanhf((u - u_csi) * k_)) - s);
      Uout[i] = x0_at(U, H, W, r, c) + dt * dU + ddt * lap_at(U, phase, H, W, r, c);
      V[i] = v + dt * dV;
      Wg[i] = w + dt * dW;
      S[i] = s + dt * dS;
    }
}

/* ---- Beeler-Reuter: br.py:125-332 ------------------------------------------------------ */
static const float kAB[12][7] = {   /* br.py:49-62 */
    {0.0005f, 0.083f, 50.f, 0.f, 0.f, 0.057f, 1.f},  {0.0013f, -0.06f, 20.f, 0.f, 0.f, -0.04f, 1.f},
    {0.f, 0.f, 47.f, -1.f, 47.f, -0.1f, -1.f},       {40.f, -0.056f, 72.f, 0.f, 0.f, 0.f, 0.f},
    {0.126f, -.25f, 77.f, 0.f, 0.f, 0.f, 0.f},       {1.7f, 0.f, 22.5f, 0.f, 0.f, -0.082f, 1.f},
    {0.055f, -.25f, 78.f, 0.f, 0.f, -0.2f, 1.f},     {0.3f, 0.f, 32.f, 0.f, 0.f, -0.1f, 1.f},
    {(float)(2 * 0.095), -0.01f, -5.f, 0.f, 0.f, -0.072f, 1.f},
    {(float)(2 * 0.07), -0.017f, 44.f, 0.f, 0.f, 0.05f, 1.f},
    {(float)(2 * 0.012), -0.008f, 28.f, 0.f, 0.f, 0.15f, 1.f},
    {(float)(2 * 0.0065), -0.02f, 30.f, 0.f, 0.f, -0.2f, 1.f}};

static inline float br_rate(float v, const float* c) {   /* br.py:255-264 */
  if (c[3] == 0.f) return (c[0] * expf(c[1] * (v + c[2]))) / (expf(c[5] * (v + c[2])) + c[6]);
  return (c[0] * expf(c[1] * (v + c[2])) + c[3] * (v + c[4])) / (expf(c[5] * (v + c[2])) + c[6]);
}

/* state planes: V C M H J D F XI; gate order of cheb rows: xi m h j d f (inf, tau interleaved).
 * n = slow-gate multiplier of this step (br.py:96-107): 1, or 5/0 under skip. */
void fib_cpu_br_step(int H, int W, const float* V, float* Vout, float* C, float* M, float* Hg,
                     float* J, float* D, float* F, float* XI, const float* phase, double dt_d,
                     double diff_d, int n, int cheby, const float* cheb /*[12][9]*/) {
  const float dt = (float)dt_d, ddt = (float)(diff_d * dt_d);
  const float ndt = (float)(-dt_d), ndts = (float)(-(dt_d * n));
  float* gate[6] = {XI, M, Hg, J, D, F};
#pragma omp parallel for schedule(static)
  for (int r = 0; r < H; ++r)
    for (int c = 0; c < W; ++c) {
      const size_t i = (size_t)r * W + c;
      const float v0 = x0_at(V, H, W, r, c);
      const float m = M[i], h = Hg[i], j = J[i], d = D[i], f = F[i], xi = XI[i], ca = C[i];
      float inf[6], tau[6];
      if (cheby) {
        const float x = (v0 - -30.0f) / 60.0f;
        float S[9];
        S[0] = 1.f; S[1] = x;
        for (int k = 2; k < 9; ++k) S[k] = (2.f * x) * S[k - 1];
        for (int g = 0; g < 6; ++g) {
          if (!(g == 1 || g == 2 || n > 0)) continue;
          for (int q = 0; q < 2; ++q) {
            const float* dd = cheb + (2 * g + q) * 9;
            float acc = dd[0] + dd[1] * S[1];
            for (int k = 2; k < 9; ++k) acc = acc + dd[k] * S[k];
            if (q == 0) inf[g] = acc; else tau[g] = acc;
          }
        }
      } else {
        for (int g = 0; g < 6; ++g) {
          if (!(g == 1 || g == 2 || n > 0)) continue;
          const float a = br_rate(v0, kAB[2 * g]), b = br_rate(v0, kAB[2 * g + 1]);
          inf[g] = a / (a + b);
          tau[g] = 1.0f / (a + b);
        }
      }
      for (int g = 0; g < 6; ++g) {
        const int fast = (g == 1 || g == 2);
        if (fast || n > 0) gate[g][i] = rush_larsen(gate[g][i], inf[g], tau[g], fast ? ndt : ndts);
      }
      const float iK1 = 0.35f * (4.f * (expf(0.04f * (v0 + 85.f)) - 1.f) /
                                     (expf(0.08f * (v0 + 53.f)) + expf(0.04f * (v0 + 53.f))) +
                                 0.2f * ((v0 + 23.0f) / (1.0f - expf(-0.04f * (v0 + 23.f)))));
      const float ix1 = xi * 0.8f * (expf(0.04f * (v0 + 77.f)) - 1.f) / expf(0.04f * (v0 + 35.f));
      const float iNa = (4.0f * m * m * m * h * j + 0.005f) * (v0 - 50.0f);
      const float ECa = -82.3f - 13.0278f * logf(ca);
      const float iCa = 0.09f * d * f * (v0 - ECa);
      const float I_sum = iK1 + ix1 + iNa + iCa;
      Vout[i] = clipf(v0 + ddt * lap_at(V, phase, H, W, r, c) - dt * I_sum / 1.0f, -85.0f, 25.0f);
      C[i] = ca + dt * (-1.0e-7f * iCa + 0.07f * (1.0e-7f - ca));
    }
}
