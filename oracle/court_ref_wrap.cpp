// TEST INFRASTRUCTURE (see oracle/README.md).  extern "C" wrapper around the reference's OWN
// courtemanche.h, which is #included from /root/reference at build time (nothing is copied into
// this repository).  Same host-compilation trick as the reference's generate_table.cpp:4-9:
// neutralise the CUDA qualifiers and supply float3.  Output: oracle/_ref/libcourt_ref.so.
#include <cmath>
#include <cstdio>

#define __device__
#define __host__
struct float3 { float x, y, z; };

#include "ionic.h"
#include "courtemanche.h"

extern "C" {
// courtemanche.h:159-285
void ref_calc_inter(float V, float* inter30) { calc_inter(V, inter30); }
// courtemanche.h:473-479  (table[150][30])
void ref_init_table(float* table) { init_table<Courtemanche>(table); }
// courtemanche.h:57-103   (state[0] = V, state[1..20] = w[0..19])
void ref_init_cell(float* state21, int stim) { init_cell<Courtemanche>(state21, state21 + 1, stim); }
// courtemanche.h:294-440  (rate[21] from state[21], LUT lookup inside)
void ref_deriv(float* state21, float* rate21, float dt, const float* table, int chronic) {
  Config cfg{};
  cfg.dt = dt;
  cfg.table = table;
  cfg.chronic = chronic != 0;
  deriv<Courtemanche>(state21, rate21, cfg);
}
// `steps` forward-Euler steps state += dt * rate for n independent cells (states[n][21]); the last
// increment dt * rate of every cell is left in incs[n][21].  A loop around the reference's deriv<>.
void ref_euler_batch(float* states, float* incs, long n, int steps, float dt, const float* table, int chronic) {
  Config cfg{};
  cfg.dt = dt;
  cfg.table = table;
  cfg.chronic = chronic != 0;
  for (long i = 0; i < n; ++i) {
    float* st = states + i * 21;
    float rate[21];
    for (int s = 0; s < steps; ++s) {
      deriv<Courtemanche>(st, rate, cfg);
      for (int k = 0; k < 21; ++k) {
        const float inc = dt * rate[k];
        incs[i * 21 + k] = inc;
        st[k] = st[k] + inc;
      }
    }
  }
}
int ref_table_rows() { return Courtemanche::TABLE_ROWS; }
int ref_table_cols() { return Courtemanche::TABLE_COLS; }
}
