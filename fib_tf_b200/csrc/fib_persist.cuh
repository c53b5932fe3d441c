// fib_persist.cuh -- the persistent on-chip kernel for the latency regime (the reference's own
// 512 x 512 configurations, BASELINE configs 1 and 2): ONE launch advances a whole run() iteration
// (10 time steps of 4v, 5 of Beeler-Reuter; fenton.py:133-138, br.py:96-107) with the state resident
// on chip between the steps.
//
// Why: a 512^2 state is 4-8 MiB.  One step per launch leaves such a grid launch-latency bound (3.1 us
// per 4v step against 1.3 us of HBM time and ~0.7 us of issue time, profiles/r1_suite.json); k-step
// temporal blocking by halo recomputation would double the arithmetic at this tile size.  Here
// nothing is recomputed and nothing goes through HBM between steps:
//
//   * the grid is cut into row tiles of TH rows x W columns (W <= 512), one tile per CTA, one CTA per SM
//     (cooperative launch: all CTAs are co-resident by construction);
//   * TMA 2-D tile loads (cp.async.bulk.tensor, mbarrier completion) bring the tile into shared memory at
//     the start; the non-diffusing planes then live in REGISTERS (a thread owns one column of its tile)
//     and the diffusing variable in a double-buffered shared-memory tile; TMA tile stores write
//     everything back once at the end;
//   * per step only the two edge rows of the diffusing variable leave the SM, as self-validating
//     8-byte words {value, step number} in a small L2-resident mailbox (the "LL" protocol of NCCL): an
//     aligned 8-byte store is one transaction, so the neighbour simply polls the words it needs until
//     they carry the step it waits for -- no flag, no memory fence, no L1 invalidation on either side.
//     A step first advances the interior rows, which need nothing from outside, and only then the
//     edge rows, so the hand-shake (one L2 store + one L2 load) is hidden behind arithmetic.
//     One __syncthreads per step (the double-buffered shared tile).
//
// Every cell goes through the SAME cell function and the same Laplacian / phase-term code as
// step_kernel (fib_kernels.cuh), with the same clamped index map, and the library is built with
// -fmad=false, so the result is BIT-IDENTICAL to one launch per step (tests/test_gpu_persist.py).
//
// Deadlock safety: inter-CTA waits exist only under a cooperative launch (fib_capi.cu refuses the
// path otherwise); every spin loop is bounded (kSpinLimit polls, ~1 s) and raises *err instead of
// hanging the device.
#pragma once
#include <cuda.h>

#include "model_br.cuh"
#include "model_fenton.cuh"

namespace fib {

constexpr int kPersistThreads = 512;      // one thread per column: W <= 512
constexpr int kPersistBox = 256;          // TMA box width (elements; the hardware limit per dimension)
constexpr int kPersistWP = 512 + 64;      // padded row of the diffusing tile: data at +32 floats (128 B)
constexpr unsigned kSpinLimit = 1u << 22;
// mailbox layout: [2 step parities][tiles][2 sides: 0 = top row, 1 = bottom row][kPersistThreads] words
__host__ __device__ constexpr size_t persist_mailbox_words(int tiles) { return (size_t)2 * tiles * 2 * kPersistThreads; }

// ---- PTX wrappers: mbarrier + TMA (SASS: SYNCS / UTMALDG / UTMASTG) ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 2-D tile load global -> shared, completion counted in bytes on `bar`; (x, y) = (column, row)
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}
// 2-D tile store shared -> global (rows / columns outside the tensor are clipped)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int x, int y, const void* src_smem) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src_smem)), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// one mailbox word: {fp32 value, step number} moved as ONE aligned 8-byte transaction
__device__ __forceinline__ void ll_store(unsigned long long* p, float v, unsigned step) {
  asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(step) : "memory");
}
__device__ __forceinline__ void ll_load(const unsigned long long* p, float& v, unsigned& step) {
  unsigned a, b;
  asm volatile("ld.relaxed.gpu.global.v2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "l"(p) : "memory");
  v = __uint_as_float(a);
  step = b;
}
// ---- arguments --------------------------------------------------------------------------------
constexpr int kPersistMaxSteps = 10;

template <int NS>
struct PersistMaps {
  CUtensorMap x[2];     // the diffusing variable's two buffers, owned rows, box {256, 1}
  CUtensorMap s[NS];    // the other planes, box {256, TH}
};

template <class MS, class MF>
struct PersistArgs {
  float* x[2];                 // the same two buffers as plain pointers (halo layout: row g at (g + 1) * pitch)
  int cur;                     // x[cur] holds the state at the start of the launch
  int nsteps;                  // time steps of this launch (dt_per_step)
  unsigned slow_mask;          // bit s: step s is an MS step (e.g. BR n > 0), else MF (BR n == 0)
  unsigned long long* mail;    // the mailbox (persist_mailbox_words), words {value, step number}
  unsigned base;               // step number of the state at the start of this launch (monotonic over the run)
  int* err;                    // set to 1 if a neighbour wait ran into the spin limit
  const float* phase;          // halo layout (one row), or nullptr
  const unsigned char* pmask;  // [rows][pmask_pitch] as in StepArgs
  int pmask_pitch;
  int bw;                      // TMA box width actually encoded (min(256, W rounded up to 4))
  typename MS::Params ps;      // parameters of the MS steps
  typename MF::Params pf;      // parameters of the MF steps
};

// ---- the kernel ----------------------------------------------------------------------------------
// MS / MF: the cell types of the "slow" and "fast" steps of a schedule (Fenton4v twice; BeelerReuter<C,true>
// and <C,false> for br.py's skip schedule).  TH: rows per tile.  PHASE: phase field present.
template <class MS, class MF, int TH, bool PHASE>
__global__ void __launch_bounds__(kPersistThreads, 1)
persist_kernel(const __grid_constant__ PersistMaps<MS::NS> maps, const Geom g, const PersistArgs<MS, MF> a) {
  constexpr int NS = MS::NS;
  static_assert(MS::NS == MF::NS, "slow / fast steps share the state layout");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* ub = reinterpret_cast<float*>(smem_raw);                         // [2][TH + 2][kPersistWP]
  float* st = ub + 2 * (TH + 2) * kPersistWP;                             // [NS][2 column blocks][TH rows of bw][.. 256]
  uint64_t* bar = reinterpret_cast<uint64_t*>(st + NS * TH * 512);

  const int t = threadIdx.x, tile = blockIdx.x, ntiles = gridDim.x;
  const int W = g.W, H = g.H, pitch = g.pitch;
  const int r0 = tile * TH;
  const int nrows = min(TH, H - r0);
  const int c = t;
  const bool active = c < W;
  const int nblk = (W + kPersistBox - 1) / kPersistBox;
  auto urow = [&](int p, int li) __attribute__((always_inline)) { return ub + (p * (TH + 2) + li) * kPersistWP + 32; };   // column 0 of local row li

  // ---- TMA in: the diffusing tile row by row into the padded buffer, the other planes as {bw, TH} boxes
  if (t == 0) {
    mbar_init(bar, 1);
    fence_async_smem();
  }
  __syncthreads();
  if (t == 0) {
    mbar_expect_tx(bar, (uint32_t)((TH + NS * TH) * nblk * a.bw * sizeof(float)));
    for (int b = 0; b < nblk; ++b) {
      for (int i = 0; i < TH; ++i) tma_load_2d(urow(0, i + 1) + b * kPersistBox, &maps.x[a.cur], b * kPersistBox, r0 + i, bar);
      for (int k = 0; k < NS; ++k) tma_load_2d(st + (k * 2 + b) * TH * kPersistBox, &maps.s[k], b * kPersistBox, r0, bar);
    }
  }
  // meanwhile: which of my cells have a non-trivial phase term (same flags as step_kernel)
  unsigned phbits = 0;
  if (PHASE && active) {
#pragma unroll
    for (int i = 0; i < TH; ++i)
      if (i < nrows && a.pmask[(r0 + i) * a.pmask_pitch + (c >> 5)]) phbits |= 1u << i;
  }
  const int ccl = clampi(c - 1, 1, W - 2), ccc = clampi(c, 1, W - 2), ccr = clampi(c + 1, 1, W - 2);
  while (!mbar_try_wait(bar, 0)) {}

  float s[NS][TH];
  {
    const int b = c / kPersistBox, cb = c - b * kPersistBox;
#pragma unroll
    for (int k = 0; k < NS; ++k)
#pragma unroll
      for (int i = 0; i < TH; ++i) s[k][i] = active ? st[(k * 2 + b) * TH * kPersistBox + i * a.bw + cb] : 0.f;
  }

  // the cell functions only look at `.p`: hand them the parameter blocks where they are (kernel
  // parameter space) instead of copies
  struct PS { const typename MS::Params& p; };
  struct PF { const typename MF::Params& p; };
  const PS sa_s{a.ps};
  const PF sa_f{a.pf};

  // Laplacian (+ phase term) of row i: exactly step_kernel's arithmetic.  nN / nC / nS: the clamped-column
  // triples of the enforced rows above / at / below.
  auto lap_row = [&](int i, const float (&nN)[3], const float (&nC)[3], const float (&nS)[3]) -> float {
    float lap = lap9(nN[1], nS[1], nC[0], nC[2], nN[0], nS[0], nN[2], nS[2], nC[1]);
    if (PHASE && ((phbits >> i) & 1u)) {
      const int gr = r0 + i;
      const float* ph = a.phase;
      const int rN = (reflecti(gr - 1, H) + 1) * pitch, rC = (gr + 1) * pitch, rS = (reflecti(gr + 1, H) + 1) * pitch;
      const float pN = ph[rN + c], pS = ph[rS + c];
      const float pW = ph[rC + clampi(reflecti(c - 1, W), 0, W - 1)], pE = ph[rC + clampi(reflecti(c + 1, W), 0, W - 1)];
      lap = __fadd_rn(lap, phase_term(nN[1], nS[1], nC[0], nC[2], pN, pS, pW, pE, ph[rC + c]));
    }
    return lap;
  };
  // one cell (row i) / two cells (rows i and j of my column as ONE packed pair, fib_math.cuh) through the
  // model's cell function; raw: the un-enforced centre value, x0: the enforced one
  auto advance = [&](bool slow, int i, float raw, float x0, float lap) __attribute__((always_inline)) -> float {
    float sl[NS], xnew;
#pragma unroll
    for (int k = 0; k < NS; ++k) sl[k] = s[k][i];
    if (slow) MS::cell(sa_s, raw, x0, lap, sl, xnew);
    else MF::cell(sa_f, raw, x0, lap, sl, xnew);
#pragma unroll
    for (int k = 0; k < NS; ++k) s[k][i] = sl[k];
    return xnew;
  };
  auto advance2 = [&](bool slow, int i, int j, f2 raw, f2 x0, f2 lap) __attribute__((always_inline)) -> f2 {
    f2 sl[NS], xnew;
#pragma unroll
    for (int k = 0; k < NS; ++k) sl[k] = f2(s[k][i], s[k][j]);
    if (slow) MS::cell(sa_s, raw, x0, lap, sl, xnew);
    else MF::cell(sa_f, raw, x0, lap, sl, xnew);
#pragma unroll
    for (int k = 0; k < NS; ++k) { s[k][i] = sl[k].x; s[k][j] = sl[k].y; }
    return xnew;
  };

  for (int step = 0; step < a.nsteps; ++step) {
    const int p = step & 1;                                  // shared buffer holding the state at this step
    const float* xg = a.x[(a.cur + step) & 1];               // global plane with the neighbours' edge rows of it
    const bool slow = (a.slow_mask >> step) & 1u;
    const unsigned want = a.base + step;                     // step number of the state this step reads
    // neighbour rows (the ring): global rows r0 - 1 and r0 + TH, columns c-1, c, c+1 (clamped)
    float rt[3] = {0.f, 0.f, 0.f}, rb[3] = {0.f, 0.f, 0.f};
    const bool need_top = tile > 0, need_bot = tile + 1 < ntiles;
    // mailbox words of step number n written by `tl`'s side sd (0 = its top row, 1 = its bottom row)
    auto box = [&](unsigned n, int tl, int sd) __attribute__((always_inline)) {
      return a.mail + (((size_t)(n & 1u) * ntiles + tl) * 2 + sd) * kPersistThreads;
    };
    if (step == 0 && active) {
      // the state at the start of a launch is complete in the global plane (previous launch / upload /
      // stimulus), whatever the mailbox holds
      if (need_top) {
        const float* q = xg + (size_t)(r0 - 1 + 1) * pitch;
        rt[0] = __ldcg(q + ccl); rt[1] = __ldcg(q + ccc); rt[2] = __ldcg(q + ccr);
      }
      if (need_bot) {
        const float* q = xg + (size_t)(r0 + TH + 1) * pitch;
        rb[0] = __ldcg(q + ccl); rb[1] = __ldcg(q + ccc); rb[2] = __ldcg(q + ccr);
      }
    }
    // My six mailbox words (three columns on each side).  All six loads are issued back to back, so one
    // poll costs ONE L2 round trip; the first poll is issued BEFORE the interior rows are advanced and only
    // looked at afterwards, which hides that round trip behind arithmetic when the neighbours are on time.
    float mv[6];
    unsigned mn[6];
    const int mcols[3] = {ccl, ccc, ccr};
    auto ring_poll = [&]() __attribute__((always_inline)) {
#pragma unroll
      for (int sd = 0; sd < 2; ++sd) {
        // the tile above publishes its BOTTOM row (side 1), the tile below its TOP row (side 0)
        const bool need = sd == 0 ? need_top : need_bot;
        const unsigned long long* w = sd == 0 ? box(want, tile - 1, 1) : box(want, tile + 1, 0);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          if (need) ll_load(w + mcols[q], mv[sd * 3 + q], mn[sd * 3 + q]);
          else { mv[sd * 3 + q] = 0.f; mn[sd * 3 + q] = want; }
        }
      }
    };
    auto ring_wait = [&]() __attribute__((always_inline)) {          // bounded: raises *err instead of hanging
      if (step == 0 || !active) return;
      unsigned spins = 0;
      for (;;) {
        bool ok = true;
#pragma unroll
        for (int q = 0; q < 6; ++q) ok &= mn[q] == want;
        if (ok) break;
        if (++spins > kSpinLimit) { *a.err = 1; break; }
        ring_poll();
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) { rt[q] = mv[q]; rb[q] = mv[3 + q]; }
    };
    if (step > 0 && active) ring_poll();

    // triple of enforced values of global row gr (already clamped by the caller) at my three columns
    auto triple = [&](int gr, float (&v)[3]) {
      const int li = gr - r0 + 1;
      if (li == 0) { v[0] = rt[0]; v[1] = rt[1]; v[2] = rt[2]; }
      else if (li == TH + 1) { v[0] = rb[0]; v[1] = rb[1]; v[2] = rb[2]; }
      else {
        const float* q = urow(p, li);
        v[0] = q[ccl]; v[1] = q[ccc]; v[2] = q[ccr];
      }
    };
    auto row_in = [&](int i, float (&nN)[3], float (&nC)[3], float (&nS)[3]) {
      const int gr = r0 + i;
      triple(clampi(gr - 1, 1, H - 2), nN);
      triple(clampi(gr, 1, H - 2), nC);
      triple(clampi(gr + 1, 1, H - 2), nS);
    };
    auto do_row = [&](int i) __attribute__((always_inline)) {
      float nN[3], nC[3], nS[3];
      row_in(i, nN, nC, nS);
      const float xnew = advance(slow, i, urow(p, i + 1)[c], nC[1], lap_row(i, nN, nC, nS));
      urow(p ^ 1, i + 1)[c] = xnew;
      return xnew;
    };
    // rows i < j as one packed pair; row j may lie beyond the grid in the last tile (junk lane, not stored)
    auto do_pair = [&](int i, int j) __attribute__((always_inline)) {
      float aN[3], aC[3], aS[3], bN[3], bC[3], bS[3];
      row_in(i, aN, aC, aS);
      const bool jok = j < nrows;
      if (jok) row_in(j, bN, bC, bS);
      else {
#pragma unroll
        for (int q = 0; q < 3; ++q) bN[q] = bC[q] = bS[q] = 0.f;
      }
      const float li = lap_row(i, aN, aC, aS), lj = jok ? lap_row(j, bN, bC, bS) : 0.f;
      const f2 xnew = advance2(slow, i, j, f2(urow(p, i + 1)[c], jok ? urow(p, j + 1)[c] : 0.f), f2(aC[1], bC[1]),
                               f2(li, lj));
      urow(p ^ 1, i + 1)[c] = xnew.x;
      if (jok) urow(p ^ 1, j + 1)[c] = xnew.y;
      return xnew;
    };
    constexpr bool kPairs = MS::PACKED && MF::PACKED && TH >= 2;

    // interior rows: need nothing from outside the tile
    if (active) {
      if constexpr (kPairs) {
#pragma unroll
        for (int i = 1; i + 1 < TH - 1; i += 2)
          if (i < nrows) do_pair(i, i + 1);
      } else {
#pragma unroll
        for (int i = 1; i < TH - 1; ++i)
          if (i < nrows) do_row(i);
      }
    }
    ring_wait();
    // edge rows, published to the neighbours straight from registers
    if (active) {
      float top, bot = 0.f;
      if constexpr (kPairs) {
        const f2 e = do_pair(0, TH - 1);
        top = e.x;
        bot = e.y;
      } else {
        top = do_row(0);
        if (TH > 1 && TH - 1 < nrows) bot = do_row(TH - 1);
      }
      // (the last step's rows are not consumed through the mailbox: the next launch starts from the plane)
      if (need_top) ll_store(box(want + 1, tile, 0) + c, top, want + 1);
      if (need_bot) ll_store(box(want + 1, tile, 1) + c, TH > 1 ? bot : top, want + 1);
    }
    __syncthreads();            // shared tile of the next step complete, the old one free for reuse
  }

  // ---- TMA out: registers -> staging, then tile stores (rows beyond the grid are clipped)
  if (active) {
    const int b = c / kPersistBox, cb = c - b * kPersistBox;
#pragma unroll
    for (int k = 0; k < NS; ++k)
#pragma unroll
      for (int i = 0; i < TH; ++i) st[(k * 2 + b) * TH * kPersistBox + i * a.bw + cb] = s[k][i];
  }
  fence_async_smem();
  __syncthreads();
  if (t == 0) {
    const int pf = a.nsteps & 1;
    const CUtensorMap* mx = &maps.x[(a.cur + a.nsteps) & 1];
    for (int b = 0; b < nblk; ++b) {
      for (int i = 0; i < TH; ++i) tma_store_2d(mx, b * kPersistBox, r0 + i, urow(pf, i + 1) + b * kPersistBox);
      for (int k = 0; k < NS; ++k)
        if (MS::stores(k) || MF::stores(k)) tma_store_2d(&maps.s[k], b * kPersistBox, r0, st + (k * 2 + b) * TH * kPersistBox);
    }
    tma_store_commit_wait();
  }
}

template <int NS, int TH>
constexpr size_t persist_smem_bytes() {
  return (size_t)(2 * (TH + 2) * kPersistWP + NS * TH * 512) * sizeof(float) + 64;
}

}  // namespace fib
