// dfx_adjoint3.cuh -- the 24-warp adjoint kernel for lattices of the cfg1 / cfg3 size (same algorithm as
// dfx_adjoint.cuh, see there for the mathematics; same numerics as dfx_adjoint2.cuh).
//
// One CTA of 768 threads per design.  Every thread has two roles:
//   bond role : thread t evaluates bond t (phase B)                                    t < n_bonds <= 768
//   unit role : threads 0..383 ("P") own the primal half of rigid unit t: u, v, the velocity-derivative history,
//               the inertia / damping quadratures; threads 384..767 ("D") own the dual half of unit t-384:
//               lambda_u, lambda_v, their derivative histories, the centroid_node_vectors quadratures, the
//               drive / t0 terms (phases A and C)
// so that phases A and C run on 24 warps instead of 12 and the bond phase needs the registers of ONE dual-number
// bond evaluation (80 per thread).  Storage tiers:
//   tensor memory (tcgen05.ld/st, one lane per thread; P and D of a unit share the lane) : derivative histories,
//                  state at the step start, stage values parked across the bond phase
//   shared memory : stage state (Us, Ws), per-bond result slots, 1/m and damping coefficients, and the running
//                  solution / error sums of the 14 unit-owned quadratures (read-modify-write every stage)
//   L2-resident scratch indexed by SM : bond constants (read only, fetched at the top of an evaluation, a phase
//                  before their use), the bond-owned quadratures, quantities touched once per step (q0, k1, k7)
// The kernel is compiled for a fixed vocabulary (template parameters + preconditions checked by the host:
// ligament energy, scalar stiffness leaves, no external load, n_blocks <= 384, n_bonds <= 768, n_npb == NPB);
// anything else runs dfx_adjoint2.cuh / dfx_adjoint.cuh.
//
// Dense scalar leaves (k_stretch, k_shear, k_rot, the three contact parameters) are not reduced with shuffles in
// every warp: each bond thread stores its partial to the scratch and ONE warp per leaf sums the 768 partials in
// phase C.  Step control lives in shared memory (written by thread 0 between two barriers), not in registers.
#pragma once

#include "dfx_adjoint2.cuh"

namespace dfx {
namespace k3 {

constexpr int TT = 768, TU = 384, NW = TT / 32, NDW = TU / 32;
// tensor-memory slots (doubles) of the two unit roles
// (the last slot of each role holds per-thread topology words that would otherwise occupy registers in the bond phase)
constexpr int P_KV = 0, P_U0 = 21, P_V0 = 24, P_TV = 27, P_TOPO = 30, P_N = 32;
constexpr int D_KLU = 0, D_KLV = 21, D_LU0 = 42, D_LV0 = 45, D_TLU = 48, D_TOPO = 51, D_N = 53;
constexpr int P_COLS = 2 * P_N, D_COLS = 2 * D_N;  // 64 + 106 columns per (P, D) warp pair, three pairs per lane quarter
static_assert(3 * (P_COLS + D_COLS) <= 512, "tensor memory columns");
constexpr int NCU = 40;   // constrained units whose drive vectors are cached in shared memory
constexpr int NDRV = 14;  // one row of the drive table: s[2], ds/dt[2], ds/dparam[2][5]
constexpr int NBC = 10;  // bond constants: r0x r0y L0 1/L0 r1x r1y r2x r2y da1 da2
constexpr int NE3 = 12;  // quadrature entries per thread: 0..9 unit role (P: inertia[3] damping[3] cnv x[4]; D: cnv y[4]), 10..11 bond role
// L2 scratch of one SM (doubles); every array is [..][TT], a thread touches its own column only
constexpr long long G_BC = 0;
constexpr long long G_SPART = G_BC + (long long)NBC * TT;   // [2 (parity of the evaluation)][6][TT] per-bond partials of k_stretch k_shear k_rot c_min c_cut k_c
constexpr long long G_BQ = G_SPART + 12LL * TT;              // [sol, err][2][TT] running sums of the reference-vector quadratures
constexpr long long G_Q = G_BQ + 4LL * TT;                  // [q0 a, q0 b, k1 a, k1 b][NE3][TT]
constexpr long long G_TOTAL = G_Q + 4LL * NE3 * TT;
// scalar leaves: running sums [k1, k7, sol, err, mid][NSLOT]
constexpr int SLOT_K = 0;                                    // k_stretch k_shear k_rot c_min c_cut k_c: four partials each (see phase C)
constexpr int SLOT_T = SLOT_K + 6 * 4;                       // t0_bar, drive[5]: one partial per D warp
constexpr int SLOT_DAMP = SLOT_T + 6 * NDW;                  // scalar damping leaf: one partial per P warp
constexpr int NSLOT = SLOT_DAMP + NDW;

struct Ctrl {
  double h, h0, d1, s0, s_target, s_cur, x;
  double hst;  // step size of the evaluation in progress (h0 for the probe)
  long long n_steps, n_acc, n_rhs, istep;
  int status, crossing, contact_seen;
  int i, par;  // output interval in progress; which copy of q0 / k1 is current
  uint32_t tmem_base;
  int ev_w[NW];  // kind of the evaluation in progress, one copy per warp (written by its lane 0): a register carried round
                 // the loop is spilled to local memory = an L2 round trip at the top of every phase
};

template <int NSL>
struct Lay {  // shared-memory layout (doubles)
  static constexpr int RED = 0, US = 40, WS = US + 5 * TU, SL = WS + 6 * TU, INVM = SL + NSL * TT, CD = INVM + 3 * TU,
                       QSP = CD + 3 * TU, QSD = QSP + 20 * TU, SQ = QSD + 8 * TU, ACC = SQ + 2 * NSCAL, DRV = ACC + 5 * NSLOT,
                       PC = DRV + 8 * NDRV /* k_stretch k_shear k_rot c_min c_cut k_c */, CV = PC + 6 /* [NCU][6] drive vectors */,
                       CTRL = CV + 6 * NCU,
                       END = CTRL + (int)((sizeof(Ctrl) + 7) / 8);
};

// wide tensor-memory loads: NC consecutive 32-bit columns of the thread's lane with ONE instruction (x2 .. x32)
template <int NC> __device__ __forceinline__ void tm_ld_cols(uint32_t taddr, uint32_t* r);
template <> __device__ __forceinline__ void tm_ld_cols<2>(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr));
}
template <> __device__ __forceinline__ void tm_ld_cols<4>(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
template <> __device__ __forceinline__ void tm_ld_cols<8>(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
template <> __device__ __forceinline__ void tm_ld_cols<16>(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}
template <> __device__ __forceinline__ void tm_ld_cols<32>(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
}

// N consecutive slots (doubles) with as few load instructions as possible and one wait
template <int N>
__device__ __forceinline__ void tm_ld(uint32_t taddr, double (&out)[N]) {
  uint32_t r[2 * N];
  constexpr int NC = 2 * N;
  int c = 0;
#pragma unroll
  for (int w = 32; w >= 2; w >>= 1) {
#pragma unroll
    for (int rep_ = 0; rep_ < 2; ++rep_) {
      if (NC - c >= w) {
        if (w == 32) tm_ld_cols<32>(taddr + c, r + c);
        else if (w == 16) tm_ld_cols<16>(taddr + c, r + c);
        else if (w == 8) tm_ld_cols<8>(taddr + c, r + c);
        else if (w == 4) tm_ld_cols<4>(taddr + c, r + c);
        else tm_ld_cols<2>(taddr + c, r + c);
        c += w;
      }
    }
  }
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < N; ++i) { tmem_pin(r[2 * i], r[2 * i + 1]); out[i] = __hiloint2double(r[2 * i + 1], r[2 * i]); }
}
template <int N>
__device__ __forceinline__ void tm_lds(uint32_t taddr, int stride, double (&out)[N]) {  // N slots `stride` apart, one wait
  uint32_t lo[N], hi[N];
#pragma unroll
  for (int i = 0; i < N; ++i) tmem_ld_issue(taddr + 2 * i * stride, lo[i], hi[i]);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < N; ++i) { tmem_pin(lo[i], hi[i]); out[i] = __hiloint2double(hi[i], lo[i]); }
}

// stage values from the derivative history: P role (u through the folded coefficients a2, v) and D role.  The history
// stages L0..L1 are fetched with one wait; the tableau entries are compile-time indexed (constant-bank operands).
template <int ST, int L0, int L1>
__device__ __forceinline__ void acc_P(uint32_t ta, const Tableau& tab, double (&au)[3], double (&av)[3]) {
  constexpr int N = 3 * (L1 - L0 + 1);
  double kv[N];
  tm_ld<N>(ta + 2 * (P_KV + 3 * L0), kv);
#pragma unroll
  for (int l = L0; l <= L1; ++l) {
    const double b = tab.beta[ST][l], b2 = tab.a2[ST][l];
#pragma unroll
    for (int j = 0; j < 3; ++j) { au[j] = fma(b2, kv[3 * (l - L0) + j], au[j]); av[j] = fma(b, kv[3 * (l - L0) + j], av[j]); }
  }
}
template <int ST>
__device__ __forceinline__ void stage_P(uint32_t ta, const Tableau& tab, double h, double (&us)[3], double (&vs)[3]) {
  double au[3] = {0, 0, 0}, av[3] = {0, 0, 0};
  if constexpr (ST <= 2) acc_P<ST, 0, ST>(ta, tab, au, av);
  else { acc_P<ST, 0, 2>(ta, tab, au, av); acc_P<ST, 3, ST>(ta, tab, au, av); }
  double y0[6];
  tm_ld<6>(ta + 2 * P_U0, y0);
  const double ha = h * tab.alpha[ST], h2 = h * h;
#pragma unroll
  for (int j = 0; j < 3; ++j) { us[j] = y0[j] - ha * y0[3 + j] - h2 * au[j]; vs[j] = y0[3 + j] + h * av[j]; }
}
template <int ST, int L0, int L1>
__device__ __forceinline__ void acc_D(uint32_t ta, const Tableau& tab, double (&alu)[3], double (&alv)[3]) {
  constexpr int N = 3 * (L1 - L0 + 1);
  double ku[N], kw[N];
  tm_ld<N>(ta + 2 * (D_KLU + 3 * L0), ku);
  tm_ld<N>(ta + 2 * (D_KLV + 3 * L0), kw);
#pragma unroll
  for (int l = L0; l <= L1; ++l) {
    const double b = tab.beta[ST][l];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      alu[j] = fma(b, ku[3 * (l - L0) + j], alu[j]);
      alv[j] = fma(b, kw[3 * (l - L0) + j], alv[j]);
    }
  }
}
template <int ST>
__device__ __forceinline__ void stage_D(uint32_t ta, const Tableau& tab, double h, double (&lus)[3], double (&lvs)[3]) {
  double alu[3] = {0, 0, 0}, alv[3] = {0, 0, 0};
  acc_D<ST, 0, (ST < 1 ? ST : 1)>(ta, tab, alu, alv);
  if constexpr (ST >= 2) acc_D<ST, 2, (ST < 3 ? ST : 3)>(ta, tab, alu, alv);
  if constexpr (ST >= 4) acc_D<ST, 4, ST>(ta, tab, alu, alv);
  double y0[6];
  tm_ld<6>(ta + 2 * D_LU0, y0);
#pragma unroll
  for (int j = 0; j < 3; ++j) { lus[j] = y0[j] + h * alu[j]; lvs[j] = y0[3 + j] + h * alv[j]; }
}

// Kind of an augmented-RHS evaluation, a compile-time tag of the phase code: 0..5 = Runge-Kutta stage ev of a step
// (the derivative lands in history slot ev + 1), 6 = f0 at the start of an output interval (slot 0), 7 = the second
// evaluation of initial_step_size (slot 1).  Quadrature "mode" of the evaluation (as in dfx_adjoint2.cuh):
// 0: k1 at interval start | 1: nothing (zero weights) | 2: first accumulating stage | 3..5: accumulating | 6: last stage |
// 7: probe.
constexpr int EV_INIT = 6, EV_PROBE = 7;
template <int V> using IC = std::integral_constant<int, V>;
template <int EV> struct Ev {
  static constexpr int kidx = EV < 6 ? EV + 1 : (EV == EV_INIT ? 0 : 1);
  static constexpr int mode = EV < 6 ? EV + 1 : (EV == EV_INIT ? 0 : 7);
};
// what the quadratures need from the step controller besides the tableau
struct QC {
  int par;        // which copy of q0 / k1 is current
  bool crossing;  // the step reaches the output time: also integrate the midpoint sum, interpolate at the end
  double h, x, atol, rtol;
};

// N quadrature entries of one thread receive their integrand values.  `qg` = the thread's column of G_Q, entry e0 + k;
// running solution / error sums at sol[k * sstride], err[k * sstride] (shared memory for the unit role, scratch for the
// bond role).  Returns the thread's contribution to the error norm (mode 6) or to d2 of initial_step_size (mode 7).
// CG: the sums live in the scratch and are also the target of red.global.add (bond role): read / written at L2.
template <int MODE, int N, bool CG = false>
__device__ __forceinline__ double quad_entries(const QC& c, const Tableau& tab, double* qg, int e0, double* sol, double* err,
                                               int sstride, const double (&val)[N]) {
  double acc = 0.0;
  auto ldv = [](const double* p) { return CG ? __ldcg(p) : *p; };
  auto stv = [](double* p, double v) { if (CG) __stcg(p, v); else *p = v; };
  constexpr int K = MODE >= 1 && MODE <= 6 ? MODE : 0;
  const double cs = tab.c_sol[K], ce = tab.c_err[K], cm = tab.c_mid[K];
  double* q0p = qg + (long long)(c.par * NE3 + e0) * TT;
  double* qnp = qg + (long long)((1 - c.par) * NE3 + e0) * TT;
  double* k1p = qg + (long long)((2 + c.par) * NE3 + e0) * TT;
  double* k7p = qg + (long long)((3 - c.par) * NE3 + e0) * TT;  // spare copy: k7, or the midpoint sum on a crossing step
  if constexpr (MODE >= 3 && MODE <= 5) {
#pragma unroll
    for (int k0 = 0; k0 < N; k0 += 4) {  // groups of four: loads in flight together, few live registers
      double s_in[4], e_in[4];
#pragma unroll
      for (int k = k0; k < k0 + 4 && k < N; ++k) { s_in[k - k0] = ldv(&sol[k * sstride]); e_in[k - k0] = ldv(&err[k * sstride]); }
#pragma unroll
      for (int k = k0; k < k0 + 4 && k < N; ++k) {
        stv(&sol[k * sstride], fma(cs, val[k], s_in[k - k0]));
        stv(&err[k * sstride], fma(ce, val[k], e_in[k - k0]));
      }
    }
    if (c.crossing) {
#pragma unroll
      for (int k = 0; k < N; ++k) stv(&k7p[k * TT], fma(cm, val[k], ldv(&k7p[k * TT])));
    }
  } else if constexpr (MODE == 2) {
    double k1[N];
#pragma unroll
    for (int k = 0; k < N; ++k) k1[k] = ldv(&k1p[k * TT]);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      stv(&sol[k * sstride], fma(cs, val[k], tab.c_sol[0] * k1[k]));
      stv(&err[k * sstride], fma(ce, val[k], tab.c_err[0] * k1[k]));
      if (c.crossing) stv(&k7p[k * TT], fma(cm, val[k], tab.c_mid[0] * k1[k]));
    }
  } else if constexpr (MODE == 6) {
    double q_in[N];
#pragma unroll
    for (int k = 0; k < N; ++k) q_in[k] = q0p[k * TT];
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const double q1 = fma(c.h, ldv(&sol[k * sstride]), q_in[k]);
      const double r = c.h * fma(ce, val[k], ldv(&err[k * sstride])) * rcp_pos(c.atol + c.rtol * fmax(fabs(q_in[k]), fabs(q1)));
      acc = fma(r, r, acc);
      if (!c.crossing) { qnp[k * TT] = q1; stv(&k7p[k * TT], val[k]); }
      else {
        const double amid = fma(cm, val[k], ldv(&k7p[k * TT]));
        qnp[k * TT] = interp_eval(q_in[k], q1, q_in[k] + c.h * amid, c.h * ldv(&k1p[k * TT]), c.h * val[k], c.x);
      }
    }
  } else if constexpr (MODE == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) stv(&k1p[k * TT], val[k]);
  } else if constexpr (MODE == 7) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const double d = (val[k] - ldv(&k1p[k * TT])) * rcp_pos(c.atol + fabs(q0p[k * TT]) * c.rtol);
      acc = fma(d, d, acc);
    }
  }
  return acc;
}

// one running-sum slot of a scalar leaf receives the total (or a warp's partial) of this evaluation
template <int MODE>
__device__ __forceinline__ void scal_slot(const Tableau& tab, double* accb, int slot, double v) {
  constexpr int K = MODE >= 1 && MODE <= 6 ? MODE : 0;
  double* k1 = accb + slot; double* k7 = k1 + NSLOT; double* sol = k7 + NSLOT; double* err = sol + NSLOT; double* mid = err + NSLOT;
  if constexpr (MODE == 0) *k1 = v;
  else if constexpr (MODE == 7) *k7 = v;
  else if constexpr (MODE == 2) {
    const double a = *k1;
    *sol = tab.c_sol[0] * a + tab.c_sol[K] * v; *err = tab.c_err[0] * a + tab.c_err[K] * v; *mid = tab.c_mid[0] * a + tab.c_mid[K] * v;
  } else if constexpr (MODE == 6) { *err += tab.c_err[K] * v; *mid += tab.c_mid[K] * v; *k7 = v; }
  else if constexpr (MODE >= 3 && MODE <= 5) { *sol += tab.c_sol[K] * v; *err += tab.c_err[K] * v; *mid += tab.c_mid[K] * v; }
}
// a further contribution of the same evaluation to a slot that scal_slot has already been applied to
template <int MODE>
__device__ __forceinline__ void scal_slot_more(const Tableau& tab, double* k1, double v) {
  constexpr int K = MODE >= 1 && MODE <= 6 ? MODE : 0;
  double* k7 = k1 + NSLOT; double* sol = k7 + NSLOT; double* err = sol + NSLOT; double* mid = err + NSLOT;
  if constexpr (MODE == 0) *k1 += v;
  else if constexpr (MODE == 7) *k7 += v;
  else if constexpr (MODE == 6) { *err += tab.c_err[K] * v; *mid += tab.c_mid[K] * v; *k7 += v; }
  else if constexpr (MODE >= 2 && MODE <= 5) { *sol += tab.c_sol[K] * v; *err += tab.c_err[K] * v; *mid += tab.c_mid[K] * v; }
}
// total of scalar leaf `which` (NSCAL numbering of dfx_adjoint.cuh) in running-sum array m (0 k1, 1 k7, 2 sol, 3 err, 4 mid)
__device__ __forceinline__ double scal_total(const double* accb, int m, int which) {
  const double* p = accb + m * NSLOT;
  int s0 = SLOT_DAMP, n = NDW;
  if (which >= SC_KS && which <= SC_KR) { s0 = SLOT_K + (which - SC_KS) * 4; n = 4; }
  else if (which >= SC_CONTACT && which < SC_CONTACT + 3) { s0 = SLOT_K + (3 + which - SC_CONTACT) * 4; n = 4; }
  else if (which == SC_T0) s0 = SLOT_T;
  else if (which >= SC_DRIVE) s0 = SLOT_T + (1 + which - SC_DRIVE) * NDW;
  double s = 0.0;
#pragma unroll 1
  for (int w = 0; w < n; ++w) s += p[s0 + w];
  return s;
}

// the five totals (k1, k7, sol, err, mid) of scalar leaf `which`, computed by one warp: lane l fetches partial l of every
// array, four shuffle steps add them up; valid in lane 0 (fixed order: deterministic)
__device__ __forceinline__ void scal_totals_warp(const double* accb, int which, int lane, double (&out)[5]) {
  int s0 = SLOT_DAMP, n = NDW;
  if (which >= SC_KS && which <= SC_KR) { s0 = SLOT_K + (which - SC_KS) * 4; n = 4; }
  else if (which >= SC_CONTACT && which < SC_CONTACT + 3) { s0 = SLOT_K + (3 + which - SC_CONTACT) * 4; n = 4; }
  else if (which == SC_T0) s0 = SLOT_T;
  else if (which >= SC_DRIVE) s0 = SLOT_T + (1 + which - SC_DRIVE) * NDW;
#pragma unroll
  for (int m = 0; m < 5; ++m) out[m] = lane < n ? accb[m * NSLOT + s0 + lane] : 0.0;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
    for (int m = 0; m < 5; ++m) out[m] += __shfl_xor_sync(0xffffffffu, out[m], o);
  }
}
// bond role, the modes that occur once per step or per interval (0, 2, 6, 7); out of line to keep the bond phase lean
static __device__ __noinline__ double bond_quad_rare_nl(int mode, const QC& c, const Tableau& tab, double* qg, double* bq, double qb0,
                                                        double qb1) {
  const double qb[2] = {qb0, qb1};
  switch (mode) {
    case 0: return quad_entries<0, 2, true>(c, tab, qg, 10, bq, bq + 2 * TT, TT, qb);
    case 2: return quad_entries<2, 2, true>(c, tab, qg, 10, bq, bq + 2 * TT, TT, qb);
    case 6: return quad_entries<6, 2, true>(c, tab, qg, 10, bq, bq + 2 * TT, TT, qb);
    default: return quad_entries<7, 2, true>(c, tab, qg, 10, bq, bq + 2 * TT, TT, qb);
  }
}

// cotangent of ys[design][i][(is_v ? n_free : 0) + f] (see cotangent_nl) with the objective's target index of f known
// (kf = index + 1, 0 = not a target, 1023 = not cached: look it up)
static __device__ __noinline__ double cotangent_k_nl(const AdjArgs& a, int design, int i, int f, bool is_v, int kf) {
  if (a.g) return __ldcs(&a.g[((long long)design * a.n_t + i) * 2 * a.topo.n_free + (is_v ? a.topo.n_free : 0) + f]);
  return objective_cotangent_k(a, design, i, f, is_v, kf == 1023 ? objective_target_index(a, f) : kf - 1);
}
static __device__ __noinline__ int target_index_nl(const AdjArgs& a, int f) { return objective_target_index(a, f); }

// one row of the drive table (out of line: several call sites, none of them inside the RHS phases)
static __device__ __noinline__ void drive_row_nl(int kind, double t, const double* g_drive, DriveTable table, double* r) {
  if (kind == DFX_DRIVE_PULSE || kind == DFX_DRIVE_HARMONIC) {
    // one channel, three parameters (amplitude, loading rate, delay): no DriveEval round trip through local memory
    // (this row is on the critical path of every step: the whole CTA waits for it at the barrier that ends the step)
    double s, dtau, dA, df;
    pulse_eval(t - g_drive[2], g_drive[0], g_drive[1], kind == DFX_DRIVE_PULSE, true, s, dtau, dA, df);
    r[0] = s; r[1] = 0.0; r[2] = dtau; r[3] = 0.0;
    r[4] = dA; r[5] = df; r[6] = -dtau; r[7] = 0.0; r[8] = 0.0;
#pragma unroll
    for (int q = 0; q < DFX_MAX_DRIVE_PARAMS; ++q) r[9 + q] = 0.0;
    return;
  }
  DriveEval de;
  drive_eval(kind, t, g_drive, true, de, table);
  r[0] = de.s[0]; r[1] = de.s[1]; r[2] = de.sdot[0]; r[3] = de.sdot[1];
#pragma unroll
  for (int q = 0; q < DFX_MAX_DRIVE_PARAMS; ++q) { r[4 + q] = de.dsdp[0][q]; r[9 + q] = de.dsdp[1][q]; }
}

}  // namespace k3

struct Adj3Args {
  AdjArgs a;
  const int* node_bond;         // [n_nodes] bond*2+side of the bond attached to the node, or -1
  long long scratch_per_slot;   // doubles of scratch per SM (>= k3::G_TOTAL)
  int scratch_slots;            // slices available; the kernel indexes them by %smid
};

// DAMP: 0 no damping leaf, 1 scalar leaf, 2 (n_damped, 3) leaf
template <int NPB, bool CONTACT, int DAMP>
__global__ void __launch_bounds__(k3::TT, 1) adjoint3_kernel(const __grid_constant__ Adj3Args A) {
  using namespace k3;
  extern __shared__ double smem[];
  constexpr int NSL = CONTACT ? 14 : 12;
  using L = Lay<NSL>;
  const AdjArgs& a = A.a;
  const DevTopo& T = a.topo;
  const Tableau& tab = a.tab;
  const int design = a.order ? a.order[blockIdx.x] : (int)blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool isD = tid >= TU;
  const int NB = T.n_blocks, NBONDS = T.n_bonds, nf = T.n_free;

  double* red = smem + L::RED;
  double* Us = smem + L::US;      // [5][TU]  x y theta sin cos
  double* Ws = smem + L::WS;      // [2][3][TU]  w = lambda_v / m, double buffered (read by P in phase C, written by D in the next phase A)
  double* SL = smem + L::SL;      // [NSL][TT] per bond: gdx gdy T1 T2 | hx hy H1 H2 | g1x g1y g2x g2y | a1 a2
  double* INVM = smem + L::INVM;  // [3][TU]
  double* CDs = smem + L::CD;     // [3][TU]
  double* Sq0 = smem + L::SQ;
  double* Sqnew = Sq0 + NSCAL;
  double* accb = smem + L::ACC;   // [5][NSLOT]
  double* drv = smem + L::DRV;
  double* PC = smem + L::PC;      // k_stretch k_shear k_rot c_min c_cut k_c
  double* CV = smem + L::CV;      // [NCU][6] drive vectors (vec0[3], vec1[3]) of the constrained units
  Ctrl* C = (Ctrl*)(smem + L::CTRL);

  unsigned smid;
  asm("mov.u32 %0, %%smid;" : "=r"(smid));
  if ((int)smid >= A.scratch_slots) {  // cannot happen on a part whose %nsmid the host sized the scratch for; fail loudly
    if (tid == 0 && a.stats) {
      DfxStats st; st.steps = 0; st.accepted = 0; st.rhs_evals = 0; st.status = DFX_STATUS_NONFINITE; st.reserved = 0; st.last_dt = 0.0;
      a.stats[design] = st;
    }
    return;
  }
  double* gbase = a.scratch + (long long)smid * A.scratch_per_slot;
  double* gcol = gbase + tid;                // + array offset + row * TT
  double* qg = gcol + G_Q;

  // ---- tensor memory: all 512 columns ---------------------------------------------------------------------
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&C->tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = C->tmem_base;
  // lane quarter of the warp; column base: the three P warps of a quarter, then its three D warps
  const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) +
                      (uint32_t)(isD ? 3 * P_COLS + ((warp - NDW) >> 2) * D_COLS : (warp >> 2) * P_COLS);
  __syncthreads();  // everyone has read the base before the control block is initialised

  // ---- roles ----------------------------------------------------------------------------------------------
  const int unit = isD ? tid - TU : tid;
  const bool has_unit = unit < NB;
  const int blk = has_unit ? unit : NB - 1;
  const bool has_bnd = tid < NBONDS;
  const int bnd = has_bnd ? tid : NBONDS - 1;

  const double* g_cnv = leaf_ptr(a.p.centroid_node_vectors, design);
  const double* g_drive = leaf_ptr(a.p.drive, design);
  const double* ts = a.ts + (long long)design * a.ts_bstride;
  const double* ys = a.ys + (long long)design * a.n_t * 2 * nf;
  const double rtol = a.rtol, atol = a.atol;

  // per-thread topology packed in one register: bit j free, bit 3+j constrained, bit 6+j damped, bit 9 contact seen by this bond
  unsigned flags = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int dof = 3 * blk + j;
    if (has_unit && T.free_of_dof[dof] >= 0) flags |= 1u << j;
    if (has_unit && T.cons_slot[dof] >= 0) flags |= 8u << j;
    if (T.damp_slot[dof] >= 0) flags |= 64u << j;
  }
  auto is_free = [&](int j) { return (flags >> j) & 1u; };
  auto is_cons = [&](int j) { return (flags >> (3 + j)) & 1u; };
  const bool has_cons = (flags & 56u) != 0;
  const bool warp_t0 = __any_sync(0xffffffffu, has_cons);
  auto fidx = [&](int j) { return T.free_of_dof[3 * blk + j]; };  // cold paths only
  // target index (+1) of each DOF in the device objective, 10 bits each; 0 = not a target, 1023 = look it up.  Kept in
  // tensor memory (cold)
  unsigned tgt = 0;
  if (!a.g) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (is_free(j)) {
        const int k = target_index_nl(a, fidx(j));
        tgt |= (unsigned)(k < 0 ? 0 : (k + 1 < 1023 ? k + 1 : 1023)) << (10 * j);
      }
    }
  }
  const uint32_t ta_topo = ta + 2 * (isD ? D_TOPO : P_TOPO);
  auto cot3 = [&](int it, bool is_v, double (&out)[3]) {  // cotangents of the three DOFs of this thread's unit (cold)
    double tw[1];
    tm_ld<1>(ta_topo, tw);
    const unsigned tg = (unsigned)__double2loint(tw[0]);
#pragma unroll
    for (int j = 0; j < 3; ++j)
      out[j] = is_free(j) ? cotangent_k_nl(a, design, it, fidx(j), is_v, (int)((tg >> (10 * j)) & 1023u)) : 0.0;
  };
  // bond * 2 + side attached to each vertex of this thread's unit (0xffff: none), two vertices per register
  unsigned nbp[2] = {0xffffffffu, 0xffffffffu};
#pragma unroll
  for (int l = 0; l < NPB; ++l) {
    const int nb_ = has_unit ? A.node_bond[blk * NPB + l] : -1;
    const unsigned f = nb_ < 0 ? 0xffffu : (unsigned)nb_;
    nbp[l >> 1] = (l & 1) ? ((nbp[l >> 1] & 0x0000ffffu) | (f << 16)) : ((nbp[l >> 1] & 0xffff0000u) | f);
  }
  tmem_st(ta_topo, __hiloint2double(0, (int)tgt));
  tmem_st(ta_topo + 2, __hiloint2double((int)nbp[1], (int)nbp[0]));
  // slot of this unit in the shared-memory table of drive vectors: its rank among the constrained units
  {
    const unsigned bal = __ballot_sync(0xffffffffu, has_cons);
    if (!isD && lane == 0) red[warp] = (double)__popc(bal);
    __syncthreads();
    if (has_cons) {
      int slot = __popc(bal & ((1u << lane) - 1u));
      for (int w = 0; w < (isD ? warp - NDW : warp); ++w) slot += (int)red[w];
      flags |= (unsigned)slot << 10;
      if (!isD) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int cs_ = T.cons_slot[3 * blk + j];
          CV[slot * 6 + j] = cs_ >= 0 ? T.drive_vec0[cs_] : 0.0;
          CV[slot * 6 + 3 + j] = cs_ >= 0 ? T.drive_vec1[cs_] : 0.0;
        }
      }
    }
    __syncthreads();
  }
  int bbp;  // the two blocks of this thread's bond, packed
  { const int2 bb = T.bond_blocks[bnd]; bbp = bb.x | (bb.y << 16); }

  // ---- constants ----------------------------------------------------------------------------------------------
  for (int i = tid; i < NSL * TT; i += TT) SL[i] = 0.0;
  for (int i = tid; i < 28 * TU; i += TT) smem[L::QSP + i] = 0.0;
  for (int i = tid; i < 2 * NSCAL + 5 * NSLOT + 8 * NDRV; i += TT) Sq0[i] = 0.0;
  if (tid == 0) {
    PC[0] = leaf_ptr(a.p.k_stretch, design)[0]; PC[1] = leaf_ptr(a.p.k_shear, design)[0]; PC[2] = leaf_ptr(a.p.k_rot, design)[0];
    PC[3] = 0.0; PC[4] = 0.0; PC[5] = 0.0;
    if (CONTACT) { const double* g_contact = leaf_ptr(a.p.contact, design); PC[3] = g_contact[0]; PC[4] = g_contact[1]; PC[5] = g_contact[2]; }
    C->h = 0; C->h0 = 0; C->d1 = 0; C->s0 = 0; C->s_target = 0; C->s_cur = 0; C->x = 0;
    C->n_steps = 0; C->n_acc = 0; C->n_rhs = 0; C->istep = 0; C->status = 0; C->crossing = 0; C->contact_seen = 0;
  }
  {
    const double* g_inertia = leaf_ptr(a.p.inertia, design);
    const double* g_damp = leaf_ptr(a.p.damping, design);
    if (!isD) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double cdv = 0.0;
        if (DAMP != 0 && is_free(j) && ((flags >> (6 + j)) & 1u)) cdv = DAMP == 2 ? g_damp[T.damp_slot[3 * blk + j]] : g_damp[0];
        INVM[j * TU + tid] = is_free(j) ? 1.0 / g_inertia[fidx(j)] : 0.0;
        CDs[j * TU + tid] = cdv;
      }
    }
  }
  {
    const double* g_ref = leaf_ptr(a.p.reference_vector, design);
    auto edge_angle = [&](int n, int dir) {  // angle of the edge from node n to its next (dir=+1) / previous (dir=-1) node
      const int b = n / NPB, l = n - b * NPB;
      const int m = b * NPB + (dir > 0 ? (l + 1 == NPB ? 0 : l + 1) : (l == 0 ? NPB - 1 : l - 1));
      return atan2(g_cnv[2 * m + 1] - g_cnv[2 * n + 1], g_cnv[2 * m] - g_cnv[2 * n]);
    };
    const int2 nd = T.bond_nodes[bnd];
    const double rx = g_ref[2 * bnd], ry = g_ref[2 * bnd + 1];
    double* bcg = gcol + G_BC;
    bcg[0 * TT] = rx; bcg[1 * TT] = ry;
    bcg[2 * TT] = sqrt(rx * rx + ry * ry); bcg[3 * TT] = 1.0 / sqrt(rx * rx + ry * ry);
    bcg[4 * TT] = g_cnv[2 * nd.x]; bcg[5 * TT] = g_cnv[2 * nd.x + 1];
    bcg[6 * TT] = g_cnv[2 * nd.y]; bcg[7 * TT] = g_cnv[2 * nd.y + 1];
    double da1 = 0.0, da2 = 0.0;
    if (CONTACT) {  // psi1 = (a1_next + th1) - (a2_prev + th2), psi2 = (a2_next + th2) - (a1_prev + th1)
      da1 = edge_angle(nd.x, +1) - edge_angle(nd.y, -1);
      da2 = edge_angle(nd.y, +1) - edge_angle(nd.x, -1);
    }
    bcg[8 * TT] = da1; bcg[9 * TT] = da2;
  }
  for (int q = 0; q < 12 + 4 + 4 * NE3; ++q) gcol[G_SPART + (long long)q * TT] = 0.0;
  // y_bar = g[-1]
  if (isD) {
    double cu[3], cv[3];
    tmem_st_wait();
    cot3(a.n_t - 1, false, cu); cot3(a.n_t - 1, true, cv);
#pragma unroll
    for (int j = 0; j < 3; ++j) { tmem_st(ta + 2 * (D_LU0 + j), cu[j]); tmem_st(ta + 2 * (D_LV0 + j), cv[j]); }
    tmem_st_wait();
  }
  __syncthreads();

  const double inv_n = 1.0 / (double)a.aug_size;
  // Drive table: row e holds the drive channels, their time derivatives and parameter derivatives at the time of
  // evaluation e (0..5 stages of the coming step, 6 interval start, 7 probe).  The rows of a whole step are filled at once
  // by the lanes of the last warp (one lane per row, SIMT-parallel) before the barrier that ends the previous step, so
  // no signal is evaluated inside phases A / B / C.
  auto fill_drive = [&](int row, double t) { drive_row_nl(T.drive_kind, t, g_drive, T.table, drv + row * NDRV); };
  const bool drive_on = T.drive_kind != DFX_DRIVE_ZERO;
  if (drive_on && tid == TT - 32 + EV_INIT && a.n_t >= 2) fill_drive(EV_INIT, ts[a.n_t - 1]);
  __syncthreads();

  // The only loop-carried register is the kind of the next evaluation; the output interval, the quadrature parity and the
  // step size live in the control block (written by thread 0 between barriers).
  constexpr int EV_STOP = 8;
  if (lane == 0) C->ev_w[warp] = a.n_t >= 2 ? EV_INIT : EV_STOP;
  if (tid == 0) { C->i = a.n_t - 1; C->par = 0; C->hst = 0.0; }
  __syncthreads();

  // scratch column of this thread, re-derived from %smid where it is used (not carried in registers)
  auto gcol_now = [&]() {
    unsigned sm_;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_));
    return a.scratch + (long long)sm_ * A.scratch_per_slot + tid;
  };
  // ---- phase A: stage state of this thread's half of its unit, published to shared memory ------------------------
  auto phaseA = [&](auto tag, double* Wcur) {
    constexpr int EV = decltype(tag)::value;
    const double hst = C->hst;
    if (!isD) {
      double us[3], vs[3];
      if constexpr (EV == EV_INIT) {
        const double* yi = ys + (long long)C->i * 2 * nf;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          us[j] = is_free(j) ? __ldcs(&yi[fidx(j)]) : 0.0;
          vs[j] = is_free(j) ? __ldcs(&yi[nf + fidx(j)]) : 0.0;
          tmem_st(ta + 2 * (P_U0 + j), us[j]); tmem_st(ta + 2 * (P_V0 + j), vs[j]);
        }
      } else if constexpr (EV == EV_PROBE) {
        double y0[6], k0[3];
        tm_ld<6>(ta + 2 * P_U0, y0);
        tm_ld<3>(ta + 2 * P_KV, k0);
#pragma unroll
        for (int j = 0; j < 3; ++j) { us[j] = y0[j] - hst * y0[3 + j]; vs[j] = y0[3 + j] + hst * k0[j]; }
      } else {
        stage_P<EV>(ta, tab, hst, us, vs);
      }
#ifndef ABL_NO_CONS
      if (has_cons && drive_on) {
        const double s0_ = drv[EV * NDRV], s1_ = drv[EV * NDRV + 1];
#pragma unroll
        const double* cvp = CV + ((flags >> 10) & 63u) * 6;
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (is_cons(j)) us[j] = cvp[j] * s0_ + cvp[3 + j] * s1_;
      }
#endif
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (!is_free(j)) { vs[j] = 0.0; if (!is_cons(j)) us[j] = 0.0; }
        tmem_st(ta + 2 * (P_TV + j), vs[j]);
      }
      if (has_unit) {
        double sn, cs;
        sincos_fast(us[2], &sn, &cs);
        Us[tid] = us[0]; Us[TU + tid] = us[1]; Us[2 * TU + tid] = us[2]; Us[3 * TU + tid] = sn; Us[4 * TU + tid] = cs;
      }
    } else {
      double lus[3], lvs[3];
      if constexpr (EV == EV_INIT) {
        double y0[6];
        tm_ld<6>(ta + 2 * D_LU0, y0);
#pragma unroll
        for (int j = 0; j < 3; ++j) { lus[j] = y0[j]; lvs[j] = y0[3 + j]; }
      } else if constexpr (EV == EV_PROBE) {
        double y0[6], k0[3], k1[3];
        tm_ld<6>(ta + 2 * D_LU0, y0);
        tm_ld<3>(ta + 2 * D_KLU, k0);
        tm_ld<3>(ta + 2 * D_KLV, k1);
#pragma unroll
        for (int j = 0; j < 3; ++j) { lus[j] = y0[j] + hst * k0[j]; lvs[j] = y0[3 + j] + hst * k1[j]; }
      } else {
        stage_D<EV>(ta, tab, hst, lus, lvs);
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (!is_free(j)) { lus[j] = 0.0; lvs[j] = 0.0; }
        tmem_st(ta + 2 * (D_TLU + j), lus[j]);
        if (has_unit) Wcur[j * TU + unit] = lvs[j] * INVM[j * TU + unit];
      }
    }
    tmem_st_wait();
  };

  // ---- phase C: gather the bond slots of this thread's half of its unit, stage derivatives, quadratures -----------
  // Returns the thread's contribution to the error norm (last stage) or to d2 of initial_step_size (probe).
  auto phaseC = [&](auto tag, const double* Wcur) -> double {
    constexpr int EV = decltype(tag)::value;
    constexpr int KIDX = Ev<EV>::kidx, MODE = Ev<EV>::mode;
    double* const gcol = gcol_now();  // (shadows: the scratch pointers are re-derived, not carried round the loop)
    double* const qg = gcol + G_Q;
    const double* const gbase = gcol - tid;
#ifdef ABL_NO_Q
    constexpr bool WANT_Q = false;
#else
    constexpr bool WANT_Q = MODE != 1;
#endif
    double accq = 0.0;
    QC qc;
    qc.par = C->par; qc.crossing = C->crossing != 0; qc.h = C->hst; qc.x = C->x; qc.atol = a.atol; qc.rtol = a.rtol;
    // vertex -> bond table of this thread's unit (two vertices per 32-bit word, 0xffff: none), parked in tensor memory
    unsigned nbp[2];
    {
      uint32_t lo, hi;
      tmem_ld_issue(ta_topo + 2, lo, hi);
      tmem_ld_wait();
      tmem_pin(lo, hi);
      nbp[0] = lo; nbp[1] = hi;
    }
    auto nbq = [&](int l) { const int f = (int)((nbp[l >> 1] >> ((l & 1) * 16)) & 0xffffu); return f == 0xffff ? -1 : f; };
    const bool seen = CONTACT && C->contact_seen;
    if (!isD) {
      // dense scalar leaves: P warp w sums the partials of bonds [192 q, 192 q + 192), q = w / 3, of leaf w % 3 (and of
      // contact leaf w % 3 once a bond has touched), fixed order; loads issued first, consumed last
      // (accumulating stages 2..4: the partials are summed by the last warp, which has no bonds, during the NEXT bond phase)
      constexpr bool REDUCE_NOW = WANT_Q && !(EV >= 1 && EV <= 4);
      double rk[6] = {0, 0, 0, 0, 0, 0}, rc[6] = {0, 0, 0, 0, 0, 0};
#ifndef ABL_NO_REDUCE
      if (REDUCE_NOW) {
        const double* sp = gbase + G_SPART + (long long)((EV & 1) * 6 + warp % 3) * TT + (warp / 3) * 192 + lane;
#pragma unroll
        for (int q = 0; q < 6; ++q) rk[q] = __ldcg(&sp[q * 32]);
        if (seen) {
#pragma unroll
          for (int q = 0; q < 6; ++q) rc[q] = __ldcg(&sp[3 * TT + q * 32]);
        }
      }
#endif
      double F[3] = {0, 0, 0}, val[10], An[NPB], Ap[NPB];
#pragma unroll
      for (int l = 0; l < 4; ++l) val[6 + l] = 0.0;
#pragma unroll
      for (int l = 0; l < NPB; ++l) {
        An[l] = 0.0; Ap[l] = 0.0;
        const int nb_ = nbq(l);
        if (nb_ >= 0) {
          const int b = nb_ >> 1;
          const bool second = nb_ & 1;
          const double sg = second ? -1.0 : 1.0;
          F[0] += sg * SL[b]; F[1] += sg * SL[TT + b]; F[2] += SL[(second ? 3 : 2) * TT + b];
          if (WANT_Q) {
            val[6 + l] = SL[(second ? 10 : 8) * TT + b];
            if (seen) {
              // dS/dalpha = -(dual part of dE/dalpha): a1next:+e1, a1prev:-e2, a2next:+e2, a2prev:-e1
              const double e1d = SL[12 * TT + b], e2d = SL[13 * TT + b];
              An[l] = second ? -e2d : -e1d;
              Ap[l] = second ? e1d : e2d;
            }
          }
        }
      }
      // partial sums of the dense scalar leaves (the loads were issued before the gather)
      double sk = ((rk[0] + rk[1]) + (rk[2] + rk[3])) + (rk[4] + rk[5]);
      double scn = seen ? ((rc[0] + rc[1]) + (rc[2] + rc[3])) + (rc[4] + rc[5]) : 0.0;
      double vst[3];
      tm_ld<3>(ta + 2 * P_TV, vst);
      double p_damp = 0.0;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double kvv = 0.0;
        val[j] = 0.0; val[3 + j] = 0.0;
        if (is_free(j)) {
          const double wj = Wcur[j * TU + tid], cdj = CDs[j * TU + tid];
          const double acc = (F[j] - cdj * vst[j]) * INVM[j * TU + tid];
          kvv = -acc;
          val[j] = -wj * acc;
          if (DAMP != 0 && ((flags >> (6 + j)) & 1u)) { if (DAMP == 2) val[3 + j] = -wj * vst[j]; else p_damp -= wj * vst[j]; }
        }
        tmem_st(ta + 2 * (P_KV + KIDX * 3 + j), kvv);
      }
      if (WANT_Q) {
        if (seen) {
          bool any_contact = false;
#pragma unroll
          for (int l = 0; l < NPB; ++l) any_contact |= (An[l] != 0.0) | (Ap[l] != 0.0);
          if (any_contact) {
            // contact chain of the centroid_node_vectors cotangent (x components; the D thread does y): edge l -> l+1 is
            // node l's "next" edge and, reversed, node (l+1)'s "previous" edge
#pragma unroll
            for (int l = 0; l < NPB; ++l) {
              const int ln = l + 1 == NPB ? 0 : l + 1;
              const int n = blk * NPB + l, m = blk * NPB + ln;
              const double ex = g_cnv[2 * m] - g_cnv[2 * n], ey = g_cnv[2 * m + 1] - g_cnv[2 * n + 1];
              const double wx = -(An[l] + Ap[ln]) * ey / (ex * ex + ey * ey);
              val[6 + ln] += wx; val[6 + l] -= wx;
            }
          }
        }
        double* qs = smem + L::QSP + tid;
        {
          const double v3[3] = {val[0], val[1], val[2]};
          accq += quad_entries<MODE, 3>(qc, tab, qg, 0, qs, qs + 10 * TU, TU, v3);
        }
        if (DAMP == 2) {
          const double v3[3] = {val[3], val[4], val[5]};
          accq += quad_entries<MODE, 3>(qc, tab, qg, 3, qs + 3 * TU, qs + 13 * TU, TU, v3);
        }
        {
          const double v4[4] = {val[6], val[7], val[8], val[9]};
          accq += quad_entries<MODE, 4>(qc, tab, qg, 6, qs + 6 * TU, qs + 16 * TU, TU, v4);
        }
        if (DAMP == 1) { const double tot = warp_sum(p_damp); if (lane == 0) scal_slot<MODE>(tab, accb, SLOT_DAMP + warp, tot); }
#ifndef ABL_NO_REDUCE
        if (REDUCE_NOW) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            sk += __shfl_xor_sync(0xffffffffu, sk, o);
            if (CONTACT) scn += __shfl_xor_sync(0xffffffffu, scn, o);
          }
          if (lane == 0) scal_slot<MODE>(tab, accb, SLOT_K + (warp % 3) * 4 + warp / 3, sk);
          if (CONTACT && lane == 1 && seen) scal_slot<MODE>(tab, accb, SLOT_K + (3 + warp % 3) * 4 + warp / 3, scn);
        }
#endif
      }
    } else {
      double HW[3] = {0, 0, 0}, val[4] = {0, 0, 0, 0}, An[NPB], Ap[NPB];
#pragma unroll
      for (int l = 0; l < NPB; ++l) {
        An[l] = 0.0; Ap[l] = 0.0;
        const int nb_ = nbq(l);
        if (nb_ >= 0) {
          const int b = nb_ >> 1;
          const bool second = nb_ & 1;
          const double sg = second ? -1.0 : 1.0;
          HW[0] -= sg * SL[4 * TT + b]; HW[1] -= sg * SL[5 * TT + b]; HW[2] += SL[(second ? 7 : 6) * TT + b];
          if (WANT_Q) {
            val[l] = SL[(second ? 11 : 9) * TT + b];
            if (seen) {
              const double e1d = SL[12 * TT + b], e2d = SL[13 * TT + b];
              An[l] = second ? -e2d : -e1d;
              Ap[l] = second ? e1d : e2d;
            }
          }
        }
      }
      double lust[3];
      tm_ld<3>(ta + 2 * D_TLU, lust);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double kluv = 0.0, klvv = 0.0;
        if (is_free(j)) { kluv = -HW[j]; klvv = lust[j] - CDs[j * TU + unit] * Wcur[j * TU + unit]; }
        tmem_st(ta + 2 * (D_KLU + KIDX * 3 + j), kluv);
        tmem_st(ta + 2 * (D_KLV + KIDX * 3 + j), klvv);
      }
      if (WANT_Q) {
        if (seen) {
          bool any_contact = false;
#pragma unroll
          for (int l = 0; l < NPB; ++l) any_contact |= (An[l] != 0.0) | (Ap[l] != 0.0);
          if (any_contact) {
#pragma unroll
            for (int l = 0; l < NPB; ++l) {
              const int ln = l + 1 == NPB ? 0 : l + 1;
              const int n = blk * NPB + l, m = blk * NPB + ln;
              const double ex = g_cnv[2 * m] - g_cnv[2 * n], ey = g_cnv[2 * m + 1] - g_cnv[2 * n + 1];
              const double wy = (An[l] + Ap[ln]) * ex / (ex * ex + ey * ey);
              val[ln] += wy; val[l] -= wy;
            }
          }
        }
        double* qs = smem + L::QSD + unit;
        accq += quad_entries<MODE, 4>(qc, tab, qg, 0, qs, qs + 4 * TU, TU, val);
        // t0_bar and the drive parameters only receive contributions from constrained DOFs
#ifndef ABL_NO_T0
        if (warp_t0 && drive_on) {
          // every t0 / drive-parameter integrand is -(A_q S0 + B_q S1) with S_c = sum over the constrained DOFs of
          // (H w)_dof * drive_vec_c[dof] and (A_q, B_q) the derivatives of the two drive channels: two warp sums serve all leaves
          double S0 = 0.0, S1 = 0.0;
          if (has_cons) {
            const double* cvp = CV + ((flags >> 10) & 63u) * 6;
#pragma unroll
            for (int j = 0; j < 3; ++j)
              if (is_cons(j)) { S0 = fma(HW[j], cvp[j], S0); S1 = fma(HW[j], cvp[3 + j], S1); }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            S0 += __shfl_xor_sync(0xffffffffu, S0, o);
            S1 += __shfl_xor_sync(0xffffffffu, S1, o);
          }
          if (lane <= T.n_drive_params) {
            const double* dr = drv + EV * NDRV;
            const double Aq = lane == 0 ? dr[2] : dr[3 + lane], Bq = lane == 0 ? dr[3] : dr[8 + lane];
            scal_slot<MODE>(tab, accb, SLOT_T + lane * NDW + (warp - NDW), -(Aq * S0 + Bq * S1));
          }
        }
#endif
      }
    }
    tmem_st_wait();
    return accq;
  };

#define DFX_A3_DISPATCH(CALL)                                                                          \
  switch (ev) {                                                                                        \
    case 0: CALL(IC<0>{}); break;                                             \
    case 1: CALL(IC<1>{}); break;                                             \
    case 2: CALL(IC<2>{}); break;                                             \
    case 3: CALL(IC<3>{}); break;                                             \
    case 4: CALL(IC<4>{}); break;                                             \
    case 5: CALL(IC<5>{}); break;                                             \
    case EV_INIT: CALL(IC<EV_INIT>{}); break;                                 \
    default: CALL(IC<EV_PROBE>{}); break;                                     \
  }

#ifdef DFX_PHASE_TIMERS
  // cycles of lane 0 of every warp: A | wait A->B | B | wait B->C | C | what follows (DFX_PHASE_TIMERS builds only)
  long long pt_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, pt_mark = clock64();  // 6..9: parts of the step end (see the marks)
#define PT_MARK(k) do { const long long now_ = clock64(); pt_acc[k] += now_ - pt_mark; pt_mark = now_; } while (0)
#else
#define PT_MARK(k) do { } while (0)
#endif
  while (true) {
    PT_MARK(5);
    int ev = *(volatile int*)&C->ev_w[warp];
    if (ev == EV_STOP) break;
    bool running = true;
    double* Wcur = Ws + (ev & 1) * 3 * TU;  // consecutive evaluations never have the same parity (6 -> 7 -> 0..5 -> 0 | 6)
    // bond constants: fetched from L2 now, used after the barrier
    double bc[NBC];
#pragma unroll
    {
      const double* gc_ = gcol_now();
#pragma unroll
      for (int k = 0; k < (CONTACT ? NBC : NBC - 2); ++k) bc[k] = gc_[G_BC + (long long)k * TT];
    }
#define DFX_A3_A(TAG) phaseA(TAG, Wcur)
    DFX_A3_DISPATCH(DFX_A3_A)
#undef DFX_A3_A
    PT_MARK(0);
    __syncthreads();
    PT_MARK(1);

    // ================= phase B: this thread's bond (one copy of the code for every kind of evaluation) =================
    double accq = 0.0;  // this thread's contribution to the error norm / probe norm
    ev = *(volatile int*)&C->ev_w[warp];
    if (has_bnd) {
      double* const gcol = gcol_now();
      double* const qg = gcol + G_Q;
      const bool want_q = ev != 0;  // the second stage of a step has zero weight in every combination
      const int b1 = bbp & 0xffff, b2 = bbp >> 16;
      const int b = tid;
      // The ligament gradient on dual numbers (bond_gradient of dfx_device.cuh, ligament energy, with the parameter
      // cotangents), written out in two sections so that the block rotations are not kept in registers across the
      // energy arithmetic: they are read again from shared memory for the node-vector cotangents.
      Dual gdx, gdy;          // dE/d(dU)
      double a1 = 0.0, a2 = 0.0, p_c0 = 0, p_c1 = 0, p_c2 = 0;
      double gks_d, gksh_d, gkr_d;
      {
        BlockState<Dual> s1, s2;
        make_block(Us[b1], Us[TU + b1], Us[2 * TU + b1], Us[3 * TU + b1], Us[4 * TU + b1], Wcur[b1], Wcur[TU + b1], Wcur[2 * TU + b1], s1);
        make_block(Us[b2], Us[TU + b2], Us[2 * TU + b2], Us[3 * TU + b2], Us[4 * TU + b2], Wcur[b2], Wcur[TU + b2], Wcur[2 * TU + b2], s2);
        const double r0x = bc[0], r0y = bc[1], L0 = bc[2], iL0 = bc[3], r1x = bc[4], r1y = bc[5], r2x = bc[6], r2y = bc[7];
        const double ks = PC[0], ksh = PC[1], kr = PC[2];
        Dual c1m = s1.c - 1.0, c2m = s2.c - 1.0;
        Dual dx = (s2.x + c2m * r2x - s2.s * r2y) - (s1.x + c1m * r1x - s1.s * r1y) + r0x;
        Dual dy = (s2.y + s2.s * r2x + c2m * r2y) - (s1.y + s1.s * r1x + c1m * r1y) + r0y;
        {
          // d(node)/d(theta) = R'(theta) r of the two nodes: only needed for the torques at the end -- parked in this bond's
          // own (still unused) result slots instead of eight registers across the energy arithmetic
          const Dual t1x = -(s1.s * r1x) - s1.c * r1y, t1y = s1.c * r1x - s1.s * r1y;
          const Dual t2x = -(s2.s * r2x) - s2.c * r2y, t2y = s2.c * r2x - s2.s * r2y;
          SL[6 * TT + b] = t1x.v; SL[7 * TT + b] = t1x.d; SL[8 * TT + b] = t1y.v; SL[9 * TT + b] = t1y.d;
          SL[10 * TT + b] = t2x.v; SL[11 * TT + b] = t2x.d;
          if (CONTACT) { SL[12 * TT + b] = t2y.v; SL[13 * TT + b] = t2y.d; } else { SL[4 * TT + b] = t2y.v; SL[5 * TT + b] = t2y.d; }
        }
        Dual dth = s2.th - s1.th;
        Dual mean = (s2.th + s1.th) * 0.5;
        const double L0sq = L0 * L0;
        Dual L2 = dx * dx + dy * dy;
        const double rinv = rsqrt_pos(L2.v);
        Dual iL2 = inv_from(L2, rinv);
        Dual L = len_from(L2, rinv);
        double gv;
        {
          const double cp = r0x * iL0, sp = r0y * iL0;
          const double ex = dx.v * rinv, ey = dy.v * rinv;
          gv = wrap_value(angle_of_unit(cp * ey - sp * ex, cp * ex + sp * ey) - mean.v);
        }
        Dual gam(gv, (dx.v * dy.d - dy.v * dx.d) * iL2.v - mean.d);
        Dual ext = L - L0;
        Dual A = ext * ks * L * iL2;  // ks (L-L0)/L
        Dual M = gam * (ksh * L0sq);  // dE/dgamma
        Dual Bc = M * iL2;
        gdx = A * dx - Bc * dy;
        gdy = A * dy + Bc * dx;
        Dual bend = dth * kr;
        asm volatile("" ::: "memory");
        const Dual t1x(SL[6 * TT + b], SL[7 * TT + b]), t1y(SL[8 * TT + b], SL[9 * TT + b]), t2x(SL[10 * TT + b], SL[11 * TT + b]);
        const Dual t2y = CONTACT ? Dual(SL[12 * TT + b], SL[13 * TT + b]) : Dual(SL[4 * TT + b], SL[5 * TT + b]);
        Dual f1t = M * (-0.5) - (gdx * t1x + gdy * t1y) - bend;
        Dual f2t = M * (-0.5) + (gdx * t2x + gdy * t2y) + bend;
        if (CONTACT) {
          Dual psi1 = wrapT(s1.th - s2.th + bc[8]);
          Dual psi2 = wrapT(s2.th - s1.th + bc[9]);
          const double cmin = PC[3], ccut = PC[4], ckc = PC[5];
          const bool act1 = !(psi1.v < cmin) && psi1.v < ccut, act2 = !(psi2.v < cmin) && psi2.v < ccut;
          if (act1 || act2) {
            Dual e1, e2, m1, m2, c1, c2, k1, k2;
            contact_term<Dual>(psi1, cmin, ccut, ckc, e1, m1, c1, k1);
            contact_term<Dual>(psi2, cmin, ccut, ckc, e2, m2, c2, k2);
            f1t = f1t + e1 - e2;
            f2t = f2t + e2 - e1;
            a1 = e1.d; a2 = e2.d;
            p_c0 = -(m1.d + m2.d); p_c1 = -(c1.d + c2.d); p_c2 = -(k1.d + k2.d);
            if (!(flags & 512u)) { flags |= 512u; C->contact_seen = 1; }
          }
        }
        // forces on the two ends are equal and opposite: store (gdx, gdy) once, the two torques separately
        SL[b] = gdx.v; SL[TT + b] = gdy.v; SL[2 * TT + b] = -f1t.v; SL[3 * TT + b] = -f2t.v;
        SL[4 * TT + b] = gdx.d; SL[5 * TT + b] = gdy.d; SL[6 * TT + b] = f1t.d; SL[7 * TT + b] = f2t.d;
        if (CONTACT) { SL[12 * TT + b] = a1; SL[13 * TT + b] = a2; }  // (always: the slots were used as scratch above)
        // parameter cotangent integrands: -(dual part of dE/dp)
        gks_d = ext.v * ext.d; gksh_d = gam.v * gam.d * L0sq; gkr_d = dth.v * dth.d;
        Dual dE_dL0 = gam * gam * (ksh * L0) - ext * ks;
#ifndef ABL_NO_BQ
        if (want_q) {
          // reference-vector quadratures of this bond (running sums in L2, thread private): the accumulating stages add
          // their terms with fire-and-forget reductions at L2 -- no load, no latency, nothing carried into phase C
          const double qb0 = -(gdx.d + dE_dL0.d * (r0x * iL0) + M.d * (r0y / L0sq));
          const double qb1 = -(gdy.d + dE_dL0.d * (r0y * iL0) - M.d * (r0x / L0sq));
          double* bq = gcol + G_BQ;
          if (ev >= 2 && ev <= 4) {
            const double cs = tab.c_sol[ev + 1], ce = tab.c_err[ev + 1];
            atomicAdd(bq, cs * qb0); atomicAdd(bq + TT, cs * qb1);
            atomicAdd(bq + 2 * TT, ce * qb0); atomicAdd(bq + 3 * TT, ce * qb1);
            if (C->crossing) {
              const double cm = tab.c_mid[ev + 1];
              double* k7p = qg + (long long)((3 - C->par) * NE3 + 10) * TT;
              atomicAdd(k7p, cm * qb0); atomicAdd(k7p + TT, cm * qb1);
            }
          } else {
            QC qc;
            qc.par = C->par; qc.crossing = C->crossing != 0; qc.h = C->hst; qc.x = C->x; qc.atol = a.atol; qc.rtol = a.rtol;
            accq = bond_quad_rare_nl(ev == EV_INIT ? 0 : (ev == EV_PROBE ? 7 : ev + 1), qc, tab, qg, bq, qb0, qb1);
          }
        }
#endif
      }
      asm volatile("" ::: "memory");  // the block rotations are re-read below instead of being carried in registers
      {
        const double sn1 = Us[3 * TU + b1], cs1 = Us[4 * TU + b1], w1 = Wcur[2 * TU + b1];
        const double sn2 = Us[3 * TU + b2], cs2 = Us[4 * TU + b2], w2 = Wcur[2 * TU + b2];
        const Dual s1(sn1, cs1 * w1), c1m(cs1 - 1.0, -sn1 * w1), s2(sn2, cs2 * w2), c2m(cs2 - 1.0, -sn2 * w2);
        // dE/d(centroid_node_vector) of the two nodes: gr1 = -(R1 - I)^T g, gr2 = (R2 - I)^T g; stored with the sign of the integrand
        SL[8 * TT + b] = (c1m * gdx + s1 * gdy).d; SL[9 * TT + b] = -(s1 * gdx - c1m * gdy).d;
        SL[10 * TT + b] = -(c2m * gdx + s2 * gdy).d; SL[11 * TT + b] = -(c2m * gdy - s2 * gdx).d;
      }
      if (want_q) {
        // d(w.F)/dp = -(dual part of dE/dp).  The reference-vector quadratures are updated at the end of phase C; the dense
        // scalar leaves go through per-bond partials in the scratch, summed in phase C by the P warps.
#ifndef ABL_NO_SPART
        double* sp = gcol + G_SPART + (long long)((ev & 1) * 6) * TT;
        sp[0] = -gks_d; sp[TT] = -gksh_d; sp[2 * TT] = -gkr_d;
        if (CONTACT && (flags & 512u)) { sp[3 * TT] = p_c0; sp[4 * TT] = p_c1; sp[5 * TT] = p_c2; }
#endif
      }
    }
#ifndef ABL_NO_REDUCE
    if (warp == NW - 1 && ev >= 2 && ev <= 5) {
      // the dense scalar leaves of the PREVIOUS evaluation (an accumulating stage, mode = ev): this warp has no bonds, it
      // sums the 768 per-bond partials of every leaf (fixed order) while the others do their bond arithmetic
      const bool seen = CONTACT && C->contact_seen;
      const double* sp = gbase + G_SPART + (long long)(((ev - 1) & 1) * 6) * TT + lane;
      double tot[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int l = 0; l < 6; ++l) {
        if (l < 3 || seen) {
          double v[NW];
#pragma unroll
          for (int q = 0; q < NW; ++q) v[q] = __ldcg(&sp[(long long)l * TT + q * 32]);
          double sacc = 0.0;
#pragma unroll
          for (int q = 0; q < NW; ++q) sacc += v[q];
          tot[l] = sacc;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int l = 0; l < 6; ++l) if (l < 3 || CONTACT) tot[l] += __shfl_xor_sync(0xffffffffu, tot[l], o);
      }
      // lane 4 l + q updates partial slot q of leaf l: the total goes to slot 0, the others receive 0 (mode 2 rebuilds all four)
      if (lane < (seen ? 24 : 12)) {
        const int l = lane >> 2, q = lane & 3;
        double mine = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) if (k == l && q == 0) mine = tot[k];
        const int slot = SLOT_K + l * 4 + q;
        switch (ev) {
          case 2: scal_slot<2>(tab, accb, slot, mine); break;
          case 3: scal_slot<3>(tab, accb, slot, mine); break;
          case 4: scal_slot<4>(tab, accb, slot, mine); break;
          default: scal_slot<5>(tab, accb, slot, mine); break;
        }
      }
    }
#endif
    PT_MARK(2);
    __syncthreads();
    PT_MARK(3);

#define DFX_A3_C(TAG) accq += phaseC(TAG, Wcur)
    ev = *(volatile int*)&C->ev_w[warp];
    Wcur = Ws + (ev & 1) * 3 * TU;
    DFX_A3_DISPATCH(DFX_A3_C)
#undef DFX_A3_C
    PT_MARK(4);

    // ================= what follows the evaluation =================
    if (ev < 5) {
      __syncwarp();
      if (lane == 0) C->ev_w[warp] = ev + 1;
      __syncwarp();
      continue;
    }
    int i = C->i, par = C->par;
    const double hst = C->hst;
    double* const qg = gcol_now() + G_Q;
    if (ev == EV_INIT) {
      double sd0 = 0, sd1 = 0, pt = 0.0;
      if (!isD) {
        double y0[6], k0[3], cu[3], cv[3];
        tm_ld<6>(ta + 2 * P_U0, y0);
        tm_ld<3>(ta + 2 * P_KV, k0);
        cot3(i, false, cu); cot3(i, true, cv);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (is_free(j)) {
            // t_bar = func(ys[i], ts[i]) . g[i] with func = (v0, -kv[0])
            pt += y0[3 + j] * cu[j] - k0[j] * cv[j];
            const double su = atol + fabs(y0[j]) * rtol, sv = atol + fabs(y0[3 + j]) * rtol;
            const double a0 = y0[j] / su, a1 = y0[3 + j] / sv, b0 = -y0[3 + j] / su, b1 = k0[j] / sv;
            sd0 += a0 * a0 + a1 * a1;
            sd1 += b0 * b0 + b1 * b1;
          }
        }
      } else {
        double y0[6], k0[3], k1[3];
        tm_ld<6>(ta + 2 * D_LU0, y0);
        tm_ld<3>(ta + 2 * D_KLU, k0);
        tm_ld<3>(ta + 2 * D_KLV, k1);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (is_free(j)) {
            const double slu = atol + fabs(y0[j]) * rtol, slv = atol + fabs(y0[3 + j]) * rtol;
            const double a2 = y0[j] / slu, a3 = y0[3 + j] / slv, b2 = k0[j] / slu, b3 = k1[j] / slv;
            sd0 += a2 * a2 + a3 * a3;
            sd1 += b2 * b2 + b3 * b3;
          }
        }
      }
      const double t_bar = block_sum(pt, red);
      if (tid == 0) {
        if (a.ts_bar) a.ts_bar[(long long)design * a.n_t + i] = t_bar;
        Sq0[SC_T0] -= t_bar;
      }
      __syncthreads();
      // initial_step_size over the whole augmented vector
      {
        double q0v[NE3], k1v[NE3];  // all loads in flight together (one L2 round trip instead of twelve)
#pragma unroll
        for (int e = 0; e < NE3; ++e) {
          q0v[e] = __ldcg(&qg[(long long)(par * NE3 + e) * TT]); k1v[e] = __ldcg(&qg[(long long)((2 + par) * NE3 + e) * TT]);
        }
#pragma unroll
        for (int e = 0; e < NE3; ++e) {
          const double is = rcp_pos(atol + fabs(q0v[e]) * rtol);
          sd0 = fma(q0v[e] * is, q0v[e] * is, sd0); sd1 = fma(k1v[e] * is, k1v[e] * is, sd1);
        }
      }
      if (tid < NSCAL) {
        const double s = atol + fabs(Sq0[tid]) * rtol;
        const double a0 = Sq0[tid] / s, b0 = scal_total(accb, 0, tid) / s;
        sd0 += a0 * a0; sd1 += b0 * b0;
      }
      const double d0 = sqrt(block_sum(sd0, red));
      const double d1 = sqrt(block_sum(sd1, red));
      const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
      if (tid == 0) { C->s0 = -ts[i]; C->s_target = -ts[i - 1]; C->h0 = h0; C->d1 = d1; C->n_rhs += 1; C->crossing = 0; C->hst = h0; }
      if (drive_on && tid == TT - 32 + EV_PROBE) fill_drive(EV_PROBE, ts[i] - h0);
      if (lane == 0) C->ev_w[warp] = EV_PROBE;
      __syncthreads();
    } else if (ev == EV_PROBE) {
      double sd2 = accq;
      if (!isD) {
        double y0[6], k01[6], vs[3];
        tm_ld<6>(ta + 2 * P_U0, y0);
        tm_ld<6>(ta + 2 * P_KV, k01);
        tm_ld<3>(ta + 2 * P_TV, vs);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (is_free(j)) {
            const double su = atol + fabs(y0[j]) * rtol, sv = atol + fabs(y0[3 + j]) * rtol;
            const double b0 = (-vs[j] + y0[3 + j]) / su, b1 = (k01[3 + j] - k01[j]) / sv;
            sd2 += b0 * b0 + b1 * b1;
          }
        }
      } else {
        double y0[6], ku[6], kw[6];
        tm_ld<6>(ta + 2 * D_LU0, y0);
        tm_ld<6>(ta + 2 * D_KLU, ku);
        tm_ld<6>(ta + 2 * D_KLV, kw);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (is_free(j)) {
            const double slu = atol + fabs(y0[j]) * rtol, slv = atol + fabs(y0[3 + j]) * rtol;
            const double b2 = (ku[3 + j] - ku[j]) / slu, b3 = (kw[3 + j] - kw[j]) / slv;
            sd2 += b2 * b2 + b3 * b3;
          }
        }
      }
      __syncthreads();  // scalar partials of this evaluation complete
      if (tid < NSCAL) {
        const double s = atol + fabs(Sq0[tid]) * rtol;
        const double b0 = (scal_total(accb, 1, tid) - scal_total(accb, 0, tid)) / s;
        sd2 += b0 * b0;
      }
      const double h0 = C->h0, d1 = C->d1, s0 = C->s0, s_target = C->s_target;
      const double d2 = sqrt(block_sum(sd2, red)) / h0;
      double h1;
      if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
      else h1 = pow(0.01 / (a.init_step_variant == 0 ? d1 + d2 : fmax(d1, d2)), 0.2);
      const double h = fmin(100.0 * h0, h1);
      bool stop = false, empty = false;
      int status = 0;
      if (!(s0 < s_target)) empty = true;  // empty interval (repeated output time): nothing to integrate
      else if (!(h > 0.0)) { status = DFX_STATUS_DT_UNDERFLOW; stop = true; }
      else if (0 >= a.max_steps) { status = DFX_STATUS_MAX_STEPS; stop = true; }
      if (tid == 0) {
        C->n_rhs += 1; C->h = h; C->s_cur = s0; C->istep = 0; C->status |= status;
        if (!empty && !stop) { const double s_new = s0 + h; C->crossing = !(s_new < s_target); C->x = (s_target - s0) / (s_new - s0); }
      }
      if (empty) { if (--i < 1) running = false; else ev = EV_INIT; }
      else if (stop) running = false;
      else ev = 0;
      if (lane == 0) C->ev_w[warp] = running ? ev : EV_STOP;
      if (tid == 0) { C->hst = h; C->i = i; }
      if (drive_on && running && warp == NW - 1) {
        if (ev == 0 && lane < 6) fill_drive(lane, -(s0 + h * tab.alpha[lane]));
        else if (ev == EV_INIT && lane == EV_INIT) fill_drive(EV_INIT, ts[i]);
      }
      __syncthreads();
    } else {
      double se = accq;
      const double h = hst;
      const bool crossing = C->crossing != 0;
      const double xq = C->x, s_cur = C->s_cur;
      double y1a[3], y1b[3];  // P: u, v at the end of the step; D: lambda_u, lambda_v
      if (!isD) {
        double y0[6], kvh[21];  // the whole velocity-derivative history with two wide loads
        tm_ld<6>(ta + 2 * P_U0, y0);
        tm_ld<21>(ta + 2 * P_KV, kvh);

#pragma unroll
        for (int j = 0; j < 3; ++j) {
          double eu = 0.0, evv = 0.0, su = 0.0, sv = 0.0;
#pragma unroll
          for (int l = 0; l < 7; ++l) {
            const double kvl = kvh[3 * l + j];
            eu = fma(tab.e2[l], kvl, eu); evv = fma(tab.c_err[l], kvl, evv);
            su = fma(tab.s2[l], kvl, su); sv = fma(tab.c_sol[l], kvl, sv);
          }
          y1a[j] = y0[j] - h * (tab.sum_sol * y0[3 + j] + h * su);
          y1b[j] = y0[3 + j] + h * sv;
          eu = -h * (tab.sum_err * y0[3 + j] + h * eu);
          evv *= h;
          if (is_free(j)) {
            const double r0 = eu * rcp_pos(atol + rtol * fmax(fabs(y0[j]), fabs(y1a[j])));
            const double r1 = evv * rcp_pos(atol + rtol * fmax(fabs(y0[3 + j]), fabs(y1b[j])));
            se += r0 * r0 + r1 * r1;
          }
        }
      } else {
        double y0[6];
        tm_ld<6>(ta + 2 * D_LU0, y0);
#pragma unroll
        for (int q = 0; q < 2; ++q) {  // lambda_u then lambda_v: one wide load of the 7 x 3 history each
          double kh[21];
          tm_ld<21>(ta + 2 * (q == 0 ? D_KLU : D_KLV), kh);
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            double el = 0.0, sl = 0.0;
#pragma unroll
            for (int l = 0; l < 7; ++l) { el = fma(tab.c_err[l], kh[3 * l + j], el); sl = fma(tab.c_sol[l], kh[3 * l + j], sl); }
            const double y0q = y0[3 * q + j], y1q = y0q + h * sl;
            if (q == 0) y1a[j] = y1q; else y1b[j] = y1q;
            el *= h;
            if (is_free(j)) {
              const double r = el * rcp_pos(atol + rtol * fmax(fabs(y0q), fabs(y1q)));
              se += r * r;
            }
          }
        }
      }
      PT_MARK(6);  // error contributions of the thread's own state
      __syncthreads();  // scalar running sums of the last stage complete
      if (warp < NSCAL) {  // warp k: scalar leaf k
        double tot[5];
        scal_totals_warp(accb, warp, lane, tot);
        if (lane == 0) {
          const double q0 = Sq0[warp];
          const double q1 = q0 + h * tot[2];
          const double r = h * tot[3] * rcp_pos(atol + rtol * fmax(fabs(q0), fabs(q1)));
          se += r * r;
          Sqnew[warp] = crossing ? interp_eval(q0, q1, q0 + h * tot[4], h * tot[0], h * tot[1], xq) : q1;
        }
      }
      const double ratio = sqrt(block_sum(se, red) * inv_n);
      PT_MARK(7);  // barrier + scalar totals + norm
      const long long istep = C->istep + 1;
      const double s_target = C->s_target;
      bool interval_done = false, stop = false;
      int status = 0;
      double h_new = h, s_new = s_cur;
      bool accepted = false;
      if (!isfinite(ratio)) { status = DFX_STATUS_NONFINITE; stop = true; }
      else {
        accepted = ratio <= 1.0;
        if (accepted) {
          if (crossing) {
            // interval finished: cotangents interpolated at s_target, plus g[i-1]
            if (isD) {
              double y0[6], nl[6];
              tm_ld<6>(ta + 2 * D_LU0, y0);
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                double klu[7], klv[7];
                tm_lds<7>(ta + 2 * (D_KLU + j), 3, klu);
                tm_lds<7>(ta + 2 * (D_KLV + j), 3, klv);
                double mlu = 0.0, mlv = 0.0;
#pragma unroll
                for (int l = 0; l < 7; ++l) { mlu = fma(tab.c_mid[l], klu[l], mlu); mlv = fma(tab.c_mid[l], klv[l], mlv); }
                const double nlu = interp_eval(y0[j], y1a[j], y0[j] + h * mlu, h * klu[0], h * klu[6], xq);
                const double nlv = interp_eval(y0[3 + j], y1b[j], y0[3 + j] + h * mlv, h * klv[0], h * klv[6], xq);
                nl[j] = nlu; nl[3 + j] = nlv;
              }
              double cu[3], cv[3];
              cot3(i - 1, false, cu); cot3(i - 1, true, cv);
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                tmem_st(ta + 2 * (D_LU0 + j), is_free(j) ? nl[j] + cu[j] : 0.0);
                tmem_st(ta + 2 * (D_LV0 + j), is_free(j) ? nl[3 + j] + cv[j] : 0.0);
              }
            }
            interval_done = true;
          } else {
            if (!isD) {
              double k6[3];
              tm_ld<3>(ta + 2 * (P_KV + 18), k6);
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                tmem_st(ta + 2 * (P_U0 + j), y1a[j]); tmem_st(ta + 2 * (P_V0 + j), y1b[j]);
                tmem_st(ta + 2 * (P_KV + j), k6[j]);
              }
            } else {
              double k6[3], m6[3];
              tm_ld<3>(ta + 2 * (D_KLU + 18), k6);
              tm_ld<3>(ta + 2 * (D_KLV + 18), m6);
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                tmem_st(ta + 2 * (D_LU0 + j), y1a[j]); tmem_st(ta + 2 * (D_LV0 + j), y1b[j]);
                tmem_st(ta + 2 * (D_KLU + j), k6[j]); tmem_st(ta + 2 * (D_KLV + j), m6[j]);
              }
            }
            if (tid < NSLOT) accb[tid] = accb[NSLOT + tid];  // k1 <- k7 of every scalar running sum
          }
          tmem_st_wait();
          if (tid < NSCAL) Sq0[tid] = Sqnew[tid];  // (written above by lane 0 of warp tid; block_sum's barriers are in between)
          par ^= 1;  // q0 <- qnew, k1 <- k7 for every thread-private quadrature
          s_new = s_cur + h;
        }
        const double dfactor = ratio < 1.0 ? 1.0 : 0.2;
        const double factor = fmin(10.0, fmax(inv_fifth_root(ratio) * 0.9, dfactor));
        h_new = (ratio == 0.0) ? h * 10.0 : h * factor;
        if (!interval_done) {
          if (!(h_new > 0.0)) { status = DFX_STATUS_DT_UNDERFLOW; stop = true; }
          else if (istep >= a.max_steps) { status = DFX_STATUS_MAX_STEPS; stop = true; }
        }
      }
      if (tid == 0) {
        C->n_steps += 1; C->istep = istep; C->n_rhs += 6; C->status |= status;
        if (accepted) C->n_acc += 1;
        C->h = h_new; C->s_cur = s_new;
        if (!interval_done && !stop) { const double s_nn = s_new + h_new; C->crossing = !(s_nn < s_target); C->x = (s_target - s_new) / (s_nn - s_new); }
      }
      if (stop) running = false;
      else if (interval_done) { if (--i < 1) running = false; else ev = EV_INIT; }
      else ev = 0;
      if (lane == 0) C->ev_w[warp] = running ? ev : EV_STOP;
      if (tid == 0) { C->hst = h_new; C->i = i; C->par = par; }
      PT_MARK(8);  // accept / reject bookkeeping
      if (drive_on && running && warp == NW - 1) {
        if (ev == 0 && lane < 6) fill_drive(lane, -(s_new + h_new * tab.alpha[lane]));
        else if (ev == EV_INIT && lane == EV_INIT) fill_drive(EV_INIT, ts[i]);
      }
      __syncthreads();
      PT_MARK(9);  // drive table of the next step (last warp) + closing barrier
    }
  }

#ifdef DFX_PHASE_TIMERS
  if (design == 0 && lane == 0 && (warp == 0 || warp == 4 || warp == 12 || warp == 15 || warp == 23)) {
    const double n_ = (double)C->n_rhs;
    printf("warp %2d cycles/eval: A %.0f | wait A->B %.0f | B %.0f | wait B->C %.0f | C %.0f | post %.0f | total %.0f | evaluations %.0f\n", warp,
           pt_acc[0] / n_, pt_acc[1] / n_, pt_acc[2] / n_, pt_acc[3] / n_, pt_acc[4] / n_, pt_acc[5] / n_,
           (pt_acc[0] + pt_acc[1] + pt_acc[2] + pt_acc[3] + pt_acc[4] + pt_acc[5] + pt_acc[6] + pt_acc[7] + pt_acc[8] + pt_acc[9]) / n_, n_);
    printf("warp %2d step end, cycles/step: own error terms %.0f | barrier + scalar totals + norm %.0f | accept / reject %.0f | drive table + barrier %.0f | steps %lld\n",
           warp, (double)pt_acc[6] / C->n_steps, (double)pt_acc[7] / C->n_steps, (double)pt_acc[8] / C->n_steps, (double)pt_acc[9] / C->n_steps, C->n_steps);
  }
#endif
  // ---- outputs ----------------------------------------------------------------------------------------------
  __syncthreads();
  const double nanv = nan("");
  const int status = C->status, par = C->par;
  const bool bad = status != 0;
  {
    double qv[NE3];
#pragma unroll
    for (int e = 0; e < NE3; ++e) qv[e] = bad ? nanv : __ldcg(&qg[(long long)(par * NE3 + e) * TT]);
    if (!isD) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (DAMP == 2 && has_unit && a.grads.damping && ((flags >> (6 + j)) & 1u))
          a.grads.damping[(long long)design * T.n_damped * 3 + T.damp_slot[3 * blk + j]] = is_free(j) ? qv[3 + j] : (bad ? nanv : 0.0);
        if (is_free(j) && a.grads.inertia) a.grads.inertia[(long long)design * nf + fidx(j)] = qv[j];
      }
      if (has_unit && a.grads.centroid_node_vectors)
#pragma unroll
        for (int l = 0; l < NPB; ++l) a.grads.centroid_node_vectors[((long long)design * T.n_nodes + blk * NPB + l) * 2] = qv[6 + l];
    } else {
      double y0[6];
      tm_ld<6>(ta + 2 * D_LU0, y0);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (is_free(j) && a.y0_bar) {
          a.y0_bar[(long long)design * 2 * nf + fidx(j)] = bad ? nanv : y0[j];
          a.y0_bar[(long long)design * 2 * nf + nf + fidx(j)] = bad ? nanv : y0[3 + j];
        }
      }
      if (has_unit && a.grads.centroid_node_vectors)
#pragma unroll
        for (int l = 0; l < NPB; ++l) {
          const long long n = (long long)design * T.n_nodes + blk * NPB + l;
          a.grads.centroid_node_vectors[n * 2 + 1] = qv[l];
        }
    }
    if (has_bnd && a.grads.reference_vector) {
      a.grads.reference_vector[((long long)design * NBONDS + bnd) * 2] = qv[10];
      a.grads.reference_vector[((long long)design * NBONDS + bnd) * 2 + 1] = qv[11];
    }
  }
  if (tid == 0) {
    if (a.grads.k_stretch) a.grads.k_stretch[design] = bad ? nanv : Sq0[SC_KS];
    if (a.grads.k_shear) a.grads.k_shear[design] = bad ? nanv : Sq0[SC_KSH];
    if (a.grads.k_rot) a.grads.k_rot[design] = bad ? nanv : Sq0[SC_KR];
    if (a.grads.damping && DAMP == 1) a.grads.damping[design] = bad ? nanv : Sq0[SC_DAMP];
    if (a.grads.contact && CONTACT) for (int k = 0; k < 3; ++k) a.grads.contact[(long long)design * 3 + k] = bad ? nanv : Sq0[SC_CONTACT + k];
    if (a.grads.drive) for (int k = 0; k < T.n_drive_params; ++k) a.grads.drive[(long long)design * T.n_drive_params + k] = bad ? nanv : Sq0[SC_DRIVE + k];
    if (a.ts_bar) a.ts_bar[(long long)design * a.n_t] = bad ? nanv : Sq0[SC_T0];
    if (a.stats) {
      DfxStats st;
      st.steps = C->n_steps; st.accepted = C->n_acc; st.rhs_evals = C->n_rhs; st.status = status; st.reserved = 0; st.last_dt = C->h;
      a.stats[design] = st;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

}  // namespace dfx
