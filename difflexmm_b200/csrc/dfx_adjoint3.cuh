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
constexpr int P_KV = 0, P_U0 = 21, P_V0 = 24, P_TV = 27, P_N = 30;
constexpr int D_KLU = 0, D_KLV = 21, D_LU0 = 42, D_LV0 = 45, D_TLU = 48, D_N = 51;
constexpr int P_COLS = 2 * P_N, D_COLS = 2 * D_N;  // 60 + 102 columns per (P, D) warp pair, three pairs per lane quarter
static_assert(3 * (P_COLS + D_COLS) <= 512, "tensor memory columns");
constexpr int NBC = 10;  // bond constants: r0x r0y L0 1/L0 r1x r1y r2x r2y da1 da2
constexpr int NE3 = 10;  // quadrature entries per thread: 0..7 unit role (P: inertia[3] damping[3]; D: cnv x[4] y[4]), 8..9 bond role
// L2 scratch of one SM (doubles); every array is [..][TT], a thread touches its own column only
constexpr long long G_BC = 0;
constexpr long long G_SPART = G_BC + (long long)NBC * TT;   // [6][TT] partials of k_stretch k_shear k_rot c_min c_cut k_c
constexpr long long G_BQ = G_SPART + 6LL * TT;              // [sol, err][2][TT] running sums of the reference-vector quadratures
constexpr long long G_Q = G_BQ + 4LL * TT;                  // [q0 a, q0 b, k1 a, k1 b][NE3][TT]
constexpr long long G_TOTAL = G_Q + 4LL * NE3 * TT;
// scalar leaves: running sums [k1, k7, sol, err, mid][NSLOT]
constexpr int NS_DENSE = 6;                                  // k_stretch k_shear k_rot c_min c_cut k_c: one total each
constexpr int SLOT_T = NS_DENSE;                             // t0_bar, drive[5]: one partial per D warp
constexpr int SLOT_DAMP = SLOT_T + 6 * NDW;                  // scalar damping leaf: one partial per P warp
constexpr int NSLOT = SLOT_DAMP + NDW;

struct Ctrl {
  double h, h0, d1, s0, s_target, s_cur, x;
  long long n_steps, n_acc, n_rhs, istep;
  int status, crossing, contact_seen;
  uint32_t tmem_base;
};

template <int NSL>
struct Lay {  // shared-memory layout (doubles)
  static constexpr int RED = 0, US = 40, WS = US + 5 * TU, SL = WS + 6 * TU, INVM = SL + NSL * TT, CD = INVM + 3 * TU,
                       QSP = CD + 3 * TU, QSD = QSP + 12 * TU, SQ = QSD + 16 * TU, ACC = SQ + 2 * NSCAL, DRV = ACC + 5 * NSLOT,
                       CTRL = DRV + 32, END = CTRL + (int)((sizeof(Ctrl) + 7) / 8);
};

template <int N>
__device__ __forceinline__ void tm_ld(uint32_t taddr, double (&out)[N]) {  // N consecutive slots, one wait
  uint32_t lo[N], hi[N];
#pragma unroll
  for (int i = 0; i < N; ++i) tmem_ld_issue(taddr + 2 * i, lo[i], hi[i]);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < N; ++i) { tmem_pin(lo[i], hi[i]); out[i] = __hiloint2double(hi[i], lo[i]); }
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
  uint32_t lo[2 * N], hi[2 * N];
#pragma unroll
  for (int q = 0; q < N; ++q) {
    tmem_ld_issue(ta + 2 * (D_KLU + 3 * L0 + q), lo[q], hi[q]);
    tmem_ld_issue(ta + 2 * (D_KLV + 3 * L0 + q), lo[N + q], hi[N + q]);
  }
  tmem_ld_wait();
#pragma unroll
  for (int q = 0; q < 2 * N; ++q) tmem_pin(lo[q], hi[q]);
#pragma unroll
  for (int l = L0; l <= L1; ++l) {
    const double b = tab.beta[ST][l];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int q = 3 * (l - L0) + j;
      alu[j] = fma(b, __hiloint2double(hi[q], lo[q]), alu[j]);
      alv[j] = fma(b, __hiloint2double(hi[N + q], lo[N + q]), alv[j]);
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

// What one quadrature entry needs from the step controller
struct QC {
  int mode;  // 0: k1 at interval start | 1: nothing | 2..5: accumulating stage | 6: last stage | 7: initial-step probe
  int par;   // which copy of q0 / k1 is current
  bool crossing;
  double h, x, atol, rtol;
  double cs, ce, cm, cs0, ce0, cm0;
};

// N quadrature entries of one thread receive their integrand values.  `qg` = the thread's column of G_Q, entry e0 + k;
// running solution / error sums at sol[k * sstride], err[k * sstride] (shared memory for the unit role, scratch for the
// bond role).  Returns the thread's contribution to the error norm (mode 6) or to d2 of initial_step_size (mode 7).
template <int N>
__device__ __forceinline__ double quad_entries(const QC& c, double* qg, int e0, double* sol, double* err, int sstride,
                                               const double (&val)[N]) {
  double acc = 0.0;
  double* q0p = qg + (long long)(c.par * NE3 + e0) * TT;
  double* qnp = qg + (long long)((1 - c.par) * NE3 + e0) * TT;
  double* k1p = qg + (long long)((2 + c.par) * NE3 + e0) * TT;
  double* k7p = qg + (long long)((3 - c.par) * NE3 + e0) * TT;  // spare copy: k7, or the midpoint sum on a crossing step
  const int mode = c.mode;
  if (mode >= 3 && mode <= 5) {
    double s_in[N], e_in[N];
#pragma unroll
    for (int k = 0; k < N; ++k) { s_in[k] = sol[k * sstride]; e_in[k] = err[k * sstride]; }
#pragma unroll
    for (int k = 0; k < N; ++k) { sol[k * sstride] = fma(c.cs, val[k], s_in[k]); err[k * sstride] = fma(c.ce, val[k], e_in[k]); }
    if (c.crossing) {
      double m_in[N];
#pragma unroll
      for (int k = 0; k < N; ++k) m_in[k] = k7p[k * TT];
#pragma unroll
      for (int k = 0; k < N; ++k) k7p[k * TT] = fma(c.cm, val[k], m_in[k]);
    }
  } else if (mode == 2) {
    double k1[N];
#pragma unroll
    for (int k = 0; k < N; ++k) k1[k] = k1p[k * TT];
#pragma unroll
    for (int k = 0; k < N; ++k) {
      sol[k * sstride] = fma(c.cs, val[k], c.cs0 * k1[k]);
      err[k * sstride] = fma(c.ce, val[k], c.ce0 * k1[k]);
      if (c.crossing) k7p[k * TT] = fma(c.cm, val[k], c.cm0 * k1[k]);
    }
  } else if (mode == 6) {
    double q_in[N], s_in[N], e_in[N];
#pragma unroll
    for (int k = 0; k < N; ++k) { q_in[k] = q0p[k * TT]; s_in[k] = sol[k * sstride]; e_in[k] = err[k * sstride]; }
    if (!c.crossing) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const double q1 = fma(c.h, s_in[k], q_in[k]);
        const double r = c.h * fma(c.ce, val[k], e_in[k]) * rcp_pos(c.atol + c.rtol * fmax(fabs(q_in[k]), fabs(q1)));
        acc = fma(r, r, acc);
        qnp[k * TT] = q1;
        k7p[k * TT] = val[k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const double q1 = fma(c.h, s_in[k], q_in[k]);
        const double r = c.h * fma(c.ce, val[k], e_in[k]) * rcp_pos(c.atol + c.rtol * fmax(fabs(q_in[k]), fabs(q1)));
        acc = fma(r, r, acc);
        const double amid = fma(c.cm, val[k], k7p[k * TT]);
        qnp[k * TT] = interp_eval(q_in[k], q1, q_in[k] + c.h * amid, c.h * k1p[k * TT], c.h * val[k], c.x);
      }
    }
  } else if (mode == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) k1p[k * TT] = val[k];
  } else if (mode == 7) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const double d = (val[k] - k1p[k * TT]) * rcp_pos(c.atol + fabs(q0p[k * TT]) * c.rtol);
      acc = fma(d, d, acc);
    }
  }
  return acc;
}

// one running-sum slot of a scalar leaf receives the total (or a warp's partial) of this evaluation
__device__ __forceinline__ void scal_slot(const QC& c, double* accb, int slot, double v) {
  double* k1 = accb + slot; double* k7 = k1 + NSLOT; double* sol = k7 + NSLOT; double* err = sol + NSLOT; double* mid = err + NSLOT;
  switch (c.mode) {
    case 0: *k1 = v; break;
    case 7: *k7 = v; break;
    case 2: { const double a = *k1; *sol = c.cs0 * a + c.cs * v; *err = c.ce0 * a + c.ce * v; *mid = c.cm0 * a + c.cm * v; } break;
    case 6: *err += c.ce * v; *mid += c.cm * v; *k7 = v; break;
    default: *sol += c.cs * v; *err += c.ce * v; *mid += c.cm * v; break;
  }
}
// total of scalar leaf `which` (NSCAL numbering of dfx_adjoint.cuh) in running-sum array m (0 k1, 1 k7, 2 sol, 3 err, 4 mid)
__device__ __forceinline__ double scal_total(const double* accb, int m, int which) {
  const double* p = accb + m * NSLOT;
  if (which >= SC_KS && which <= SC_KR) return p[which - SC_KS];
  if (which >= SC_CONTACT && which < SC_CONTACT + 3) return p[3 + which - SC_CONTACT];
  int s0 = SLOT_DAMP;
  if (which == SC_T0) s0 = SLOT_T;
  else if (which >= SC_DRIVE) s0 = SLOT_T + (1 + which - SC_DRIVE) * NDW;
  double s = 0.0;
#pragma unroll 1
  for (int w = 0; w < NDW; ++w) s += p[s0 + w];
  return s;
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
  const int design = blockIdx.x;
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
  const int ndp = T.n_drive_params;

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
  int nbq[NPB];  // bond * 2 + side attached to each vertex of this thread's unit (-1: none)
#pragma unroll
  for (int l = 0; l < NPB; ++l) nbq[l] = has_unit ? A.node_bond[blk * NPB + l] : -1;
  int bbp;  // the two blocks of this thread's bond, packed
  { const int2 bb = T.bond_blocks[bnd]; bbp = bb.x | (bb.y << 16); }

  // ---- constants ----------------------------------------------------------------------------------------------
  for (int i = tid; i < NSL * TT; i += TT) SL[i] = 0.0;
  for (int i = tid; i < 28 * TU; i += TT) smem[L::QSP + i] = 0.0;
  for (int i = tid; i < 2 * NSCAL + 5 * NSLOT + 32; i += TT) Sq0[i] = 0.0;
  if (tid == 0) {
    drv[30] = nan(""); drv[31] = nan("");
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
  for (int q = 0; q < 6 + 4 + 4 * NE3; ++q) gcol[G_SPART + (long long)q * TT] = 0.0;
  double cmin = 0, ccut = 0, ckc = 0;
  if (CONTACT) { const double* g_contact = leaf_ptr(a.p.contact, design); cmin = g_contact[0]; ccut = g_contact[1]; ckc = g_contact[2]; }
  // y_bar = g[-1]
  if (isD) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      tmem_st(ta + 2 * (D_LU0 + j), is_free(j) ? cotangent_nl(a, design, a.n_t - 1, fidx(j), false) : 0.0);
      tmem_st(ta + 2 * (D_LV0 + j), is_free(j) ? cotangent_nl(a, design, a.n_t - 1, fidx(j), true) : 0.0);
    }
    tmem_st_wait();
  }
  __syncthreads();

  const double inv_n = 1.0 / (double)a.aug_size;
  constexpr int EV_INIT = 6, EV_PROBE = 7;
  int ev = EV_INIT, i = a.n_t - 1, par = 0;
  unsigned nev = 0;  // evaluation counter: selects the Ws buffer
  double hst = 0.0;  // step size of the stage about to be evaluated (h0 for the probe)
  bool running = i >= 1;

  while (running) {
    // ================= stage state (phase A) =================
    QC qc;
    qc.par = par; qc.atol = atol; qc.rtol = rtol;
    const double s_cur = C->s_cur;
    double time;
    int kidx;
    if (ev == EV_INIT) { time = ts[i]; kidx = 0; qc.mode = 0; }
    else if (ev == EV_PROBE) { time = -(C->s0 + hst); kidx = 1; qc.mode = 7; }
    else { kidx = ev + 1; time = -(s_cur + hst * tab.alpha[ev]); qc.mode = kidx; }
    qc.h = hst; qc.crossing = C->crossing != 0; qc.x = C->x;
    qc.cs = tab.c_sol[kidx]; qc.ce = tab.c_err[kidx]; qc.cm = tab.c_mid[kidx];
    qc.cs0 = tab.c_sol[0]; qc.ce0 = tab.c_err[0]; qc.cm0 = tab.c_mid[0];
    const bool want_q = qc.mode != 1;
    const double time_next = ev < 5 ? -(s_cur + hst * tab.alpha[ev + 1]) : nan("");
    double* Wcur = Ws + (nev & 1u) * 3 * TU;
    ++nev;
    // bond constants: fetched from L2 now, used after the barrier
    double bc[NBC];
#pragma unroll
    for (int k = 0; k < NBC; ++k) bc[k] = gcol[G_BC + (long long)k * TT];

    if (!isD) {
      double us[3], vs[3];
      if (ev == EV_INIT) {
        const double* yi = ys + (long long)i * 2 * nf;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          us[j] = is_free(j) ? __ldcs(&yi[fidx(j)]) : 0.0;
          vs[j] = is_free(j) ? __ldcs(&yi[nf + fidx(j)]) : 0.0;
          tmem_st(ta + 2 * (P_U0 + j), us[j]); tmem_st(ta + 2 * (P_V0 + j), vs[j]);
        }
      } else if (ev == EV_PROBE) {
        double y0[6], k0[3];
        tm_ld<6>(ta + 2 * P_U0, y0);
        tm_ld<3>(ta + 2 * P_KV, k0);
#pragma unroll
        for (int j = 0; j < 3; ++j) { us[j] = y0[j] - hst * y0[3 + j]; vs[j] = y0[3 + j] + hst * k0[j]; }
      } else {
        switch (ev) {
          case 0: stage_P<0>(ta, tab, hst, us, vs); break;
          case 1: stage_P<1>(ta, tab, hst, us, vs); break;
          case 2: stage_P<2>(ta, tab, hst, us, vs); break;
          case 3: stage_P<3>(ta, tab, hst, us, vs); break;
          case 4: stage_P<4>(ta, tab, hst, us, vs); break;
          default: stage_P<5>(ta, tab, hst, us, vs); break;
        }
      }
      if (has_cons && T.drive_kind != DFX_DRIVE_ZERO) {
        double s0_, s1_;
        if (drv[30] == time) { s0_ = drv[28]; s1_ = drv[29]; }
        else {
          DriveEval de;
          drive_eval(T.drive_kind, time, g_drive, false, de, T.table);
          s0_ = de.s[0]; s1_ = de.s[1];
        }
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (is_cons(j)) { const int cs_ = T.cons_slot[3 * blk + j]; us[j] = T.drive_vec0[cs_] * s0_ + T.drive_vec1[cs_] * s1_; }
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (!is_free(j)) { vs[j] = 0.0; if (!is_cons(j)) us[j] = 0.0; }
        tmem_st(ta + 2 * (P_TV + j), vs[j]);
      }
      if (has_unit) {
        double sn, cs;
        sincos(us[2], &sn, &cs);
        Us[tid] = us[0]; Us[TU + tid] = us[1]; Us[2 * TU + tid] = us[2]; Us[3 * TU + tid] = sn; Us[4 * TU + tid] = cs;
      }
    } else {
      double lus[3], lvs[3];
      if (ev == EV_INIT) {
        double y0[6];
        tm_ld<6>(ta + 2 * D_LU0, y0);
#pragma unroll
        for (int j = 0; j < 3; ++j) { lus[j] = y0[j]; lvs[j] = y0[3 + j]; }
      } else if (ev == EV_PROBE) {
        double y0[6], k0[3], k1[3];
        tm_ld<6>(ta + 2 * D_LU0, y0);
        tm_ld<3>(ta + 2 * D_KLU, k0);
        tm_ld<3>(ta + 2 * D_KLV, k1);
#pragma unroll
        for (int j = 0; j < 3; ++j) { lus[j] = y0[j] + hst * k0[j]; lvs[j] = y0[3 + j] + hst * k1[j]; }
      } else {
        switch (ev) {
          case 0: stage_D<0>(ta, tab, hst, lus, lvs); break;
          case 1: stage_D<1>(ta, tab, hst, lus, lvs); break;
          case 2: stage_D<2>(ta, tab, hst, lus, lvs); break;
          case 3: stage_D<3>(ta, tab, hst, lus, lvs); break;
          case 4: stage_D<4>(ta, tab, hst, lus, lvs); break;
          default: stage_D<5>(ta, tab, hst, lus, lvs); break;
        }
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (!is_free(j)) { lus[j] = 0.0; lvs[j] = 0.0; }
        tmem_st(ta + 2 * (D_TLU + j), lus[j]);
        if (has_unit) Wcur[j * TU + unit] = lvs[j] * INVM[j * TU + unit];
      }
    }
    tmem_st_wait();
    __syncthreads();

    // ================= phase B: this thread's bond =================
    double accq = 0.0;  // this thread's contribution to the error norm / probe norm
    if (tid == TT - 1 && T.drive_kind != DFX_DRIVE_ZERO) {  // drive channels: now (with derivatives) and next
      DriveEval de;
      drive_eval(T.drive_kind, time, g_drive, true, de, T.table);
      drv[2] = de.sdot[0]; drv[3] = de.sdot[1];
#pragma unroll
      for (int q = 0; q < DFX_MAX_DRIVE_PARAMS; ++q) { drv[4 + q] = de.dsdp[0][q]; drv[9 + q] = de.dsdp[1][q]; }
      drive_eval(T.drive_kind, time_next, g_drive, false, de, T.table);
      drv[28] = de.s[0]; drv[29] = de.s[1]; drv[30] = time_next;
    }
    {
      double p_c0 = 0, p_c1 = 0, p_c2 = 0;
      bool act = false;
      if (has_bnd) {
        const int b1 = bbp & 0xffff, b2 = bbp >> 16;
        BlockState<Dual> s1, s2;
        make_block(Us[b1], Us[TU + b1], Us[2 * TU + b1], Us[3 * TU + b1], Us[4 * TU + b1], Wcur[b1], Wcur[TU + b1], Wcur[2 * TU + b1], s1);
        make_block(Us[b2], Us[TU + b2], Us[2 * TU + b2], Us[3 * TU + b2], Us[4 * TU + b2], Wcur[b2], Wcur[TU + b2], Wcur[2 * TU + b2], s2);
        const double* g_ks = leaf_ptr(a.p.k_stretch, design);
        const double* g_ksh = leaf_ptr(a.p.k_shear, design);
        const double* g_kr = leaf_ptr(a.p.k_rot, design);
        BondConst bcs = {bc[0], bc[1], bc[2], bc[3]};
        BondOut<Dual> o;
        bond_gradient<Dual, true>(DFX_BOND_LIGAMENT, s1, s2, bc[4], bc[5], bc[6], bc[7], bcs, g_ks[0], g_ksh[0], g_kr[0], o);
        double a1 = 0.0, a2 = 0.0;
        if (CONTACT) {
          Dual psi1 = wrapT(s1.th - s2.th + bc[8]);
          Dual psi2 = wrapT(s2.th - s1.th + bc[9]);
          const bool act1 = !(psi1.v < cmin) && psi1.v < ccut, act2 = !(psi2.v < cmin) && psi2.v < ccut;
          act = act1 || act2;
          if (act) {
            Dual e1, e2, m1, m2, c1, c2, k1, k2;
            contact_term<Dual>(psi1, cmin, ccut, ckc, e1, m1, c1, k1);
            contact_term<Dual>(psi2, cmin, ccut, ckc, e2, m2, c2, k2);
            o.f1[2] = o.f1[2] + e1 - e2;
            o.f2[2] = o.f2[2] + e2 - e1;
            a1 = e1.d; a2 = e2.d;
            p_c0 = -(m1.d + m2.d); p_c1 = -(c1.d + c2.d); p_c2 = -(k1.d + k2.d);
            if (!(flags & 512u)) { flags |= 512u; C->contact_seen = 1; }
          }
        }
        const int b = tid;
        // forces on the two ends are equal and opposite: store (gdx, gdy) once, the two torques separately
        SL[b] = o.f2[0].v; SL[TT + b] = o.f2[1].v; SL[2 * TT + b] = -o.f1[2].v; SL[3 * TT + b] = -o.f2[2].v;
        SL[4 * TT + b] = o.f2[0].d; SL[5 * TT + b] = o.f2[1].d; SL[6 * TT + b] = o.f1[2].d; SL[7 * TT + b] = o.f2[2].d;
        SL[8 * TT + b] = -o.gr1[0].d; SL[9 * TT + b] = -o.gr1[1].d;
        SL[10 * TT + b] = -o.gr2[0].d; SL[11 * TT + b] = -o.gr2[1].d;
        if (CONTACT && (flags & 512u)) { SL[12 * TT + b] = a1; SL[13 * TT + b] = a2; }
        if (want_q) {
          // d(w.F)/dp = -(dual part of dE/dp)
          const double qb[2] = {-o.gr0[0].d, -o.gr0[1].d};
          accq = quad_entries<2>(qc, qg, 8, gcol + G_BQ, gcol + G_BQ + 2 * TT, TT, qb);
          double* sp = gcol + G_SPART;
          sp[0] = -o.gks.d; sp[TT] = -o.gksh.d; sp[2 * TT] = -o.gkr.d;
          if (CONTACT && (flags & 512u)) { sp[3 * TT] = p_c0; sp[4 * TT] = p_c1; sp[5 * TT] = p_c2; }
        }
      }
    }
    __syncthreads();

    // ================= phase C: this thread's half of its unit =================
    if (!isD) {
      double F[3] = {0, 0, 0};
#pragma unroll
      for (int l = 0; l < NPB; ++l) {
        const int nb_ = nbq[l];
        if (nb_ >= 0) {
          const int b = nb_ >> 1;
          const bool second = nb_ & 1;
          const double sg = second ? -1.0 : 1.0;
          F[0] += sg * SL[b]; F[1] += sg * SL[TT + b]; F[2] += SL[(second ? 3 : 2) * TT + b];
        }
      }
      double vst[3];
      tm_ld<3>(ta + 2 * P_TV, vst);
      double val[6];
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
        tmem_st(ta + 2 * (P_KV + kidx * 3 + j), kvv);
      }
      if (want_q) {
        double* qs = smem + L::QSP + tid;
        if (DAMP == 2) accq += quad_entries<6>(qc, qg, 0, qs, qs + 6 * TU, TU, val);
        else { const double v3[3] = {val[0], val[1], val[2]}; accq += quad_entries<3>(qc, qg, 0, qs, qs + 6 * TU, TU, v3); }
        if (DAMP == 1) { const double tot = warp_sum(p_damp); if (lane == 0) scal_slot(qc, accb, SLOT_DAMP + warp, tot); }
        // dense scalar leaves: warp k sums the 768 per-bond partials of leaf k (fixed order)
        if (warp < (CONTACT ? 6 : 3) && (warp < 3 || C->contact_seen)) {
          const double* sp = gbase + G_SPART + (long long)warp * TT + lane;
          double s = 0.0;
#pragma unroll
          for (int q0_ = 0; q0_ < NW; q0_ += 8) {
            double v8[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v8[q] = __ldcg(&sp[(q0_ + q) * 32]);
#pragma unroll
            for (int q = 0; q < 8; ++q) s += v8[q];
          }
          s = warp_sum(s);
          if (lane == 0) scal_slot(qc, accb, warp, s);
        }
      }
    } else {
      double HW[3] = {0, 0, 0}, val[8], An[NPB], Ap[NPB];
      const bool seen = CONTACT && C->contact_seen;
#pragma unroll
      for (int l = 0; l < 4; ++l) { val[l] = 0.0; val[4 + l] = 0.0; }
#pragma unroll
      for (int l = 0; l < NPB; ++l) {
        An[l] = 0.0; Ap[l] = 0.0;
        const int nb_ = nbq[l];
        if (nb_ >= 0) {
          const int b = nb_ >> 1;
          const bool second = nb_ & 1;
          const double sg = second ? -1.0 : 1.0;
          HW[0] -= sg * SL[4 * TT + b]; HW[1] -= sg * SL[5 * TT + b]; HW[2] += SL[(second ? 7 : 6) * TT + b];
          val[l] = SL[(second ? 10 : 8) * TT + b]; val[4 + l] = SL[(second ? 11 : 9) * TT + b];
          if (seen) {
            // dS/dalpha = -(dual part of dE/dalpha): a1next:+e1, a1prev:-e2, a2next:+e2, a2prev:-e1
            const double e1d = SL[12 * TT + b], e2d = SL[13 * TT + b];
            An[l] = second ? -e2d : -e1d;
            Ap[l] = second ? e1d : e2d;
          }
        }
      }
      double lust[3];
      tm_ld<3>(ta + 2 * D_TLU, lust);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double kluv = 0.0, klvv = 0.0;
        if (is_free(j)) { kluv = -HW[j]; klvv = lust[j] - CDs[j * TU + unit] * Wcur[j * TU + unit]; }
        tmem_st(ta + 2 * (D_KLU + kidx * 3 + j), kluv);
        tmem_st(ta + 2 * (D_KLV + kidx * 3 + j), klvv);
      }
      if (want_q) {
        if (seen) {
          bool any_contact = false;
#pragma unroll
          for (int l = 0; l < NPB; ++l) any_contact |= (An[l] != 0.0) | (Ap[l] != 0.0);
          if (any_contact) {
            // contact chain of the centroid_node_vectors cotangent: edge l -> l+1 is node l's "next" edge and,
            // reversed, node (l+1)'s "previous" edge; both angles have the same derivative w.r.t. the end points
#pragma unroll
            for (int l = 0; l < NPB; ++l) {
              const int ln = l + 1 == NPB ? 0 : l + 1;
              const int n = blk * NPB + l, m = blk * NPB + ln;
              const double ex = g_cnv[2 * m] - g_cnv[2 * n], ey = g_cnv[2 * m + 1] - g_cnv[2 * n + 1];
              const double inv = 1.0 / (ex * ex + ey * ey);
              const double wx = -(An[l] + Ap[ln]) * ey * inv, wy = (An[l] + Ap[ln]) * ex * inv;
              val[ln] += wx; val[4 + ln] += wy;
              val[l] -= wx; val[4 + l] -= wy;
            }
          }
        }
        double* qs = smem + L::QSD + unit;
        accq += quad_entries<8>(qc, qg, 0, qs, qs + 8 * TU, TU, val);
        // t0_bar and the drive parameters only receive contributions from constrained DOFs
        if (warp_t0 && T.drive_kind != DFX_DRIVE_ZERO) {
          double p[6] = {0, 0, 0, 0, 0, 0};
          if (has_cons) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              if (is_cons(j)) {
                const int cs_ = T.cons_slot[3 * blk + j];
                const double v0_ = T.drive_vec0[cs_], v1_ = T.drive_vec1[cs_];
                p[0] -= HW[j] * (v0_ * drv[2] + v1_ * drv[3]);
#pragma unroll
                for (int q = 0; q < DFX_MAX_DRIVE_PARAMS; ++q) p[1 + q] -= HW[j] * (v0_ * drv[4 + q] + v1_ * drv[9 + q]);
              }
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int q = 0; q < 6; ++q) p[q] += __shfl_xor_sync(0xffffffffu, p[q], o);
          }
          if (lane <= ndp) {
            double mine = p[0];
#pragma unroll
            for (int q = 1; q < 6; ++q) if (lane == q) mine = p[q];
            scal_slot(qc, accb, SLOT_T + lane * NDW + (warp - NDW), mine);
          }
        }
      }
    }
    tmem_st_wait();

    // ================= what follows the evaluation =================
    if (ev == EV_INIT) {
      double sd0 = 0, sd1 = 0, pt = 0.0;
      if (!isD) {
        double y0[6], k0[3];
        tm_ld<6>(ta + 2 * P_U0, y0);
        tm_ld<3>(ta + 2 * P_KV, k0);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (is_free(j)) {
            // t_bar = func(ys[i], ts[i]) . g[i] with func = (v0, -kv[0])
            pt += y0[3 + j] * cotangent_nl(a, design, i, fidx(j), false) - k0[j] * cotangent_nl(a, design, i, fidx(j), true);
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
#pragma unroll 1
      for (int e = 0; e < NE3; ++e) {
        const double q0 = qg[(long long)(par * NE3 + e) * TT], k1 = qg[(long long)((2 + par) * NE3 + e) * TT];
        const double s = atol + fabs(q0) * rtol;
        sd0 += (q0 / s) * (q0 / s); sd1 += (k1 / s) * (k1 / s);
      }
      if (tid < NSCAL) {
        const double s = atol + fabs(Sq0[tid]) * rtol;
        const double a0 = Sq0[tid] / s, b0 = scal_total(accb, 0, tid) / s;
        sd0 += a0 * a0; sd1 += b0 * b0;
      }
      const double d0 = sqrt(block_sum(sd0, red));
      const double d1 = sqrt(block_sum(sd1, red));
      const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
      if (tid == 0) { C->s0 = -ts[i]; C->s_target = -ts[i - 1]; C->h0 = h0; C->d1 = d1; C->n_rhs += 1; C->crossing = 0; }
      hst = h0;
      ev = EV_PROBE;
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
      hst = h;
      if (empty) { if (--i < 1) running = false; else ev = EV_INIT; }
      else if (stop) running = false;
      else ev = 0;
      __syncthreads();
    } else if (ev < 5) {
      ++ev;
    } else {
      double se = accq;
      const double h = hst;
      const bool crossing = qc.crossing;
      double y1a[3], y1b[3];  // P: u, v at the end of the step; D: lambda_u, lambda_v
      if (!isD) {
        double y0[6];
        tm_ld<6>(ta + 2 * P_U0, y0);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          double kv[7];
          tm_lds<7>(ta + 2 * (P_KV + j), 3, kv);
          double eu = 0.0, evv = 0.0, su = 0.0, sv = 0.0;
#pragma unroll
          for (int l = 0; l < 7; ++l) {
            eu = fma(tab.e2[l], kv[l], eu); evv = fma(tab.c_err[l], kv[l], evv);
            su = fma(tab.s2[l], kv[l], su); sv = fma(tab.c_sol[l], kv[l], sv);
          }
          y1a[j] = y0[j] - h * (tab.sum_sol * y0[3 + j] + h * su);
          y1b[j] = y0[3 + j] + h * sv;
          eu = -h * (tab.sum_err * y0[3 + j] + h * eu);
          evv *= h;
          if (is_free(j)) {
            const double r0 = eu / (atol + rtol * fmax(fabs(y0[j]), fabs(y1a[j])));
            const double r1 = evv / (atol + rtol * fmax(fabs(y0[3 + j]), fabs(y1b[j])));
            se += r0 * r0 + r1 * r1;
          }
        }
      } else {
        double y0[6];
        tm_ld<6>(ta + 2 * D_LU0, y0);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          double klu[7], klv[7];
          tm_lds<7>(ta + 2 * (D_KLU + j), 3, klu);
          tm_lds<7>(ta + 2 * (D_KLV + j), 3, klv);
          double elu = 0.0, elv = 0.0, slu = 0.0, slv = 0.0;
#pragma unroll
          for (int l = 0; l < 7; ++l) {
            elu = fma(tab.c_err[l], klu[l], elu); elv = fma(tab.c_err[l], klv[l], elv);
            slu = fma(tab.c_sol[l], klu[l], slu); slv = fma(tab.c_sol[l], klv[l], slv);
          }
          y1a[j] = y0[j] + h * slu;
          y1b[j] = y0[3 + j] + h * slv;
          elu *= h; elv *= h;
          if (is_free(j)) {
            const double r2 = elu / (atol + rtol * fmax(fabs(y0[j]), fabs(y1a[j])));
            const double r3 = elv / (atol + rtol * fmax(fabs(y0[3 + j]), fabs(y1b[j])));
            se += r2 * r2 + r3 * r3;
          }
        }
      }
      __syncthreads();  // scalar running sums of the last stage complete
      if (tid < NSCAL) {
        const double q0 = Sq0[tid], k1 = scal_total(accb, 0, tid), k7 = scal_total(accb, 1, tid);
        const double q1 = q0 + h * scal_total(accb, 2, tid);
        const double r = h * scal_total(accb, 3, tid) / (atol + rtol * fmax(fabs(q0), fabs(q1)));
        se += r * r;
        Sqnew[tid] = crossing ? interp_eval(q0, q1, q0 + h * scal_total(accb, 4, tid), h * k1, h * k7, qc.x) : q1;
      }
      const double ratio = sqrt(block_sum(se, red) * inv_n);
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
              double y0[6];
              tm_ld<6>(ta + 2 * D_LU0, y0);
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                double klu[7], klv[7];
                tm_lds<7>(ta + 2 * (D_KLU + j), 3, klu);
                tm_lds<7>(ta + 2 * (D_KLV + j), 3, klv);
                double mlu = 0.0, mlv = 0.0;
#pragma unroll
                for (int l = 0; l < 7; ++l) { mlu = fma(tab.c_mid[l], klu[l], mlu); mlv = fma(tab.c_mid[l], klv[l], mlv); }
                const double nlu = interp_eval(y0[j], y1a[j], y0[j] + h * mlu, h * klu[0], h * klu[6], qc.x);
                const double nlv = interp_eval(y0[3 + j], y1b[j], y0[3 + j] + h * mlv, h * klv[0], h * klv[6], qc.x);
                tmem_st(ta + 2 * (D_LU0 + j), is_free(j) ? nlu + cotangent_nl(a, design, i - 1, fidx(j), false) : 0.0);
                tmem_st(ta + 2 * (D_LV0 + j), is_free(j) ? nlv + cotangent_nl(a, design, i - 1, fidx(j), true) : 0.0);
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
          if (tid < NSCAL) Sq0[tid] = Sqnew[tid];
          par ^= 1;  // q0 <- qnew, k1 <- k7 for every thread-private quadrature
          s_new = s_cur + h;
        }
        const double dfactor = ratio < 1.0 ? 1.0 : 0.2;
        const double factor = fmin(10.0, fmax(pow(ratio, -0.2) * 0.9, dfactor));
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
      hst = h_new;
      if (stop) running = false;
      else if (interval_done) { if (--i < 1) running = false; else ev = EV_INIT; }
      else ev = 0;
      __syncthreads();
    }
  }

  // ---- outputs ----------------------------------------------------------------------------------------------
  __syncthreads();
  const double nanv = nan("");
  const int status = C->status;
  const bool bad = status != 0;
  {
    double qv[NE3];
#pragma unroll
    for (int e = 0; e < NE3; ++e) qv[e] = bad ? nanv : qg[(long long)(par * NE3 + e) * TT];
    if (!isD) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (DAMP == 2 && has_unit && a.grads.damping && ((flags >> (6 + j)) & 1u))
          a.grads.damping[(long long)design * T.n_damped * 3 + T.damp_slot[3 * blk + j]] = is_free(j) ? qv[3 + j] : (bad ? nanv : 0.0);
        if (is_free(j) && a.grads.inertia) a.grads.inertia[(long long)design * nf + fidx(j)] = qv[j];
      }
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
          a.grads.centroid_node_vectors[n * 2] = qv[l];
          a.grads.centroid_node_vectors[n * 2 + 1] = qv[4 + l];
        }
    }
    if (has_bnd && a.grads.reference_vector) {
      a.grads.reference_vector[((long long)design * NBONDS + bnd) * 2] = qv[8];
      a.grads.reference_vector[((long long)design * NBONDS + bnd) * 2 + 1] = qv[9];
    }
  }
  if (tid == 0) {
    if (a.grads.k_stretch) a.grads.k_stretch[design] = bad ? nanv : Sq0[SC_KS];
    if (a.grads.k_shear) a.grads.k_shear[design] = bad ? nanv : Sq0[SC_KSH];
    if (a.grads.k_rot) a.grads.k_rot[design] = bad ? nanv : Sq0[SC_KR];
    if (a.grads.damping && DAMP == 1) a.grads.damping[design] = bad ? nanv : Sq0[SC_DAMP];
    if (a.grads.contact && CONTACT) for (int k = 0; k < 3; ++k) a.grads.contact[(long long)design * 3 + k] = bad ? nanv : Sq0[SC_CONTACT + k];
    if (a.grads.drive) for (int k = 0; k < ndp; ++k) a.grads.drive[(long long)design * ndp + k] = bad ? nanv : Sq0[SC_DRIVE + k];
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
