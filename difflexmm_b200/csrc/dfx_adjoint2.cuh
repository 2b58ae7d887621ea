// dfx_adjoint2.cuh -- the fast adjoint kernel (same algorithm as dfx_adjoint.cuh, see there for
// the mathematics).  One CTA per design; thread t owns rigid unit t (its three DOFs, their
// cotangents, their 7-stage derivative history and every parameter-cotangent quadrature attached
// to that unit) and bonds t and t+T.  Everything a thread owns is *thread private* and lives in a
// three-tier store chosen at compile time:
//     tensor memory (TMEM, tcgen05.ld/st, 256 KB per SM, one lane per thread)  -- stage history
//     shared memory [slot][T]                                                    -- constants
//     L2-resident global scratch [slot][T] (coalesced)                           -- cold quadratures
// Only the stage state (x, y, theta, sin, cos, w) and the per-bond result slots are exchanged
// between threads, through shared memory, with two CTA barriers per RHS evaluation.  Bond results
// are stored once per bond (forces on the two ends are equal and opposite) and gathered by the
// owning unit through a node->(bond, side) table: no atomics, deterministic summation order.
//
// Preconditions (checked by the host, else the generic kernel of dfx_adjoint.cuh runs):
//   n_blocks <= T <= 512, n_bonds <= 2 T, n_npb <= 4.
#pragma once

#include <cstdio>
#include <type_traits>

#include "dfx_adjoint.cuh"

namespace dfx {

// ---- thread-private slot map (doubles per thread), hottest first ------------------------------------------
constexpr int S_KV = 0, S_KLU = 21, S_KLV = 42;              // derivative history [7 stages][3 dofs]
constexpr int S_U0 = 63, S_V0 = 66, S_LU0 = 69, S_LV0 = 72;   // state at the step start
constexpr int S_TV = 75, S_TLU = 78, S_TW = 81;               // stage v, lambda_u, w parked across the bond phase
constexpr int S_INVM = 84, S_CD = 87;                         // 1/m and damping coefficient of the 3 DOFs
constexpr int S_BOND = 90;                                    // per bond (2 per thread) constants
constexpr int BC_R0X = 0, BC_R0Y = 1, BC_L0 = 2, BC_IL0 = 3, BC_R1X = 4, BC_R1Y = 5, BC_R2X = 6, BC_R2Y = 7,
              BC_DA1 = 8, BC_DA2 = 9, BC_N = 10;
constexpr int S_KPBC = S_BOND + 2 * BC_N;                     // 110: per-bond stiffness values [2 bonds][3] (per-bond leaves only)
constexpr int S_NCONST = S_KPBC + 6;                          // 116 slots in the TMEM / shared / global tiers
// parameter-cotangent quadratures owned by a thread (always in the L2-resident global tier)
constexpr int E_INERTIA = 0, E_DAMP = 3, E_CNV = 6 /* x[4 nodes] then y[4 nodes] */, E_REF = 14 /* [2 bonds][2] */,
              E_KPB = 18 /* [2 bonds][3] */, NE = 24;
constexpr int QA_SOL = 0, QA_ERR = 1, QA_Q0 = 2 /* parity pair 2,3 */, QA_K1 = 4 /* parity pair 4,5; the spare copy holds k7 / amid */,
              NQA = 6, QA_VAL = 6 /* staging of the integrand values for the rarely executed modes */;

// ---- tensor memory helpers (tcgen05, sm_100a) ---------------------------------------------------------
__device__ __forceinline__ void tmem_ld_issue(uint32_t taddr, uint32_t& lo, uint32_t& hi) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_pin(uint32_t& lo, uint32_t& hi) { asm volatile("" : "+r"(lo), "+r"(hi)); }
__device__ __forceinline__ void tmem_st(uint32_t taddr, double v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(__double2loint(v)), "r"(__double2hiint(v))
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// fast reciprocal of a positive normal double: hardware approximation + two Newton steps (~1 ulp, no branches)
__device__ __forceinline__ double rcp_pos(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  return r;
}

// thread-private store: slots [0,NT) in TMEM, [NT,NT+ns) in shared memory, the rest in global scratch.
// TT = CTA size (compile time, so every address is base + immediate); NS >= 0 fixes the shared tier at
// compile time, NS < 0 reads it from `ns` (lattices whose constants do not all fit in shared memory).
template <int NT, int NS, int TT>
struct TP {
  uint32_t taddr;  // TMEM address of slot 0 of this thread (lane base of the warp, column base of the warp)
  double* s;       // shared  + tid
  double* g;       // scratch + tid
  int ns;
  __device__ __forceinline__ int nsv() const { return NS >= 0 ? NS : ns; }
  __device__ __forceinline__ double ldc(int slot) const {  // a slot beyond the TMEM tier
    const int r = slot - NT;
    return r < nsv() ? s[r * TT] : g[(r - nsv()) * TT];
  }
  __device__ __forceinline__ void stc(int slot, double v) const {
    const int r = slot - NT;
    if (r < nsv()) s[r * TT] = v; else g[(r - nsv()) * TT] = v;
  }
  __device__ __forceinline__ double ld(int slot) const {
    if (NT > 0 && slot < NT) {
      uint32_t lo, hi;
      tmem_ld_issue(taddr + 2 * slot, lo, hi);
      tmem_ld_wait();
      tmem_pin(lo, hi);
      return __hiloint2double(hi, lo);
    }
    return ldc(slot);
  }
  __device__ __forceinline__ void st(int slot, double v) const {
    if (NT > 0 && slot < NT) tmem_st(taddr + 2 * slot, v);
    else stc(slot, v);
  }
  // slot known to lie in [0, LIM) (the derivative history): with NT >= LIM it is in TMEM whatever its run-time value,
  // so the tier test and the code of the other tiers disappear
  template <int LIM>
  __device__ __forceinline__ void st_below(int slot, double v) const {
    if (NT >= LIM) tmem_st(taddr + 2 * slot, v);
    else st(slot, v);
  }
  template <int LIM, int N>
  __device__ __forceinline__ void ldn_below(int slot0, int stride, double (&out)[N]) const {
    if (NT >= LIM) {
      uint32_t lo[N], hi[N];
#pragma unroll
      for (int i = 0; i < N; ++i) tmem_ld_issue(taddr + 2 * (slot0 + i * stride), lo[i], hi[i]);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < N; ++i) { tmem_pin(lo[i], hi[i]); out[i] = __hiloint2double(hi[i], lo[i]); }
    } else {
      ldn<N>(slot0, stride, out);
    }
  }
  // N strided loads with a single TMEM wait: out[i] = slot0 + i*stride
  template <int N>
  __device__ __forceinline__ void ldn(int slot0, int stride, double (&out)[N]) const {
    if (NT > 0 && slot0 + (N - 1) * stride < NT) {
      uint32_t lo[N], hi[N];
#pragma unroll
      for (int i = 0; i < N; ++i) tmem_ld_issue(taddr + 2 * (slot0 + i * stride), lo[i], hi[i]);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < N; ++i) { tmem_pin(lo[i], hi[i]); out[i] = __hiloint2double(hi[i], lo[i]); }
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) out[i] = ld(slot0 + i * stride);
    }
  }
  __device__ __forceinline__ void fence_st() const { if (NT > 0) tmem_st_wait(); }
};

// Scalar-leaf update, kept out of line: the call (register save / restore around it at 168 live registers) costs more
// than the work, so ONE call handles two groups of up to three consecutive leaves each (nB may be 0); the six warp
// reductions are interleaved.
static __device__ __noinline__ void scal_update6_nl(int mode, double cs, double ce, double cm, double cs0, double ce0, double cm0,
                                             double* wbase, int whichA, int nA, double a0, double a1, double a2,
                                             int whichB, int nB, double b0, double b1, double b2) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    b0 += __shfl_xor_sync(0xffffffffu, b0, o);
    b1 += __shfl_xor_sync(0xffffffffu, b1, o);
    b2 += __shfl_xor_sync(0xffffffffu, b2, o);
  }
  // every lane holds the six totals (butterfly): lane k updates leaf k, so the per-warp partials of the (up to) six
  // leaves are updated side by side instead of one after the other by lane 0
  const int k = threadIdx.x & 31;
  if (k >= nA + nB) return;
  double* wk1 = wbase; double* wk7 = wk1 + NSCAL * SCW; double* wsol = wk7 + NSCAL * SCW;
  double* werr = wsol + NSCAL * SCW; double* wmid = werr + NSCAL * SCW;
  const bool inA = k < nA;
  const int kk = inA ? k : k - nA;
  const double v = inA ? (kk == 0 ? a0 : (kk == 1 ? a1 : a2)) : (kk == 0 ? b0 : (kk == 1 ? b1 : b2));
  const int idx = ((inA ? whichA : whichB) + kk) * SCW + (threadIdx.x >> 5);
  switch (mode) {
    case 0: wk1[idx] = v; break;
    case 7: wk7[idx] = v; break;
    case 2: {
      const double k1 = wk1[idx];
      wsol[idx] = cs0 * k1 + cs * v; werr[idx] = ce0 * k1 + ce * v; wmid[idx] = cm0 * k1 + cm * v;
    } break;
    case 6: werr[idx] += ce * v; wmid[idx] += cm * v; wk7[idx] = v; break;
    default: wsol[idx] += cs * v; werr[idx] += ce * v; wmid[idx] += cm * v; break;
  }
}

struct Adj2Args {
  AdjArgs a;                 // same inputs / outputs as the generic kernel (placement unused)
  const int* node_bond;      // [n_nodes] bond*2+side of the bond attached to the node, or -1
  long long tp_scratch_per_design;  // doubles of global scratch per design: constants tier overflow + quadratures
  int ns_slots;              // thread-private slots placed in shared memory
  int scratch_slots;         // > 0: scratch indexed by SM id (slots available); 0: by design
  int tmem_cols_per_warp;    // columns of TMEM owned by one warp
};

// Cotangent of ys[design][i][(is_v ? n_free : 0) + f]: read from the caller's tensor `g`, or (g == NULL) formed here
// for the device objective.  Called once per output time and DOF (cold); kept out of line so that it costs the
// integration loop no registers.
static __device__ __noinline__ double cotangent_nl(const AdjArgs& a, int design, int i, int f, bool is_v) {
  if (a.g) return __ldcs(&a.g[((long long)design * a.n_t + i) * 2 * a.topo.n_free + (is_v ? a.topo.n_free : 0) + f]);
  return objective_cotangent(a, design, i, f, is_v);
}

template <int NT, int NS, int TT>
__global__ void __launch_bounds__(TT, 1) adjoint2_kernel(const __grid_constant__ Adj2Args A) {
  extern __shared__ double smem[];
  __shared__ uint32_t tmem_base_sh;
  const AdjArgs& a = A.a;
  const DevTopo& T = a.topo;
  const Tableau& tab = a.tab;
  const int design = a.order ? a.order[blockIdx.x] : (int)blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int nthr = TT, nwarp = TT / 32;
  const int NB = T.n_blocks, NN = T.n_nodes, NBONDS = T.n_bonds, npb = T.n_npb, nf = T.n_free;
  // array strides of the shared stage / slot arrays: compile-time in the specialised variant (addresses become immediates)
  const int NBS = NS >= 0 ? TT : NB, NDS = NS >= 0 ? 2 * TT : NBONDS;

  // ---- carve shared memory: red[40] | Us[5][NB] | Ws[3][NB] | SL[14][NBONDS] | SC | drv[32] | nodeb[NN] | tp[ns][T]
  double* red = smem;
  double* Us = red + 40;
  double* Ws = Us + 5 * NBS;
  double* SL = Ws + 3 * NBS;  // per bond: gdx gdy T1 T2 | hx hy H1 H2 | g1x g1y g2x g2y | a1 a2
  double* SC = SL + 14 * NDS;
  double* drv = SC + (2 * NSCAL + 5 * NSCAL * SCW);
  int* nodeb = (int*)(drv + 32);
  double* tp_s = (double*)(nodeb + ((NN + 1) & ~1));

  // ---- tensor memory: one allocation of all 512 columns per CTA -----------------------------------------
  if (NT > 0) {
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       (uint32_t)__cvta_generic_to_shared(&tmem_base_sh)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  TP<NT, NS, TT> tp;
  tp.ns = A.ns_slots;
  tp.s = tp_s + tid;
  // the scratch is indexed by SM, not by design: one CTA is resident per SM (shared-memory footprint), so the same
  // lines are reused by every design the SM processes and stay in L2 (A.scratch_slots = %nsmid bound, else per design)
  unsigned smid;
  asm("mov.u32 %0, %%smid;" : "=r"(smid));
  if (A.scratch_slots > 0 && (int)smid >= A.scratch_slots) {
    // the scratch has one slice per SM id (the host sized it for %nsmid of sm_100 parts); an id beyond it must not fall
    // back to the design index (out of bounds / shared with another SM): fail this design loudly instead
    if (tid == 0 && a.stats) {
      DfxStats st; st.steps = 0; st.accepted = 0; st.rhs_evals = 0; st.status = DFX_STATUS_NONFINITE; st.reserved = 0; st.last_dt = 0.0;
      a.stats[design] = st;
    }
    if (NT > 0) {
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_sh), "r"(512));
    }
    return;
  }
  double* gbase = a.scratch + (long long)(A.scratch_slots > 0 ? (int)smid : design) * A.tp_scratch_per_design;
  tp.g = gbase + tid;
  tp.taddr = NT > 0 ? (tmem_base_sh + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * A.tmem_cols_per_warp)) : 0u;
  // quadrature arrays [NQA][NE][T] behind the constants-tier overflow
  const int n_over = S_NCONST - NT - tp.nsv() > 0 ? S_NCONST - NT - tp.nsv() : 0;
  double* qg = gbase + n_over * nthr + tid;
  auto Q = [&](int arr, int e) -> double& { return qg[(arr * NE + e) * TT]; };

  // ---- roles ----------------------------------------------------------------------------------------------
  const bool has_blk = tid < NB;
  const int blk = has_blk ? tid : NB - 1;
  int bnd[2];
  bool has_bnd[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int b = tid + i * nthr;
    has_bnd[i] = b < NBONDS;
    bnd[i] = has_bnd[i] ? b : NBONDS - 1;
  }

  const double* g_cnv = leaf_ptr(a.p.centroid_node_vectors, design);
  const double* g_ref = leaf_ptr(a.p.reference_vector, design);
  const double* g_ks = leaf_ptr(a.p.k_stretch, design);
  const double* g_ksh = leaf_ptr(a.p.k_shear, design);
  const double* g_kr = leaf_ptr(a.p.k_rot, design);
  const double* g_inertia = leaf_ptr(a.p.inertia, design);
  const double* g_damp = leaf_ptr(a.p.damping, design);
  const double* g_contact = leaf_ptr(a.p.contact, design);
  const double* g_drive = leaf_ptr(a.p.drive, design);
  const double* ts = a.ts + (long long)design * a.ts_bstride;
  const double* ys = a.ys + (long long)design * a.n_t * 2 * nf;
  const double rtol = a.rtol, atol = a.atol;
  const bool ks_pb = a.p.k_per_bond[0], ksh_pb = a.p.k_per_bond[1], kr_pb = a.p.k_per_bond[2];
  const bool any_pb = ks_pb || ksh_pb || kr_pb;
  const bool damp_pd = a.p.damping_per_dof != 0, has_damp = T.n_damped > 0 && g_damp != nullptr;
  const bool contact = T.contact != 0;
  const int ndp = T.n_drive_params;
  const int ne_used = any_pb ? NE : E_KPB;  // quadrature entries in use
  double cmin = 0, ccut = 0, ckc = 0;
  if (contact) { cmin = g_contact[0]; ccut = g_contact[1]; ckc = g_contact[2]; }
  const double ks_u = g_ks[0], ksh_u = g_ksh[0], kr_u = g_kr[0];  // uniform stiffnesses (scalar leaves)

  // ---- per-thread topology --------------------------------------------------------------------------------
  int fidx[3], cslot[3], dslot[3];
  bool is_free[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int dof = 3 * blk + j;
    fidx[j] = T.free_of_dof[dof];
    is_free[j] = has_blk && fidx[j] >= 0;
    cslot[j] = has_blk ? T.cons_slot[dof] : -1;
    dslot[j] = T.damp_slot[dof];
  }
  // cotangent of output i for DOF j of this unit (displacement part, velocity part); cold, out of line
  auto cot_u = [&](int i, int j) { return cotangent_nl(a, design, i, fidx[j], false); };
  auto cot_v = [&](int i, int j) { return cotangent_nl(a, design, i, fidx[j], true); };
  const bool has_cons = cslot[0] >= 0 || cslot[1] >= 0 || cslot[2] >= 0;
  bool has_load = false;
  if (T.load_kind != DFX_LOAD_NONE && has_blk)
    for (int j = 0; j < 3; ++j) has_load |= T.load_mul[3 * blk + j] != 0.0;
  const bool warp_t0 = __any_sync(0xffffffffu, has_cons || has_load);
  bool warp_contact = false;
  const int2 bb0 = T.bond_blocks[bnd[0]], bb1 = T.bond_blocks[bnd[1]];

  // ---- constants into the thread-private store -------------------------------------------------------------
  for (int i = tid; i < NN; i += nthr) nodeb[i] = A.node_bond[i];
  int nbq[4];  // bond * 2 + side attached to each vertex of this thread's unit (-1: none), kept in registers
#pragma unroll
  for (int l = 0; l < 4; ++l) nbq[l] = (has_blk && l < npb) ? A.node_bond[blk * npb + l] : -1;
  for (int i = tid; i < 2 * NSCAL + 5 * NSCAL * SCW; i += nthr) SC[i] = 0.0;
  for (int i = tid; i < 14 * NDS; i += nthr) SL[i] = 0.0;
  for (int i = tid; i < 32; i += nthr) drv[i] = 0.0;
  if (tid == 0) { drv[30] = nan(""); drv[31] = nan(""); }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    tp.st(S_INVM + j, is_free[j] ? 1.0 / g_inertia[fidx[j]] : 0.0);
    tp.st(S_CD + j, (is_free[j] && dslot[j] >= 0 && has_damp) ? (damp_pd ? g_damp[dslot[j]] : g_damp[0]) : 0.0);
  }
  auto edge_angle = [&](int n, int dir) {  // angle of the edge from node n to its next (dir=+1) / previous (dir=-1) node
    const int b = n / npb, l = n - b * npb;
    const int m = b * npb + (dir > 0 ? (l + 1 == npb ? 0 : l + 1) : (l == 0 ? npb - 1 : l - 1));
    return atan2(g_cnv[2 * m + 1] - g_cnv[2 * n + 1], g_cnv[2 * m] - g_cnv[2 * n]);
  };
#pragma unroll 1
  for (int i = 0; i < 2; ++i) {
    const int s0 = S_BOND + i * BC_N, b = bnd[i];
    const int2 nd = T.bond_nodes[b];
    const double rx = g_ref[2 * b], ry = g_ref[2 * b + 1];
    tp.st(s0 + BC_R0X, rx); tp.st(s0 + BC_R0Y, ry);
    tp.st(s0 + BC_L0, sqrt(rx * rx + ry * ry)); tp.st(s0 + BC_IL0, 1.0 / sqrt(rx * rx + ry * ry));
    tp.st(s0 + BC_R1X, g_cnv[2 * nd.x]); tp.st(s0 + BC_R1Y, g_cnv[2 * nd.x + 1]);
    tp.st(s0 + BC_R2X, g_cnv[2 * nd.y]); tp.st(s0 + BC_R2Y, g_cnv[2 * nd.y + 1]);
    double da1 = 0.0, da2 = 0.0;
    if (contact) {  // psi1 = (a1_next + th1) - (a2_prev + th2), psi2 = (a2_next + th2) - (a1_prev + th1)
      da1 = edge_angle(nd.x, +1) - edge_angle(nd.y, -1);
      da2 = edge_angle(nd.y, +1) - edge_angle(nd.x, -1);
    }
    tp.st(s0 + BC_DA1, da1); tp.st(s0 + BC_DA2, da2);
    if (any_pb) {
      tp.st(S_KPBC + 3 * i, g_ks[ks_pb ? b : 0]); tp.st(S_KPBC + 3 * i + 1, g_ksh[ksh_pb ? b : 0]);
      tp.st(S_KPBC + 3 * i + 2, g_kr[kr_pb ? b : 0]);
    }
  }
  for (int q = 0; q < (NQA + 1) * NE; ++q) qg[q * TT] = 0.0;
  // y_bar = g[-1]
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    tp.st(S_LU0 + j, is_free[j] ? cot_u(a.n_t - 1, j) : 0.0);
    tp.st(S_LV0 + j, is_free[j] ? cot_v(a.n_t - 1, j) : 0.0);
  }
  tp.fence_st();
  __syncthreads();

  double* Sq0 = SC;
  double* Sqnew = SC + NSCAL;
  ScalCtx sc;
  sc.wk1 = SC + 2 * NSCAL; sc.wk7 = sc.wk1 + NSCAL * SCW; sc.wsol = sc.wk7 + NSCAL * SCW;
  sc.werr = sc.wsol + NSCAL * SCW; sc.wmid = sc.werr + NSCAL * SCW;

#define SCAL6(wA, nA, a0, a1, a2, wB, nB, b0, b1, b2) \
  scal_update6_nl(qc.mode, qc.cs, qc.ce, qc.cm, qc.cs0, qc.ce0, qc.cm0, sc.wk1, wA, nA, a0, a1, a2, wB, nB, b0, b1, b2)
  QuadCtx qc;  // pointer fields unused here
  qc.atol = atol; qc.rtol = rtol; qc.crossing = false; qc.x = 0; qc.h = 0; qc.mode = 0;
  qc.cs = qc.ce = qc.cm = qc.cs0 = qc.ce0 = qc.cm0 = 0;
  int par = 0;  // parity: which copy of Q0 / K1 is current

  // Apply the integrand values qv[0..ne_used) of this evaluation to the thread's quadratures.  Loads are
  // issued for all entries before any dependent arithmetic so the L2 latency is paid once.
  auto quad_apply = [&](const double (&qv)[NE]) -> double {
    double acc = 0.0;
    const int mode = qc.mode;
    const int a_q0 = QA_Q0 + par, a_qn = QA_Q0 + 1 - par, a_k1 = QA_K1 + par, a_k7 = QA_K1 + 1 - par;
    if (mode >= 2 && mode <= 5) {
      // handled by the prefetch / commit path of aug_BC
    } else if (mode == 6) {
      if (!qc.crossing) {
#pragma unroll
        for (int c0 = 0; c0 < NE; c0 += 8) {
          if (c0 < ne_used) {
            double s_in[8], e_in[8], q_in[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { s_in[k] = Q(QA_SOL, c0 + k); e_in[k] = Q(QA_ERR, c0 + k); q_in[k] = Q(a_q0, c0 + k); }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const double q1 = fma(qc.h, s_in[k], q_in[k]);
              const double r = qc.h * fma(qc.ce, qv[c0 + k], e_in[k]) * rcp_pos(atol + rtol * fmax(fabs(q_in[k]), fabs(q1)));
              acc = fma(r, r, acc);
              Q(a_qn, c0 + k) = q1;
              Q(a_k7, c0 + k) = qv[c0 + k];
            }
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < NE; ++e) Q(QA_VAL, e) = qv[e];
#pragma unroll 1
        for (int e = 0; e < ne_used; ++e) {
          const double val = Q(QA_VAL, e);
          const double q0 = Q(a_q0, e), q1 = fma(qc.h, Q(QA_SOL, e), q0);
          const double r = qc.h * fma(qc.ce, val, Q(QA_ERR, e)) * rcp_pos(atol + rtol * fmax(fabs(q0), fabs(q1)));
          acc = fma(r, r, acc);
          const double amid = fma(qc.cm, val, Q(a_k7, e));
          Q(a_qn, e) = interp_eval(q0, q1, q0 + qc.h * amid, qc.h * Q(a_k1, e), qc.h * val, qc.x);
        }
      }
    } else if (mode == 0) {
#pragma unroll
      for (int e = 0; e < NE; ++e) if (e < ne_used) Q(a_k1, e) = qv[e];
    } else if (mode == 7) {
#pragma unroll
      for (int e = 0; e < NE; ++e) Q(QA_VAL, e) = qv[e];
#pragma unroll 1
      for (int e = 0; e < ne_used; ++e) {
        const double d = (Q(QA_VAL, e) - Q(a_k1, e)) * rcp_pos(atol + fabs(Q(a_q0, e)) * rtol);
        acc = fma(d, d, acc);
      }
    }
    return acc;
  };

  // Accumulating stages (RK stages 2..5 of a step): the running solution / error sums of a group of entries are
  // fetched early (before the bond arithmetic or the CTA barrier) and committed once the integrands are known, so
  // the L2 latency of the thread-private quadratures overlaps with computation.
  struct AccCoef { int a_s, a_e, a_m, a_k7; double fs, fe, fm; bool on; };
  auto acc_coef = [&]() {
    AccCoef c;
    const int mode = qc.mode;
    c.on = mode >= 2 && mode <= 5;
    const bool first = mode == 2;
    const int a_k1 = QA_K1 + par;
    c.a_k7 = QA_K1 + 1 - par;
    c.a_s = first ? a_k1 : QA_SOL; c.a_e = first ? a_k1 : QA_ERR; c.a_m = first ? a_k1 : c.a_k7;
    c.fs = first ? qc.cs0 : 1.0; c.fe = first ? qc.ce0 : 1.0; c.fm = first ? qc.cm0 : 1.0;
    return c;
  };

  // ---- one augmented RHS evaluation: phases B and C.  The stage v, lambda_u and w of this thread's unit
  // are parked in the thread-private store by publish(); derivative stage `kidx` is written there too.
  // `time_next`: real time of the next evaluation when it is already known (drive channels are prepared for it).
#ifdef DFX_PHASE_TIMERS
  // A | wait A->B | B bond math + slots | wait B->C | C rest | step logic | B bond quadrature | B scalars | C gather | C k + store | C unit quadrature
  long long pt_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, pt_mark = 0;
#define PT_MARK(k) do { const long long now_ = clock64(); pt_acc[k] += now_ - pt_mark; pt_mark = now_; } while (0)
#else
#define PT_MARK(k) do { } while (0)
#endif
  auto aug_BC = [&](double time, int kidx, double time_next) -> double {
    const bool want_q = qc.mode != 1;
    constexpr int NEB = NE - E_REF;  // bond-owned entries: reference vector (+ per-bond stiffnesses)
    double qb[NEB];                  // integrands of the bond-owned entries (live from phase B to the end of phase C)
#pragma unroll
    for (int e = 0; e < NEB; ++e) qb[e] = 0.0;
    PT_MARK(0);
    __syncthreads();
    PT_MARK(1);
    // ============ phase B: bonds ============
    double p_ks = 0, p_ksh = 0, p_kr = 0, p_c0 = 0, p_c1 = 0, p_c2 = 0;
    const AccCoef ac = acc_coef();
    if (tid == nthr - 1 && T.drive_kind != DFX_DRIVE_ZERO) {  // drive channels: now (with derivatives) and next
      DriveEval de;
      drive_eval(T.drive_kind, time, g_drive, true, de, T.table);
      drv[2] = de.sdot[0]; drv[3] = de.sdot[1];
#pragma unroll
      for (int q = 0; q < DFX_MAX_DRIVE_PARAMS; ++q) { drv[4 + q] = de.dsdp[0][q]; drv[9 + q] = de.dsdp[1][q]; }
      drive_eval(T.drive_kind, time_next, g_drive, false, de, T.table);
      drv[28] = de.s[0]; drv[29] = de.s[1]; drv[30] = time_next;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int b = bnd[i];
      const int b1 = i == 0 ? bb0.x : bb1.x, b2 = i == 0 ? bb0.y : bb1.y;
      double c[BC_N];
      tp.template ldn<BC_N>(S_BOND + i * BC_N, 1, c);
      double ks = ks_u, ksh = ksh_u, kr = kr_u;
      if (any_pb) { ks = tp.ldc(S_KPBC + 3 * i); ksh = tp.ldc(S_KPBC + 3 * i + 1); kr = tp.ldc(S_KPBC + 3 * i + 2); }
      BlockState<Dual> s1, s2;
      make_block(Us[b1], Us[NBS + b1], Us[2 * NBS + b1], Us[3 * NBS + b1], Us[4 * NBS + b1], Ws[b1], Ws[NBS + b1], Ws[2 * NBS + b1], s1);
      make_block(Us[b2], Us[NBS + b2], Us[2 * NBS + b2], Us[3 * NBS + b2], Us[4 * NBS + b2], Ws[b2], Ws[NBS + b2], Ws[2 * NBS + b2], s2);
      BondConst bc = {c[BC_R0X], c[BC_R0Y], c[BC_L0], c[BC_IL0]};
      BondOut<Dual> o;
      bond_gradient<Dual, true>(T.bond_energy, s1, s2, c[BC_R1X], c[BC_R1Y], c[BC_R2X], c[BC_R2Y], bc, ks, ksh, kr, o);
      double a1 = 0.0, a2 = 0.0;
      if (contact) {
        Dual psi1 = wrapT(s1.th - s2.th + c[BC_DA1]);
        Dual psi2 = wrapT(s2.th - s1.th + c[BC_DA2]);
        const bool act1 = !(psi1.v < cmin) && psi1.v < ccut, act2 = !(psi2.v < cmin) && psi2.v < ccut;
        if (act1 || act2) {
          Dual e1, e2, m1, m2, c1, c2, k1, k2;
          contact_term<Dual>(psi1, cmin, ccut, ckc, e1, m1, c1, k1);
          contact_term<Dual>(psi2, cmin, ccut, ckc, e2, m2, c2, k2);
          o.f1[2] = o.f1[2] + e1 - e2;
          o.f2[2] = o.f2[2] + e2 - e1;
          a1 = e1.d; a2 = e2.d;
          if (has_bnd[i]) { p_c0 -= m1.d + m2.d; p_c1 -= c1.d + c2.d; p_c2 -= k1.d + k2.d; }
        }
      }
      if (has_bnd[i]) {
        // forces on the two ends are equal and opposite: store (gdx, gdy) once, the two torques separately
        SL[b] = o.f2[0].v; SL[NDS + b] = o.f2[1].v; SL[2 * NDS + b] = -o.f1[2].v; SL[3 * NDS + b] = -o.f2[2].v;
        SL[4 * NDS + b] = o.f2[0].d; SL[5 * NDS + b] = o.f2[1].d; SL[6 * NDS + b] = o.f1[2].d; SL[7 * NDS + b] = o.f2[2].d;
        SL[8 * NDS + b] = -o.gr1[0].d; SL[9 * NDS + b] = -o.gr1[1].d;
        SL[10 * NDS + b] = -o.gr2[0].d; SL[11 * NDS + b] = -o.gr2[1].d;
        if (contact) { SL[12 * NDS + b] = a1; SL[13 * NDS + b] = a2; }
        // d(w.F)/dp = -(dual part of dE/dp)
        // (static indices keep qv in registers although the bond loop is not unrolled)
        if (i == 0) { qb[0] = -o.gr0[0].d; qb[1] = -o.gr0[1].d; }
        else { qb[2] = -o.gr0[0].d; qb[3] = -o.gr0[1].d; }
        constexpr int KP = E_KPB - E_REF;
        if (ks_pb) { if (i == 0) qb[KP] = -o.gks.d; else qb[KP + 3] = -o.gks.d; } else p_ks -= o.gks.d;
        if (ksh_pb) { if (i == 0) qb[KP + 1] = -o.gksh.d; else qb[KP + 4] = -o.gksh.d; } else p_ksh -= o.gksh.d;
        if (kr_pb) { if (i == 0) qb[KP + 2] = -o.gkr.d; else qb[KP + 5] = -o.gkr.d; } else p_kr -= o.gkr.d;
      }
    }
    PT_MARK(2);
    if (ac.on && want_q) {
      // bond-owned entries: all loads first, then the updates
      double bs_in[NEB], be_in[NEB];
#pragma unroll
      for (int k = 0; k < NEB; ++k) if (E_REF + k < ne_used) { bs_in[k] = Q(ac.a_s, E_REF + k); be_in[k] = Q(ac.a_e, E_REF + k); }
#pragma unroll
      for (int k = 0; k < NEB; ++k) if (E_REF + k < ne_used) {
        Q(QA_SOL, E_REF + k) = fma(qc.cs, qb[k], ac.fs * bs_in[k]);
        Q(QA_ERR, E_REF + k) = fma(qc.ce, qb[k], ac.fe * be_in[k]);
      }
    }
    PT_MARK(6);
    if (want_q) {
      // a warp takes part in the contact scalars from the first time one of its bonds touches (sticky)
      if (contact) warp_contact = warp_contact || __any_sync(0xffffffffu, p_c0 != 0.0 || p_c1 != 0.0 || p_c2 != 0.0);
      const int n_k = (!ks_pb || !ksh_pb || !kr_pb) ? 3 : 0, n_c = (contact && warp_contact) ? 3 : 0;
      if (n_k + n_c > 0)
        SCAL6(SC_KS, n_k, ks_pb ? 0.0 : p_ks, ksh_pb ? 0.0 : p_ksh, kr_pb ? 0.0 : p_kr, SC_CONTACT, n_c, p_c0, p_c1, p_c2);
    }
    PT_MARK(7);
    __syncthreads();
    PT_MARK(3);
    // ============ phase C: this thread's unit ============
    double qv[NE];  // unit-owned integrands in [0, E_REF); the bond-owned ones are appended only for the rare modes
#pragma unroll
    for (int e = 0; e < E_REF; ++e) qv[e] = 0.0;
    double F[3] = {0, 0, 0}, HW[3] = {0, 0, 0}, An[4] = {0, 0, 0, 0}, Ap[4] = {0, 0, 0, 0};
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      if (l < npb) {
        const int nb_ = nbq[l];
        if (nb_ >= 0) {
          const int b = nb_ >> 1;
          const bool second = nb_ & 1;
          const double sg = second ? -1.0 : 1.0;
          F[0] += sg * SL[b]; F[1] += sg * SL[NDS + b]; F[2] += SL[(second ? 3 : 2) * NDS + b];
          HW[0] -= sg * SL[4 * NDS + b]; HW[1] -= sg * SL[5 * NDS + b]; HW[2] += SL[(second ? 7 : 6) * NDS + b];
          qv[E_CNV + l] = SL[(second ? 10 : 8) * NDS + b]; qv[E_CNV + 4 + l] = SL[(second ? 11 : 9) * NDS + b];
          if (contact) {
            // dS/dalpha = -(dual part of dE/dalpha): a1next:+e1, a1prev:-e2, a2next:+e2, a2prev:-e1
            const double e1d = SL[12 * NDS + b], e2d = SL[13 * NDS + b];
            An[l] = second ? -e2d : -e1d;
            Ap[l] = second ? e1d : e2d;
          }
        }
      }
    }
    PT_MARK(8);
    double ls = 0.0, lsd = 0.0;
    if (T.load_kind != DFX_LOAD_NONE) load_eval(T.load_kind, time, T.load_consts, ls, lsd);
    double p_t0 = 0, p_damp = 0, p_dr0 = 0, p_dr1 = 0, p_dr2 = 0, p_dr3 = 0, p_dr4 = 0;
    double im[3], cdv[3], vst[3], lust[3], wst[3];
    tp.template ldn<3>(S_INVM, 1, im);
    tp.template ldn<3>(S_CD, 1, cdv);
    tp.template ldn<3>(S_TV, 1, vst);
    tp.template ldn<3>(S_TLU, 1, lust);
    tp.template ldn<3>(S_TW, 1, wst);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double kvv = 0.0, kluv = 0.0, klvv = 0.0;
      if (is_free[j]) {
        double lm = 0.0, Fj = F[j];
        if (T.load_kind != DFX_LOAD_NONE) { lm = T.load_mul[3 * blk + j]; Fj += lm * ls; }
        const double acc = (Fj - cdv[j] * vst[j]) * im[j];
        kvv = -acc; kluv = -HW[j]; klvv = lust[j] - cdv[j] * wst[j];
        qv[E_INERTIA + j] = -wst[j] * acc;
        if (has_damp && dslot[j] >= 0) { if (damp_pd) qv[E_DAMP + j] = -wst[j] * vst[j]; else p_damp -= wst[j] * vst[j]; }
        p_t0 += wst[j] * lm * lsd;
      }
      tp.template st_below<63>(S_KV + kidx * 3 + j, kvv);
      tp.template st_below<63>(S_KLU + kidx * 3 + j, kluv);
      tp.template st_below<63>(S_KLV + kidx * 3 + j, klvv);
    }
    if (has_cons && want_q && T.drive_kind != DFX_DRIVE_ZERO) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (cslot[j] >= 0) {
          const double v0_ = T.drive_vec0[cslot[j]], v1_ = T.drive_vec1[cslot[j]];
          p_t0 -= HW[j] * (v0_ * drv[2] + v1_ * drv[3]);
          p_dr0 -= HW[j] * (v0_ * drv[4] + v1_ * drv[9]);
          p_dr1 -= HW[j] * (v0_ * drv[5] + v1_ * drv[10]);
          p_dr2 -= HW[j] * (v0_ * drv[6] + v1_ * drv[11]);
          p_dr3 -= HW[j] * (v0_ * drv[7] + v1_ * drv[12]);
          p_dr4 -= HW[j] * (v0_ * drv[8] + v1_ * drv[13]);
        }
      }
    }
    PT_MARK(9);
    double probe = 0.0;
    if (want_q) {
      if (contact) {
        bool any_contact = false;
#pragma unroll
        for (int l = 0; l < 4; ++l) any_contact |= (An[l] != 0.0) | (Ap[l] != 0.0);
        if (any_contact) {
          // contact chain of the centroid_node_vectors cotangent: edge l -> l+1 is node l's "next" edge and,
          // reversed, node (l+1)'s "previous" edge; both angles have the same derivative w.r.t. the end points
#pragma unroll
          for (int l = 0; l < 4; ++l) {
            if (l < npb) {
              const bool last = l + 1 == npb;
              const int ln = last ? 0 : l + 1;
              const int n = blk * npb + l, m = blk * npb + ln;
              const double ex = g_cnv[2 * m] - g_cnv[2 * n], ey = g_cnv[2 * m + 1] - g_cnv[2 * n + 1];
              const double inv = 1.0 / (ex * ex + ey * ey);
              const double ap_n = last ? Ap[0] : Ap[l + 1 < 4 ? l + 1 : 0];
              const double wx = -(An[l] + ap_n) * ey * inv, wy = (An[l] + ap_n) * ex * inv;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k == ln) { qv[E_CNV + k] += wx; qv[E_CNV + 4 + k] += wy; }
                if (k == l) { qv[E_CNV + k] -= wx; qv[E_CNV + 4 + k] -= wy; }
              }
            }
          }
        }
      }
      PT_MARK(10);
      if (ac.on) {
        // unit-owned entries in two groups: all loads of a group are in flight before the first dependent use
#pragma unroll
        for (int c0 = 0; c0 < E_REF; c0 += 7) {
          double s_in[7], e_in[7];
#pragma unroll
          for (int k = 0; k < 7; ++k) { s_in[k] = Q(ac.a_s, c0 + k); e_in[k] = Q(ac.a_e, c0 + k); }
#pragma unroll
          for (int k = 0; k < 7; ++k) {
            Q(QA_SOL, c0 + k) = fma(qc.cs, qv[c0 + k], ac.fs * s_in[k]);
            Q(QA_ERR, c0 + k) = fma(qc.ce, qv[c0 + k], ac.fe * e_in[k]);
          }
        }
        if (qc.crossing) {  // midpoint sums (one step in ~20): loads of a group in flight together, then the stores
#pragma unroll
          for (int c0 = 0; c0 < E_REF; c0 += 7) {
            double m_in[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) m_in[k] = Q(ac.a_m, c0 + k);
#pragma unroll
            for (int k = 0; k < 7; ++k) Q(ac.a_k7, c0 + k) = fma(qc.cm, qv[c0 + k], ac.fm * m_in[k]);
          }
          double mb_in[NEB];
#pragma unroll
          for (int k = 0; k < NEB; ++k) mb_in[k] = E_REF + k < ne_used ? Q(ac.a_m, E_REF + k) : 0.0;
#pragma unroll
          for (int k = 0; k < NEB; ++k) if (E_REF + k < ne_used) Q(ac.a_k7, E_REF + k) = fma(qc.cm, qb[k], ac.fm * mb_in[k]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < NEB; ++k) qv[E_REF + k] = qb[k];
        probe = quad_apply(qv);
      }
      PT_MARK(11);
      // t0_bar and the drive parameters only receive contributions from constrained or loaded DOFs
      if (warp_t0) {
        SCAL6(SC_T0, 1, p_t0, 0.0, 0.0, SC_DRIVE, ndp < 3 ? ndp : 3, p_dr0, p_dr1, p_dr2);
        if (ndp > 3) SCAL6(SC_DRIVE + 3, ndp - 3, p_dr3, p_dr4, 0.0, SC_DAMP, 0, 0.0, 0.0, 0.0);
      }
      if (has_damp && !damp_pd) SCAL6(SC_DAMP, 1, p_damp, 0.0, 0.0, SC_DAMP, 0, 0.0, 0.0, 0.0);
    }
    tp.fence_st();
    return probe;
  };

  // publish the stage state of this thread's unit (time = real time of the stage); parks v, lambda_u, w
  auto publish = [&](double (&u)[3], double (&v)[3], double (&lu)[3], double (&lv)[3], double time) {
    double im[3];
    tp.template ldn<3>(S_INVM, 1, im);
    if (has_cons && T.drive_kind != DFX_DRIVE_ZERO) {
      double s0_, s1_;
      if (drv[30] == time) { s0_ = drv[28]; s1_ = drv[29]; }
      else {
        DriveEval de;
        drive_eval(T.drive_kind, time, g_drive, false, de, T.table);
        s0_ = de.s[0]; s1_ = de.s[1];
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) if (cslot[j] >= 0) u[j] = T.drive_vec0[cslot[j]] * s0_ + T.drive_vec1[cslot[j]] * s1_;
    }
    double w[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (!is_free[j]) { v[j] = 0.0; lu[j] = 0.0; lv[j] = 0.0; if (cslot[j] < 0) u[j] = 0.0; }
      w[j] = lv[j] * im[j];
      tp.st(S_TV + j, v[j]); tp.st(S_TLU + j, lu[j]); tp.st(S_TW + j, w[j]);
    }
    if (has_blk) {
      double sn, cs;
      sincos_fast(u[2], &sn, &cs);
      Us[blk] = u[0]; Us[NBS + blk] = u[1]; Us[2 * NBS + blk] = u[2]; Us[3 * NBS + blk] = sn; Us[4 * NBS + blk] = cs;
      Ws[blk] = w[0]; Ws[NBS + blk] = w[1]; Ws[2 * NBS + blk] = w[2];
    }
    tp.fence_st();
  };

  auto wtotal = [&](const double* wa, int which) {
    double s = 0.0;
#pragma unroll 1
    for (int w = 0; w < nwarp; ++w) s += wa[which * SCW + w];
    return s;
  };

  // stage state of RK stage st (run time) from the stored history
  auto stage_state = [&](int st, double h, double (&us)[3], double (&vs)[3], double (&lus)[3], double (&lvs)[3]) {
    const double ha = h * tab.alpha[st], h2 = h * h;
    double au[3] = {0, 0, 0}, av[3] = {0, 0, 0}, alu[3] = {0, 0, 0}, alv[3] = {0, 0, 0};
#pragma unroll 1
    for (int l = 0; l <= st; ++l) {
      const double b = tab.beta[st][l], b2 = tab.a2[st][l];
      double kv[3], klu[3], klv[3];
      tp.template ldn_below<63, 3>(S_KV + 3 * l, 1, kv);
      tp.template ldn_below<63, 3>(S_KLU + 3 * l, 1, klu);
      tp.template ldn_below<63, 3>(S_KLV + 3 * l, 1, klv);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        au[j] = fma(b2, kv[j], au[j]);
        av[j] = fma(b, kv[j], av[j]);
        alu[j] = fma(b, klu[j], alu[j]);
        alv[j] = fma(b, klv[j], alv[j]);
      }
    }
    double y0[12];
    tp.template ldn<12>(S_U0, 1, y0);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      us[j] = y0[j] - ha * y0[3 + j] - h2 * au[j];
      vs[j] = y0[3 + j] + h * av[j];
      lus[j] = y0[6 + j] + h * alu[j];
      lvs[j] = y0[9 + j] + h * alv[j];
    }
  };

  long long n_steps = 0, n_acc = 0, n_rhs = 0, istep = 0;
  int status = 0;
  double h = 0.0, h0 = 0.0, d0 = 0.0, d1 = 0.0, s0 = 0.0, s_target = 0.0, s_cur = 0.0, se = 0.0;
  const double inv_n = 1.0 / (double)a.aug_size;

  // One loop iteration = one augmented RHS evaluation, so the heavy bond code exists once in the binary.
  // ev: 0..5 = Runge-Kutta stage, EV_INIT = f0 at the start of an output interval, EV_PROBE = the second
  // evaluation of initial_step_size.
  constexpr int EV_INIT = 6, EV_PROBE = 7;
  int ev = EV_INIT;
  int i = a.n_t - 1;
  bool running = i >= 1;
  auto begin_step = [&]() {  // returns false when the integration must stop
    if (!(h > 0.0)) { status |= DFX_STATUS_DT_UNDERFLOW; return false; }
    if (istep >= a.max_steps) { status |= DFX_STATUS_MAX_STEPS; return false; }
    const double s_new = s_cur + h;
    qc.h = h;
    qc.crossing = !(s_new < s_target);
    qc.x = (s_target - s_cur) / (s_new - s_cur);
    se = 0.0;
    ev = 0;
    return true;
  };

#ifdef DFX_PHASE_TIMERS
  pt_mark = clock64();
#endif
  while (running) {
    PT_MARK(5);
    double us[3], vs[3], lus[3], lvs[3], time;
    int kidx;
    // ---------------- stage state ----------------
    switch (ev) {
      case EV_INIT: {
        s0 = -ts[i]; s_target = -ts[i - 1];
        const double* yi = ys + (long long)i * 2 * nf;
        tp.template ldn<3>(S_LU0, 1, lus);
        tp.template ldn<3>(S_LV0, 1, lvs);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          us[j] = is_free[j] ? __ldcs(&yi[fidx[j]]) : 0.0;
          vs[j] = is_free[j] ? __ldcs(&yi[nf + fidx[j]]) : 0.0;
          tp.st(S_U0 + j, us[j]); tp.st(S_V0 + j, vs[j]);
        }
        time = -s0; kidx = 0; qc.mode = 0;
      } break;
      case EV_PROBE: {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          double y0[4], k0[3];
          tp.template ldn<4>(S_U0 + j, 3, y0);
          tp.template ldn<3>(S_KV + j, 21, k0);  // kv[0][j], klu[0][j], klv[0][j]
          us[j] = y0[0] - h0 * y0[1]; vs[j] = y0[1] + h0 * k0[0]; lus[j] = y0[2] + h0 * k0[1]; lvs[j] = y0[3] + h0 * k0[2];
        }
        time = -(s0 + h0); kidx = 1; qc.mode = 7;
      } break;
      default: stage_state(ev, h, us, vs, lus, lvs); break;
    }
    if (ev < 6) {
      kidx = ev + 1;
      time = -(s_cur + h * tab.alpha[ev]);
      qc.mode = kidx;
      qc.cs = tab.c_sol[kidx]; qc.ce = tab.c_err[kidx]; qc.cm = tab.c_mid[kidx];
      qc.cs0 = tab.c_sol[0]; qc.ce0 = tab.c_err[0]; qc.cm0 = tab.c_mid[0];
    }
    publish(us, vs, lus, lvs, time);
    // real time of the next evaluation when it is already determined (stages 0..4 of a step)
    const double time_next = ev < 5 ? -(s_cur + h * tab.alpha[ev + 1]) : nan("");
    const double probe = aug_BC(time, kidx, time_next);
    PT_MARK(4);
    n_rhs++;
    // ---------------- what follows the evaluation ----------------
    if (ev == EV_INIT) {
      double sd0 = 0, sd1 = 0, pt = 0.0;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double y0[4], k0[3];
        tp.template ldn<4>(S_U0 + j, 3, y0);
        tp.template ldn<3>(S_KV + j, 21, k0);
        if (is_free[j]) {
          // t_bar = func(ys[i], ts[i]) . g[i] with func = (v0, -kv[0])
          pt += y0[1] * cot_u(i, j) - k0[0] * cot_v(i, j);
          const double su = atol + fabs(y0[0]) * rtol, sv = atol + fabs(y0[1]) * rtol;
          const double slu = atol + fabs(y0[2]) * rtol, slv = atol + fabs(y0[3]) * rtol;
          const double a0 = y0[0] / su, a1 = y0[1] / sv, a2 = y0[2] / slu, a3 = y0[3] / slv;
          const double b0 = -y0[1] / su, b1 = k0[0] / sv, b2 = k0[1] / slu, b3 = k0[2] / slv;
          sd0 += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
          sd1 += b0 * b0 + b1 * b1 + b2 * b2 + b3 * b3;
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
      for (int e = 0; e < ne_used; ++e) {
        const double q0 = Q(QA_Q0 + par, e), k1 = Q(QA_K1 + par, e);
        const double s = atol + fabs(q0) * rtol;
        sd0 += (q0 / s) * (q0 / s); sd1 += (k1 / s) * (k1 / s);
      }
      if (tid < NSCAL) {
        const double s = atol + fabs(Sq0[tid]) * rtol;
        const double a0 = Sq0[tid] / s, b0 = wtotal(sc.wk1, tid) / s;
        sd0 += a0 * a0; sd1 += b0 * b0;
      }
      d0 = sqrt(block_sum(sd0, red));
      d1 = sqrt(block_sum(sd1, red));
      h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
      ev = EV_PROBE;
    } else if (ev == EV_PROBE) {
      double sd2 = probe;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double y0[4], k0[3], k1[3];
        tp.template ldn<4>(S_U0 + j, 3, y0);
        tp.template ldn<3>(S_KV + j, 21, k0);
        tp.template ldn<3>(S_KV + 3 + j, 21, k1);
        if (is_free[j]) {
          const double su = atol + fabs(y0[0]) * rtol, sv = atol + fabs(y0[1]) * rtol;
          const double slu = atol + fabs(y0[2]) * rtol, slv = atol + fabs(y0[3]) * rtol;
          const double b0 = (-vs[j] + y0[1]) / su, b1 = (k1[0] - k0[0]) / sv;
          const double b2 = (k1[1] - k0[1]) / slu, b3 = (k1[2] - k0[2]) / slv;
          sd2 += b0 * b0 + b1 * b1 + b2 * b2 + b3 * b3;
        }
      }
      __syncthreads();
      if (tid < NSCAL) {
        const double s = atol + fabs(Sq0[tid]) * rtol;
        const double b0 = (wtotal(sc.wk7, tid) - wtotal(sc.wk1, tid)) / s;
        sd2 += b0 * b0;
      }
      const double d2 = sqrt(block_sum(sd2, red)) / h0;
      double h1;
      if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
      else h1 = pow(0.01 / (a.init_step_variant == 0 ? d1 + d2 : fmax(d1, d2)), 0.2);
      h = fmin(100.0 * h0, h1);
      s_cur = s0;
      istep = 0;
      if (!(s_cur < s_target)) {  // empty interval (repeated output time): nothing to integrate
        if (--i < 1) running = false; else ev = EV_INIT;
      } else if (!begin_step()) running = false;
    } else if (ev < 5) {
      se += probe;
      ++ev;
    } else {
      se += probe;
      // solution y1 = y0 + h*dot(c_sol, k) and error estimate of the dynamic entries
      double y1[3][4];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double kv[7], klu[7], klv[7], y0[4];
        tp.template ldn<7>(S_KV + j, 3, kv);
        tp.template ldn<7>(S_KLU + j, 3, klu);
        tp.template ldn<7>(S_KLV + j, 3, klv);
        tp.template ldn<4>(S_U0 + j, 3, y0);
        double eu = 0.0, evv = 0.0, elu = 0.0, elv = 0.0, su = 0.0, sv = 0.0, slu = 0.0, slv = 0.0;
#pragma unroll
        for (int l = 0; l < 7; ++l) {
          eu = fma(tab.e2[l], kv[l], eu);
          evv = fma(tab.c_err[l], kv[l], evv);
          elu = fma(tab.c_err[l], klu[l], elu);
          elv = fma(tab.c_err[l], klv[l], elv);
          su = fma(tab.s2[l], kv[l], su);
          sv = fma(tab.c_sol[l], kv[l], sv);
          slu = fma(tab.c_sol[l], klu[l], slu);
          slv = fma(tab.c_sol[l], klv[l], slv);
        }
        y1[j][0] = y0[0] - h * (tab.sum_sol * y0[1] + h * su);
        y1[j][1] = y0[1] + h * sv;
        y1[j][2] = y0[2] + h * slu;
        y1[j][3] = y0[3] + h * slv;
        eu = -h * (tab.sum_err * y0[1] + h * eu);
        evv *= h; elu *= h; elv *= h;
        if (is_free[j]) {
          const double r0 = eu / (atol + rtol * fmax(fabs(y0[0]), fabs(y1[j][0])));
          const double r1 = evv / (atol + rtol * fmax(fabs(y0[1]), fabs(y1[j][1])));
          const double r2 = elu / (atol + rtol * fmax(fabs(y0[2]), fabs(y1[j][2])));
          const double r3 = elv / (atol + rtol * fmax(fabs(y0[3]), fabs(y1[j][3])));
          se += r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3;
        }
      }
      __syncthreads();  // per-warp scalar partials of the last stage complete
      if (tid < NSCAL) {
        const double q0 = Sq0[tid], k1 = wtotal(sc.wk1, tid), k7 = wtotal(sc.wk7, tid);
        const double q1 = q0 + h * wtotal(sc.wsol, tid);
        const double r = h * wtotal(sc.werr, tid) / (atol + rtol * fmax(fabs(q0), fabs(q1)));
        se += r * r;
        Sqnew[tid] = qc.crossing ? interp_eval(q0, q1, q0 + h * wtotal(sc.wmid, tid), h * k1, h * k7, qc.x) : q1;
      }
      const double ratio = sqrt(block_sum(se, red) * inv_n);
      ++n_steps; ++istep;
      bool interval_done = false;
      if (!isfinite(ratio)) { status |= DFX_STATUS_NONFINITE; running = false; }
      else {
        if (ratio <= 1.0) {
          ++n_acc;
          if (qc.crossing) {
            // interval finished: cotangents interpolated at s_target, plus g[i-1]
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              double klu[7], klv[7], y0[4];
              tp.template ldn<7>(S_KLU + j, 3, klu);
              tp.template ldn<7>(S_KLV + j, 3, klv);
              tp.template ldn<4>(S_U0 + j, 3, y0);
              double mlu = 0.0, mlv = 0.0;
#pragma unroll
              for (int l = 0; l < 7; ++l) { mlu = fma(tab.c_mid[l], klu[l], mlu); mlv = fma(tab.c_mid[l], klv[l], mlv); }
              const double nlu = interp_eval(y0[2], y1[j][2], y0[2] + h * mlu, h * klu[0], h * klu[6], qc.x);
              const double nlv = interp_eval(y0[3], y1[j][3], y0[3] + h * mlv, h * klv[0], h * klv[6], qc.x);
              tp.st(S_LU0 + j, is_free[j] ? nlu + cot_u(i - 1, j) : 0.0);
              tp.st(S_LV0 + j, is_free[j] ? nlv + cot_v(i - 1, j) : 0.0);
            }
            interval_done = true;
          } else {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              double k6[3];
              tp.template ldn<3>(S_KV + 18 + j, 21, k6);
              tp.st(S_U0 + j, y1[j][0]); tp.st(S_V0 + j, y1[j][1]); tp.st(S_LU0 + j, y1[j][2]); tp.st(S_LV0 + j, y1[j][3]);
              tp.st(S_KV + j, k6[0]); tp.st(S_KLU + j, k6[1]); tp.st(S_KLV + j, k6[2]);
            }
            if (tid < NSCAL) for (int w = 0; w < nwarp; ++w) sc.wk1[tid * SCW + w] = sc.wk7[tid * SCW + w];
          }
          tp.fence_st();
          if (tid < NSCAL) Sq0[tid] = Sqnew[tid];
          par ^= 1;  // Q0 <- QNEW, K1 <- K7 for every thread-private quadrature
          s_cur = s_cur + h;
        }
        const double dfactor = ratio < 1.0 ? 1.0 : 0.2;
        const double factor = fmin(10.0, fmax(inv_fifth_root(ratio) * 0.9, dfactor));
        h = (ratio == 0.0) ? h * 10.0 : h * factor;
        __syncthreads();
        if (interval_done) {
          if (--i < 1) running = false; else ev = EV_INIT;
        } else if (!begin_step()) running = false;
      }
    }
  }

#ifdef DFX_PHASE_TIMERS
  if (design == 0 && (tid == 0 || tid == 352))
    printf("phase timers design 0 thread %d [cycles]: A %lld | wait A->B %lld | B bond math+slots %lld | B bond quadrature %lld | B scalars %lld | wait B->C %lld | "
           "C gather %lld | C k+store %lld | C contact chain %lld | C unit quadrature %lld | C scalars+rest %lld | step logic %lld | evaluations %lld\n",
           tid, pt_acc[0], pt_acc[1], pt_acc[2], pt_acc[6], pt_acc[7], pt_acc[3], pt_acc[8], pt_acc[9], pt_acc[10], pt_acc[11], pt_acc[4], pt_acc[5], n_rhs);
#endif
  // ---- outputs ----------------------------------------------------------------------------------------------
  __syncthreads();
  const double nanv = nan("");
  const bool bad = status != 0;
  {
    double qv[NE], lu0[3], lv0[3];
    for (int e = 0; e < NE; ++e) qv[e] = bad ? nanv : Q(QA_Q0 + par, e);
    tp.template ldn<3>(S_LU0, 1, lu0);
    tp.template ldn<3>(S_LV0, 1, lv0);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (has_blk && a.grads.damping && has_damp && damp_pd && dslot[j] >= 0)
        a.grads.damping[(long long)design * T.n_damped * 3 + dslot[j]] = is_free[j] ? qv[E_DAMP + j] : (bad ? nanv : 0.0);
      if (!is_free[j]) continue;
      if (a.y0_bar) {
        a.y0_bar[(long long)design * 2 * nf + fidx[j]] = bad ? nanv : lu0[j];
        a.y0_bar[(long long)design * 2 * nf + nf + fidx[j]] = bad ? nanv : lv0[j];
      }
      if (a.grads.inertia) a.grads.inertia[(long long)design * nf + fidx[j]] = qv[E_INERTIA + j];
    }
    if (has_blk && a.grads.centroid_node_vectors)
      for (int l = 0; l < npb; ++l) {
        const long long n = (long long)design * NN + blk * npb + l;
        a.grads.centroid_node_vectors[n * 2] = qv[E_CNV + l];
        a.grads.centroid_node_vectors[n * 2 + 1] = qv[E_CNV + 4 + l];
      }
    for (int i = 0; i < 2; ++i)
      if (has_bnd[i] && a.grads.reference_vector) {
        a.grads.reference_vector[((long long)design * NBONDS + bnd[i]) * 2] = qv[E_REF + 2 * i];
        a.grads.reference_vector[((long long)design * NBONDS + bnd[i]) * 2 + 1] = qv[E_REF + 2 * i + 1];
      }
  }
  {
    double* outs[3] = {a.grads.k_stretch, a.grads.k_shear, a.grads.k_rot};
    const bool pb[3] = {ks_pb, ksh_pb, kr_pb};
    for (int k = 0; k < 3; ++k) {
      if (!outs[k]) continue;
      if (pb[k]) {
        for (int i = 0; i < 2; ++i) {
          const double v = bad ? nanv : Q(QA_Q0 + par, E_KPB + 3 * i + k);
          if (has_bnd[i]) outs[k][(long long)design * NBONDS + bnd[i]] = v;
        }
      } else if (tid == 0) outs[k][design] = bad ? nanv : Sq0[SC_KS + k];
    }
  }
  if (tid == 0) {
    if (a.grads.damping && has_damp && !damp_pd) a.grads.damping[design] = bad ? nanv : Sq0[SC_DAMP];
    if (a.grads.contact && contact) for (int k = 0; k < 3; ++k) a.grads.contact[(long long)design * 3 + k] = bad ? nanv : Sq0[SC_CONTACT + k];
    if (a.grads.drive) for (int k = 0; k < ndp; ++k) a.grads.drive[(long long)design * ndp + k] = bad ? nanv : Sq0[SC_DRIVE + k];
    if (a.ts_bar) a.ts_bar[(long long)design * a.n_t] = bad ? nanv : Sq0[SC_T0];
    if (a.stats) {
      DfxStats st;
      st.steps = n_steps; st.accepted = n_acc; st.rhs_evals = n_rhs; st.status = status; st.reserved = 0; st.last_dt = h;
      a.stats[design] = st;
    }
  }
  if (NT > 0) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_sh), "r"(512));
  }
}

}  // namespace dfx
