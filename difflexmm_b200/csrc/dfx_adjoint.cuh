// dfx_adjoint.cuh -- continuous adjoint of the forward solve: what jax.grad derives for the
// odeint call at dynamics.py:166 (jax.experimental.ode._odeint_rev).
//
// One persistent CTA per design walks the stored outputs ys[i] backwards.  On every output
// interval it restarts a fresh adaptive Dopri5 solve (own initial step size) of the augmented
// system in negated time s = -t
//     z = (y, y_bar, t0_bar, args_bar)      dz/ds = (-f, vjp_y, vjp_t, vjp_args)
// whose RMS error norm runs over ALL entries of z, parameter cotangents included.  The stored
// outputs act as checkpoints: the forward state is recomputed from ys[i] inside each interval.
//
// Augmented RHS (three phases, two CTA barriers), with w = lambda_v / m:
//   A  per block-DOF : stage values of u, v, lambda_u, lambda_v; sin/cos(theta); w
//   B  per bond      : analytic gradient evaluated on dual numbers seeded with w
//                      -> force slots, (H w) slots, parameter cotangent integrands
//   C  per block-DOF : gather slots;  dv/ds = -(F - c v + load)/m,  dlambda_u/ds = -(H w),
//                      dlambda_v/ds = lambda_u - c w,  quadrature integrands for inertia, damping,
//                      centroid_node_vectors (contact chain), drive parameters and t0_bar
// Parameter cotangents are pure quadratures: their stage values never feed the RHS.  Stages 3..6 only STORE
// their integrand (no read-modify-write round trip to the L2-resident arrays); the solution / error / midpoint
// combinations are formed once, at the last stage.  Scalar leaves (k_stretch, ..., contact and drive parameters, t0_bar) are reduced with
// warp shuffles into per-warp partials and combined once per step.
#pragma once

#include "dfx_forward.cuh"

namespace dfx {

// arrays of the adjoint kernel in placement-priority order
enum {
  AA_US = 0, AA_WS, AA_VS, AA_LUS, AA_LVS,  // stage communication
  AA_FS, AA_HS, AA_GS, AA_GA,               // node slots
  AA_SC,                                    // scalar leaves (small)
  AA_INVM, AA_CD,
  AA_U0, AA_V0, AA_LU0, AA_LV0,
  AA_KV, AA_KLU, AA_KLV,
  AA_BONDC, AA_CNV, AA_ALPHA, AA_EDGED,
  AA_QK3, AA_QK4, AA_QK5, AA_QK6, AA_QK1, AA_QK7, AA_Q0, AA_QNEW,
  AA_COUNT
};

struct PlacementA { long long off[AA_COUNT]; };

constexpr int SCW = 16;    // per-warp partial slots per scalar leaf (CTAs have at most 512 threads = 16 warps)
constexpr int NSCAL = 13;  // 0 t0_bar | 1-3 k_stretch,k_shear,k_rot | 4 damping | 5-7 contact | 8-12 drive
constexpr int SC_T0 = 0, SC_KS = 1, SC_KSH = 2, SC_KR = 3, SC_DAMP = 4, SC_CONTACT = 5, SC_DRIVE = 8;

struct AdjArgs {
  DevTopo topo;
  DfxParams p;
  Tableau tab;
  PlacementA place;
  const double* ys; const double* ts; long long ts_bstride; int n_t;
  const double* g;            // cotangent of ys, or NULL when the kinetic objective generates it in the kernel
  // objective whose cotangent is formed in the kernel (g == NULL): see objective_cotangent()
  int obj_kind; const int* obj_ids; int obj_n; const double* obj_w; const double* obj_arm; long long obj_arm_bstride;
  double rtol, atol;
  long long aug_size;
  int init_step_variant; long long max_steps;
  double* y0_bar; double* ts_bar;
  DfxParamGrads grads;
  DfxStats* stats;
  double* scratch; long long scratch_per_design;
  // quadrature layout (entries)
  int qo_cnv, qo_ref, qo_ks, qo_ksh, qo_kr, qo_damp, qo_inertia, qo_cen, nq;
  int group;  // CL = 2: CTAs per design
  const int* order;  // DfxOptions.design_order
};

// Cotangent dJ/d ys[design][i][(is_v ? n_free : 0) + f] of the device objectives (include/dfx.h), w = weights[design]:
//   kinetic:  dJ/dv_f = w m_f v_f on the target DOFs
//   angular:  J = sum (arm + u) x (m v) + I omega  ->  dJ/du_x = w m_y v_y, dJ/du_y = -w m_x v_x,
//             dJ/dv_x = -w (arm_y + u_y) m_x, dJ/dv_y = w (arm_x + u_x) m_y, dJ/domega = w I
// position of free DOF f in the objective's target list, or -1
__device__ inline int objective_target_index(const AdjArgs& a, int f) {
  int k = -1;
  for (int q = 0; q < a.obj_n; ++q) if (a.obj_ids[q] == f) k = q;
  return k;
}
// k = objective_target_index(a, f), looked up once per thread by the fast kernels
__device__ inline double objective_cotangent_k(const AdjArgs& a, int design, int i, int f, bool is_v, int k) {
  if (k < 0) return 0.0;
  const int nf = a.topo.n_free;
  const double w = a.obj_w ? a.obj_w[design] : 1.0;
  const double* m = a.p.inertia.ptr + (long long)design * a.p.inertia.bstride;
  const double* y = a.ys + ((long long)design * a.n_t + i) * 2 * nf;
  if (a.obj_kind == DFX_OBJ_KINETIC) return is_v ? w * m[f] * y[nf + f] : 0.0;
  const int kb = k / 3, c = k - 3 * kb;
  const int fx = a.obj_ids[3 * kb], fy = a.obj_ids[3 * kb + 1];
  const double* arm = a.obj_arm + (long long)design * a.obj_arm_bstride + 2 * kb;
  if (!is_v) return c == 0 ? w * m[fy] * y[nf + fy] : (c == 1 ? -w * m[fx] * y[nf + fx] : 0.0);
  if (c == 0) return -w * (arm[1] + y[fy]) * m[fx];
  if (c == 1) return w * (arm[0] + y[fx]) * m[fy];
  return w * m[f];
}
__device__ inline double objective_cotangent(const AdjArgs& a, int design, int i, int f, bool is_v) {
  return objective_cotangent_k(a, design, i, f, is_v, objective_target_index(a, f));
}

struct QuadCtx {
  double *q0, *qnew, *k1, *k7, *ks[4];  // ks: stage values k3..k6
  const Tableau* tab;
  double h, atol, rtol, x;
  int mode;  // 0: k1 at interval start | 2..5: stage | 6: last stage | 7: initial-step probe
  bool crossing;
  double cs, ce, cm, cs0, ce0, cm0;  // tableau weights of this stage (and of k1 for mode 2)
};

__device__ __forceinline__ double interp_eval(double y0, double y1, double ymid, double d0, double d1, double x) {
  const double ca = -2. * d0 + 2. * d1 - 8. * y0 - 8. * y1 + 16. * ymid;
  const double cb = 5. * d0 - 3. * d1 + 18. * y0 + 14. * y1 - 32. * ymid;
  const double cc = -4. * d0 + d1 - 11. * y0 - 5. * y1 + 16. * ymid;
  return (((ca * x + cb) * x + cc) * x + d0) * x + y0;
}

// one elementwise quadrature entry receives its integrand value for the current stage
__device__ __forceinline__ void quad_update(const QuadCtx& c, int idx, double val, double& acc) {
  switch (c.mode) {
    case 0: c.k1[idx] = val; break;
    case 7: {
      const double sc = c.atol + fabs(c.q0[idx]) * c.rtol;
      const double d = (val - c.k1[idx]) / sc;
      acc += d * d;
    } break;
    case 6: {
      const Tableau& t = *c.tab;
      const double k1 = c.k1[idx], k3 = c.ks[0][idx], k4 = c.ks[1][idx], k5 = c.ks[2][idx], k6 = c.ks[3][idx];
      const double asol = t.c_sol[0] * k1 + t.c_sol[2] * k3 + t.c_sol[3] * k4 + t.c_sol[4] * k5 + t.c_sol[5] * k6;
      const double aerr = t.c_err[0] * k1 + t.c_err[2] * k3 + t.c_err[3] * k4 + t.c_err[4] * k5 + t.c_err[5] * k6 + t.c_err[6] * val;
      const double amid = t.c_mid[0] * k1 + t.c_mid[2] * k3 + t.c_mid[3] * k4 + t.c_mid[4] * k5 + t.c_mid[5] * k6 + t.c_mid[6] * val;
      const double q0 = c.q0[idx], q1 = q0 + c.h * asol;
      const double tol = c.atol + c.rtol * fmax(fabs(q0), fabs(q1));
      const double r = c.h * aerr / tol;
      acc += r * r;
      c.k7[idx] = val;
      c.qnew[idx] = c.crossing ? interp_eval(q0, q1, q0 + c.h * amid, c.h * k1, c.h * val, c.x) : q1;
    } break;
    default: c.ks[c.mode - 2][idx] = val; break;  // stages 3..6 of the step (mode 2..5)
  }
}

// scalar leaves: per-warp partial sums of the same linear recurrences (combined once per step)
struct ScalCtx {
  double *wk1, *wk7, *wsol, *werr, *wmid;  // [NSCAL][SCW], consecutive, private to the CTA
};
constexpr int kRedDoublesDev = 40;  // reduction scratch at the start of shared memory (kRedDoubles of dfx_api.cu)
constexpr int kScalDoubles = 2 * NSCAL + 5 * NSCAL * SCW;                 // Sq0, Sqnew + per-warp partials
constexpr int kClusterReserve = 8 + 2 * kMaxGroup + 5 * NSCAL * kMaxGroup;  // barrier counter, group-sum partials, CTA totals
__device__ __forceinline__ void scal_update(const QuadCtx& c, const ScalCtx& s, int which, double partial) {
  // most warps contribute nothing to the sparse scalars (contact, drive, t0): skip their shuffles
  const double v = __any_sync(0xffffffffu, partial != 0.0) ? warp_sum(partial) : 0.0;
  if ((threadIdx.x & 31) != 0) return;
  const int idx = which * SCW + (threadIdx.x >> 5);
  switch (c.mode) {
    case 0: s.wk1[idx] = v; break;
    case 7: s.wk7[idx] = v; break;
    case 2: {
      const double k1 = s.wk1[idx];
      s.wsol[idx] = c.cs0 * k1 + c.cs * v;
      s.werr[idx] = c.ce0 * k1 + c.ce * v;
      s.wmid[idx] = c.cm0 * k1 + c.cm * v;
    } break;
    case 6:
      s.werr[idx] += c.ce * v;
      s.wmid[idx] += c.cm * v;
      s.wk7[idx] = v;
      break;
    default:
      s.wsol[idx] += c.cs * v;
      s.werr[idx] += c.ce * v;
      s.wmid[idx] += c.cm * v;
      break;
  }
}

// Several scalar leaves at once: the N warp reductions are interleaved and lane k updates the per-warp partial of leaf
// which[k] (which[k] < 0: skipped), instead of N separate reductions each finished by lane 0.
template <int N>
__device__ __forceinline__ void scal_update_many(const QuadCtx& c, const ScalCtx& s, const int (&which)[N], double (&v)[N]) {
  bool any = false;
#pragma unroll
  for (int k = 0; k < N; ++k) any = any || (which[k] >= 0 && v[k] != 0.0);
  // warps that contribute nothing to these leaves still have to write their (zero) partial in the modes that assign
  const bool assign = c.mode == 0 || c.mode == 7 || c.mode == 2 || c.mode == 6;
  if (!__any_sync(0xffffffffu, any) && !assign) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  const int lane = threadIdx.x & 31;
  double mine = 0.0;
  int w = -1;
#pragma unroll
  for (int k = 0; k < N; ++k) if (lane == k) { mine = v[k]; w = which[k]; }
  if (w < 0) return;
  const int idx = w * SCW + (threadIdx.x >> 5);
  switch (c.mode) {
    case 0: s.wk1[idx] = mine; break;
    case 7: s.wk7[idx] = mine; break;
    case 2: {
      const double k1 = s.wk1[idx];
      s.wsol[idx] = c.cs0 * k1 + c.cs * mine;
      s.werr[idx] = c.ce0 * k1 + c.ce * mine;
      s.wmid[idx] = c.cm0 * k1 + c.cm * mine;
    } break;
    case 6:
      s.werr[idx] += c.ce * mine;
      s.wmid[idx] += c.cm * mine;
      s.wk7[idx] = mine;
      break;
    default:
      s.wsol[idx] += c.cs * mine;
      s.werr[idx] += c.ce * mine;
      s.wmid[idx] += c.cm * mine;
      break;
  }
}

// CL = 0: one CTA per design.  CL = 1: one thread-block cluster per design.  CL = 2: a group of co-resident CTAs of a
// cooperative launch per design (see forward_kernel).
template <int CL>
__global__ void __launch_bounds__(512, 1) adjoint_kernel(const __grid_constant__ AdjArgs a) {
  extern __shared__ double smem[];
  const DevTopo& T = a.topo;
  const Tableau& tab = a.tab;
  const int ncta = CL == 1 ? (int)cluster_nctarank() : (CL == 2 ? a.group : 1);
  const int crank = CL == 1 ? (int)cluster_ctarank() : (CL == 2 ? (int)(blockIdx.x % ncta) : 0);
  const int design = a.order ? a.order[blockIdx.x / ncta] : (int)(blockIdx.x / ncta);
  const int tid = crank * blockDim.x + threadIdx.x, nthr = ncta * blockDim.x;
  const int lane = threadIdx.x & 31, cwarp = threadIdx.x >> 5, cnwarp = (blockDim.x + 31) >> 5;
  const int NB = T.n_blocks, NN = T.n_nodes, ND = 3 * NB, NBONDS = T.n_bonds, npb = T.n_npb, nf = T.n_free;
  const int NQ = a.nq;
  double* red = smem;
  double* scratch = a.scratch ? a.scratch + (long long)design * a.scratch_per_design : nullptr;
  // CL: the scratch starts with the barrier counter, the group-sum partials and the per-CTA totals of the scalar
  // leaves [5][NSCAL][kMaxGroup] (kClusterReserve doubles)
  GroupCtx grp = {CL, crank, ncta, (unsigned long long*)scratch, 0ULL, scratch + 8, 0};
  double* ctot = scratch + 8 + 2 * kMaxGroup;
  auto SYNC = [&]() { if (CL) group_sync(grp); else __syncthreads(); };
  auto SUM = [&](double v) { return CL ? group_sum(v, red, grp) : block_sum(v, red); };
  // Work distribution: CTA `crank` owns a contiguous slice of the blocks (all three DOF rows of it) and of the bonds,
  // so that every CTA of a cluster / group gets the same share of each phase.  One CTA: e = i, b = i.
  const int perB = (T.n_blocks + ncta - 1) / ncta;
  const int blk0 = min(T.n_blocks, crank * perB), cntB = min(T.n_blocks, blk0 + perB) - blk0;
  const int perL = (T.n_bonds + ncta - 1) / ncta;
  const int bond0 = min(T.n_bonds, crank * perL), bond1 = min(T.n_bonds, bond0 + perL);
#define FOR_E(e) for (int _i = threadIdx.x, e = 0; _i < 3 * cntB && ((e = (_i / cntB) * T.n_blocks + blk0 + _i % cntB), true); _i += blockDim.x)
#define FOR_B(b) for (int b = bond0 + threadIdx.x; b < bond1; b += blockDim.x)
  auto P = [&](int i) -> double* {
    const long long o = a.place.off[i];
    return o >= 0 ? smem + o : scratch + (-(o + 1));
  };
  double *Us = P(AA_US), *Ws = P(AA_WS), *Vs = P(AA_VS), *Lus = P(AA_LUS), *Lvs = P(AA_LVS);
  double *Fs = P(AA_FS), *Hs = P(AA_HS), *Gs = P(AA_GS), *Ga = T.contact ? P(AA_GA) : nullptr;
  // [2][NSCAL] q0,qnew + [5][NSCAL][SCW] per-warp partials; always private to the CTA (shared memory in cluster mode)
  double* SC = CL ? smem + kRedDoublesDev : P(AA_SC);
  double *invm = P(AA_INVM), *cd = P(AA_CD);
  double *u0 = P(AA_U0), *v0 = P(AA_V0), *lu0 = P(AA_LU0), *lv0 = P(AA_LV0);
  double *kv = P(AA_KV), *klu = P(AA_KLU), *klv = P(AA_KLV);
  const bool c_angle = T.contact == DFX_CONTACT_ANGLE, c_dist = T.contact == DFX_CONTACT_DISTANCE;
  double *bondc = P(AA_BONDC), *cnv = P(AA_CNV), *alpha = c_angle ? P(AA_ALPHA) : nullptr;
  double* edged = c_angle ? P(AA_EDGED) : nullptr;  // [4][NN] d atan2(edge)/d(edge) of the next / previous edge of every node
  QuadCtx qc;
  qc.ks[0] = P(AA_QK3); qc.ks[1] = P(AA_QK4); qc.ks[2] = P(AA_QK5); qc.ks[3] = P(AA_QK6); qc.tab = &a.tab;
  qc.k1 = P(AA_QK1); qc.k7 = P(AA_QK7); qc.q0 = P(AA_Q0); qc.qnew = P(AA_QNEW);
  qc.atol = a.atol; qc.rtol = a.rtol; qc.crossing = false; qc.x = 0; qc.h = 0; qc.mode = 0;
  double* Sq0 = SC;
  double* Sqnew = SC + NSCAL;
  ScalCtx sc;
  sc.wk1 = SC + 2 * NSCAL; sc.wk7 = sc.wk1 + NSCAL * SCW; sc.wsol = sc.wk7 + NSCAL * SCW;
  sc.werr = sc.wsol + NSCAL * SCW; sc.wmid = sc.werr + NSCAL * SCW;

  const double* g_ks = leaf_ptr(a.p.k_stretch, design);
  const double* g_ksh = leaf_ptr(a.p.k_shear, design);
  const double* g_kr = leaf_ptr(a.p.k_rot, design);
  const double* g_contact = leaf_ptr(a.p.contact, design);
  const double* g_cen = leaf_ptr(a.p.block_centroids, design);  // distance-based contact only
  const double* g_drive = leaf_ptr(a.p.drive, design);
  const double* ts = a.ts + (long long)design * a.ts_bstride;
  const double* ys = a.ys + (long long)design * a.n_t * 2 * nf;
  const double* gg = a.g + (long long)design * a.n_t * 2 * nf;
  const double rtol = a.rtol, atol = a.atol;
  const bool ks_pb = a.p.k_per_bond[0], ksh_pb = a.p.k_per_bond[1], kr_pb = a.p.k_per_bond[2];
  const bool damp_pd = a.p.damping_per_dof != 0, has_damp = T.n_damped > 0 && a.p.damping.ptr != nullptr;
  const int ndp = T.n_drive_params;

  setup_design_constants(T, a.p, design, bondc, cnv, alpha, invm, cd, tid, nthr);
  if (edged) {
    const double* g_cnv = leaf_ptr(a.p.centroid_node_vectors, design);
    for (int n = tid; n < NN; n += nthr) {
      const int blk = n / npb, l = n - blk * npb;
      const int nn = blk * npb + (l + 1 == npb ? 0 : l + 1), np = blk * npb + (l == 0 ? npb - 1 : l - 1);
      const double rx = g_cnv[2 * n], ry = g_cnv[2 * n + 1];
      const double e1x = g_cnv[2 * nn] - rx, e1y = g_cnv[2 * nn + 1] - ry;  // own next edge (also nn's previous edge, reversed)
      const double e2x = g_cnv[2 * np] - rx, e2y = g_cnv[2 * np + 1] - ry;  // own previous edge (also np's next edge, reversed)
      const double i1 = 1.0 / (e1x * e1x + e1y * e1y), i2 = 1.0 / (e2x * e2x + e2y * e2y);
      // d atan2(e)/d e = (-e_y, e_x)/|e|^2
      edged[n] = -e1y * i1; edged[NN + n] = e1x * i1; edged[2 * NN + n] = -e2y * i2; edged[3 * NN + n] = e2x * i2;
    }
  }
  for (int i = tid; i < 3 * NN; i += nthr) { Fs[i] = 0.0; Hs[i] = 0.0; }
  for (int i = tid; i < 2 * NN; i += nthr) Gs[i] = 0.0;
  if (Ga) for (int i = tid; i < (c_dist ? 6 : 2) * NN; i += nthr) Ga[i] = 0.0;
  for (int i = tid; i < NQ; i += nthr) { qc.q0[i] = 0.0; qc.k1[i] = 0.0; qc.k7[i] = 0.0; qc.qnew[i] = 0.0; qc.ks[0][i] = 0.0; qc.ks[1][i] = 0.0; qc.ks[2][i] = 0.0; qc.ks[3][i] = 0.0; }
  for (int i = threadIdx.x; i < kScalDoubles; i += blockDim.x) SC[i] = 0.0;
  double cmin = 0, ccut = 0, ckc = 0;
  if (T.contact) { cmin = g_contact[0]; ccut = g_contact[1]; ckc = g_contact[2]; }
  // y_bar = g[-1]
  FOR_E(e) {
    const int j = e / NB, blk = e - j * NB;
    const int f = T.free_of_dof[3 * blk + j];
    lu0[e] = f >= 0 ? gg[(long long)(a.n_t - 1) * 2 * nf + f] : 0.0;
    lv0[e] = f >= 0 ? gg[(long long)(a.n_t - 1) * 2 * nf + nf + f] : 0.0;
  }
  SYNC();

  // ---- augmented RHS, phases B and C.  kidx: which k-slot (0..6) receives the dynamic derivatives
  auto aug_BC = [&](double time, double* kv_out, double* klu_out, double* klv_out) {
    const bool want_q = qc.mode != 1;
    SYNC();
    double p_ks = 0, p_ksh = 0, p_kr = 0, p_c0 = 0, p_c1 = 0, p_c2 = 0, probe = 0;
    FOR_B(b) {
      const int2 nd = T.bond_nodes[b], bl = T.bond_blocks[b];
      BlockState<Dual> s1, s2;
      make_block(Us[bl.x], Us[NB + bl.x], Us[2 * NB + bl.x], Us[3 * NB + bl.x], Us[4 * NB + bl.x], Ws[bl.x], Ws[NB + bl.x], Ws[2 * NB + bl.x], s1);
      make_block(Us[bl.y], Us[NB + bl.y], Us[2 * NB + bl.y], Us[3 * NB + bl.y], Us[4 * NB + bl.y], Ws[bl.y], Ws[NB + bl.y], Ws[2 * NB + bl.y], s2);
      BondConst bc = {bondc[b], bondc[NBONDS + b], bondc[2 * NBONDS + b], bondc[3 * NBONDS + b]};
      const double ks = g_ks[ks_pb ? b : 0], ksh = g_ksh[ksh_pb ? b : 0], kr = g_kr[kr_pb ? b : 0];
      BondOut<Dual> o;
      if (want_q) bond_gradient<Dual, true, true>(T.bond_energy, s1, s2, cnv[nd.x], cnv[NN + nd.x], cnv[nd.y], cnv[NN + nd.y], bc, ks, ksh, kr, o);
      else bond_gradient<Dual, false, true>(T.bond_energy, s1, s2, cnv[nd.x], cnv[NN + nd.x], cnv[nd.y], cnv[NN + nd.y], bc, ks, ksh, kr, o);
      Dual dgr1[2] = {Dual(0.0), Dual(0.0)}, dgr2[2] = {Dual(0.0), Dual(0.0)};  // distance contact: d/d(cnv) of the bond's own nodes
      if (c_dist) {
        double r6[6][2];
        const int ends[2] = {nd.x, nd.y};
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          const int n = ends[side], blk = n / npb, l = n - blk * npb;
          const int nn = blk * npb + (l + 1 == npb ? 0 : l + 1), np = blk * npb + (l == 0 ? npb - 1 : l - 1);
          r6[3 * side][0] = cnv[n]; r6[3 * side][1] = cnv[NN + n];
          r6[3 * side + 1][0] = cnv[nn]; r6[3 * side + 1][1] = cnv[NN + nn];
          r6[3 * side + 2][0] = cnv[np]; r6[3 * side + 2][1] = cnv[NN + np];
        }
        DistanceContactOut<Dual> dc;
        distance_contact<Dual>(s1, s2, g_cen + 2 * bl.x, g_cen + 2 * bl.y, r6, cmin, ccut, ckc, dc);
#pragma unroll
        for (int j = 0; j < 3; ++j) { o.f1[j] = o.f1[j] + dc.f1[j]; o.f2[j] = o.f2[j] + dc.f2[j]; }
        if (want_q) {
          // integrand of a leaf p: -(dual part of dE/dp).  Vertex cotangents of the next / previous vertex go to slots
          // indexed by the bond node (each node belongs to one bond: no conflicts); block centroids see the contact
          // part of the force pair.
          dgr1[0] = dc.gr[0][0]; dgr1[1] = dc.gr[0][1]; dgr2[0] = dc.gr[3][0]; dgr2[1] = dc.gr[3][1];
          Ga[nd.x] = -dc.gr[1][0].d; Ga[NN + nd.x] = -dc.gr[1][1].d; Ga[2 * NN + nd.x] = -dc.gr[2][0].d; Ga[3 * NN + nd.x] = -dc.gr[2][1].d;
          Ga[nd.y] = -dc.gr[4][0].d; Ga[NN + nd.y] = -dc.gr[4][1].d; Ga[2 * NN + nd.y] = -dc.gr[5][0].d; Ga[3 * NN + nd.y] = -dc.gr[5][1].d;
          Ga[4 * NN + nd.x] = -dc.f1[0].d; Ga[5 * NN + nd.x] = -dc.f1[1].d;
          Ga[4 * NN + nd.y] = -dc.f2[0].d; Ga[5 * NN + nd.y] = -dc.f2[1].d;
          p_c0 -= dc.gmin.d; p_c1 -= dc.gcut.d; p_c2 -= dc.gkc.d;
        }
      } else if (c_angle) {
        Dual psi1 = wrapT(s1.th - s2.th + (alpha[nd.x] - alpha[NN + nd.y]));
        Dual psi2 = wrapT(s2.th - s1.th + (alpha[nd.y] - alpha[NN + nd.x]));
        Dual e1, e2, m1, m2, c1, c2, k1, k2;
        contact_term<Dual>(psi1, cmin, ccut, ckc, e1, m1, c1, k1);
        contact_term<Dual>(psi2, cmin, ccut, ckc, e2, m2, c2, k2);
        o.f1[2] = o.f1[2] + e1 - e2;
        o.f2[2] = o.f2[2] + e2 - e1;
        if (want_q) {
          // dS/dalpha = -(dual part of dE/dalpha): alpha_1next:+e1, alpha_1prev:-e2, alpha_2next:+e2, alpha_2prev:-e1
          Ga[nd.x] = -e1.d; Ga[NN + nd.x] = e2.d; Ga[nd.y] = -e2.d; Ga[NN + nd.y] = e1.d;
          p_c0 -= m1.d + m2.d; p_c1 -= c1.d + c2.d; p_c2 -= k1.d + k2.d;
        }
      }
      Fs[nd.x] = -o.f1[0].v; Fs[NN + nd.x] = -o.f1[1].v; Fs[2 * NN + nd.x] = -o.f1[2].v;
      Fs[nd.y] = -o.f2[0].v; Fs[NN + nd.y] = -o.f2[1].v; Fs[2 * NN + nd.y] = -o.f2[2].v;
      Hs[nd.x] = o.f1[0].d; Hs[NN + nd.x] = o.f1[1].d; Hs[2 * NN + nd.x] = o.f1[2].d;
      Hs[nd.y] = o.f2[0].d; Hs[NN + nd.y] = o.f2[1].d; Hs[2 * NN + nd.y] = o.f2[2].d;
      if (want_q) {
        // d(w.F)/dp = -(dual part of dE/dp)
        Gs[nd.x] = -(o.gr1[0].d + dgr1[0].d); Gs[NN + nd.x] = -(o.gr1[1].d + dgr1[1].d);
        Gs[nd.y] = -(o.gr2[0].d + dgr2[0].d); Gs[NN + nd.y] = -(o.gr2[1].d + dgr2[1].d);
        quad_update(qc, a.qo_ref + b, -o.gr0[0].d, probe);
        quad_update(qc, a.qo_ref + NBONDS + b, -o.gr0[1].d, probe);
        if (ks_pb) quad_update(qc, a.qo_ks + b, -o.gks.d, probe); else p_ks -= o.gks.d;
        if (ksh_pb) quad_update(qc, a.qo_ksh + b, -o.gksh.d, probe); else p_ksh -= o.gksh.d;
        if (kr_pb) quad_update(qc, a.qo_kr + b, -o.gkr.d, probe); else p_kr -= o.gkr.d;
      }
    }
    if (want_q) {
      const int wb[6] = {ks_pb ? -1 : SC_KS, ksh_pb ? -1 : SC_KSH, kr_pb ? -1 : SC_KR,
                         T.contact ? SC_CONTACT : -1, T.contact ? SC_CONTACT + 1 : -1, T.contact ? SC_CONTACT + 2 : -1};
      double vb[6] = {p_ks, p_ksh, p_kr, p_c0, p_c1, p_c2};
      scal_update_many<6>(qc, sc, wb, vb);
    }
    SYNC();
    double ls = 0.0, lsd = 0.0;
    if (T.load_kind != DFX_LOAD_NONE) load_eval(T.load_kind, time, T.load_consts, ls, lsd);
    double p_t0 = 0, p_damp = 0, p_dr[DFX_MAX_DRIVE_PARAMS] = {0, 0, 0, 0, 0};
    FOR_E(e) {
      const int j = e / NB, blk = e - j * NB, dof = 3 * blk + j;
      double F = 0.0, HW = 0.0;
      {
        const double* fslot = Fs + (long long)j * NN + blk * npb;
        const double* hslot = Hs + (long long)j * NN + blk * npb;
        for (int l = 0; l < npb; ++l) { F += fslot[l]; HW += hslot[l]; }
      }
      const double im = invm[e];
      if (im != 0.0) {
        const double w = Ws[e], v = Vs[e], c = cd[e];
        double lm = 0.0;
        if (T.load_kind != DFX_LOAD_NONE) { lm = T.load_mul[dof]; F += lm * ls; }
        const double acc = (F - c * v) * im;
        kv_out[e] = -acc;
        klu_out[e] = -HW;
        klv_out[e] = Lus[e] - c * w;
        if (want_q) {
          quad_update(qc, a.qo_inertia + e, -w * acc, probe);
          if (has_damp && T.damp_slot[dof] >= 0) { if (damp_pd) quad_update(qc, a.qo_damp + e, -w * v, probe); else p_damp -= w * v; }
          p_t0 += w * lm * lsd;
        }
      } else {
        kv_out[e] = 0.0; klu_out[e] = 0.0; klv_out[e] = 0.0;
        if (want_q) {
          quad_update(qc, a.qo_inertia + e, 0.0, probe);
          const int c = T.cons_slot[dof];
          if (c >= 0 && T.drive_kind != DFX_DRIVE_ZERO) {
            DriveEval de;
            drive_eval(T.drive_kind, time, g_drive, true, de, T.table);
            const double v0_ = T.drive_vec0[c], v1_ = T.drive_vec1[c];
            p_t0 -= HW * (v0_ * de.sdot[0] + v1_ * de.sdot[1]);
            for (int q = 0; q < ndp; ++q) p_dr[q] -= HW * (v0_ * de.dsdp[0][q] + v1_ * de.dsdp[1][q]);
          }
        }
      }
      if (want_q && j < 2) {
        // centroid_node_vectors cotangent integrand of every node of this block, component j.  All loads first
        // (one round trip), then the updates.
        double vals[4];
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          vals[l] = 0.0;
          if (l < npb) {
            const int n = blk * npb + l;
            double val = Gs[j * NN + n];
            if (c_angle) {
              const int nn = blk * npb + (l + 1 == npb ? 0 : l + 1), np = blk * npb + (l == 0 ? npb - 1 : l - 1);
              const double d1 = edged[j * NN + n], d2 = edged[(2 + j) * NN + n];
              // own edges: d/dr_n = -d/de ; as far end of nn's previous edge (e = r_n - r_nn = -e1): also -d1
              val += -Ga[n] * d1 - Ga[NN + n] * d2 - Ga[NN + nn] * d1 - Ga[np] * d2;
            } else if (c_dist) {
              // this vertex is the "next" of its previous vertex and the "previous" of its next vertex
              const int nn = blk * npb + (l + 1 == npb ? 0 : l + 1), np = blk * npb + (l == 0 ? npb - 1 : l - 1);
              val += Ga[j * NN + np] + Ga[(2 + j) * NN + nn];
            }
            vals[l] = val;
          }
        }
#pragma unroll
        for (int l = 0; l < 4; ++l)
          if (l < npb) quad_update(qc, a.qo_cnv + j * NN + blk * npb + l, vals[l], probe);
        for (int l = 4; l < npb; ++l) {  // polygons with more than 4 vertices
          const int n = blk * npb + l;
          double val = Gs[j * NN + n];
          const int nn = blk * npb + (l + 1 == npb ? 0 : l + 1), np = n - 1;
          if (c_angle) {
            const double d1 = edged[j * NN + n], d2 = edged[(2 + j) * NN + n];
            val += -Ga[n] * d1 - Ga[NN + n] * d2 - Ga[NN + nn] * d1 - Ga[np] * d2;
          } else if (c_dist) {
            val += Ga[j * NN + np] + Ga[(2 + j) * NN + nn];
          }
          quad_update(qc, a.qo_cnv + j * NN + n, val, probe);
        }
        if (c_dist) {  // block_centroids leaf: contact part of the force on this block, component j
          double cb = 0.0;
          for (int l = 0; l < npb; ++l) cb += Ga[(4 + j) * NN + blk * npb + l];
          quad_update(qc, a.qo_cen + j * NB + blk, cb, probe);
        }
      }
    }
    if (want_q) {
      const int wc[7] = {SC_T0, (has_damp && !damp_pd) ? SC_DAMP : -1, ndp > 0 ? SC_DRIVE : -1, ndp > 1 ? SC_DRIVE + 1 : -1,
                         ndp > 2 ? SC_DRIVE + 2 : -1, ndp > 3 ? SC_DRIVE + 3 : -1, ndp > 4 ? SC_DRIVE + 4 : -1};
      double vc[7] = {p_t0, p_damp, p_dr[0], p_dr[1], p_dr[2], p_dr[3], p_dr[4]};
      scal_update_many<7>(qc, sc, wc, vc);
    }
    return probe;
  };

  // stage values of DOF e -> shared stage arrays (time = real time of the stage)
  auto put_stage = [&](int e, double u, double v, double lu, double lv, double time) {
    const int j = e / NB, blk = e - j * NB;
    const double im = invm[e];
    if (im == 0.0) {
      v = 0.0; lu = 0.0; lv = 0.0; u = 0.0;
      const int c = T.cons_slot[3 * blk + j];
      if (c >= 0 && T.drive_kind != DFX_DRIVE_ZERO) {
        DriveEval de;
        drive_eval(T.drive_kind, time, g_drive, false, de, T.table);
        u = T.drive_vec0[c] * de.s[0] + T.drive_vec1[c] * de.s[1];
      }
    }
    Us[e] = u; Vs[e] = v; Lus[e] = lu; Lvs[e] = lv; Ws[e] = lv * im;
    if (j == 2) {
      double sn, cs;
      sincos(u, &sn, &cs);
      Us[3 * NB + blk] = sn; Us[4 * NB + blk] = cs;
    }
  };

  // Scalar leaves: every warp integrates its own partial sums (arrays 0 wk1, 1 wk7, 2 wsol, 3 werr, 4 wmid).
  // scal_sync(mask): barrier after which the totals of the arrays in `mask` can be read with wtotal().  In cluster
  // mode each CTA first publishes its own totals (ctot); leaf `which` is owned by warp (which % cnwarp) of the first
  // CTA, whose lane 0 keeps Sq0 / Sqnew[which].
  auto scal_sync = [&](int mask) {
    __syncthreads();
    if (CL) {
      for (int w = cwarp; w < NSCAL; w += cnwarp)
        for (int k = 0; k < 5; ++k) {
          if (!((mask >> k) & 1)) continue;
          double v = lane < cnwarp ? sc.wk1[(k * NSCAL + w) * SCW + lane] : 0.0;
          v = warp_sum(v);
          if (lane == 0) ctot[(k * NSCAL + w) * kMaxGroup + crank] = v;
        }
      group_sync(grp);
    }
  };
  auto wtotal = [&](int k, int which) {  // whole warp
    double v;
    if (CL) { v = 0.0; for (int r = lane; r < ncta; r += 32) v += ctot[(k * NSCAL + which) * kMaxGroup + r]; }
    else v = lane < cnwarp ? sc.wk1[(k * NSCAL + which) * SCW + lane] : 0.0;
    return warp_sum(v);
  };

  long long n_steps = 0, n_acc = 0, n_rhs = 0;
  int status = 0;
  double h = 0.0;
  const long long n_pad_total = a.aug_size;  // full size of the reference's augmented state
  const double inv_n = 1.0 / (double)n_pad_total;

  for (int i = a.n_t - 1; i >= 1 && status == 0; --i) {
    // ---- restart: z0 = (ys[i], y_bar, t0_bar, args_bar) at s0 = -ts[i] --------------------------
    const double s0 = -ts[i], s_target = -ts[i - 1];
    const double* yi = ys + (long long)i * 2 * nf;
    const double* gi = gg + (long long)i * 2 * nf;
    FOR_E(e) {
      const int j = e / NB, blk = e - j * NB;
      const int f = T.free_of_dof[3 * blk + j];
      u0[e] = f >= 0 ? yi[f] : 0.0;
      v0[e] = f >= 0 ? yi[nf + f] : 0.0;
      put_stage(e, u0[e], v0[e], lu0[e], lv0[e], -s0);
    }
    qc.mode = 0;
    aug_BC(-s0, kv, klu, klv);
    n_rhs++;
    // t_bar = func(ys[i], ts[i]) . g[i];  func = (v, acc) = (v0, -kv[0])
    double pt = 0.0;
    FOR_E(e) {
      if (invm[e] == 0.0) continue;
      const int j = e / NB, blk = e - j * NB;
      const int f = T.free_of_dof[3 * blk + j];
      pt += v0[e] * gi[f] - kv[e] * gi[nf + f];
    }
    const double t_bar = SUM(pt);
    if (tid == 0) {
      if (a.ts_bar) a.ts_bar[(long long)design * a.n_t + i] = t_bar;
      Sq0[SC_T0] -= t_bar;
    }
    scal_sync(1);
    // ---- initial_step_size over the whole augmented vector ---------------------------------------
    {
      double sd0 = 0, sd1 = 0;
      FOR_E(e) {
        if (invm[e] == 0.0) continue;
        const double su = atol + fabs(u0[e]) * rtol, sv = atol + fabs(v0[e]) * rtol;
        const double slu = atol + fabs(lu0[e]) * rtol, slv = atol + fabs(lv0[e]) * rtol;
        const double a0 = u0[e] / su, a1 = v0[e] / sv, a2 = lu0[e] / slu, a3 = lv0[e] / slv;
        const double b0 = -v0[e] / su, b1 = kv[e] / sv, b2 = klu[e] / slu, b3 = klv[e] / slv;
        sd0 += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
        sd1 += b0 * b0 + b1 * b1 + b2 * b2 + b3 * b3;
      }
      for (int q = tid; q < NQ; q += nthr) {
        const double s = atol + fabs(qc.q0[q]) * rtol;
        const double a0 = qc.q0[q] / s, b0 = qc.k1[q] / s;
        sd0 += a0 * a0; sd1 += b0 * b0;
      }
      if (crank == 0) for (int w = cwarp; w < NSCAL; w += cnwarp) {
        const double k1 = wtotal(0, w);
        if (lane == 0) {
          const double s = atol + fabs(Sq0[w]) * rtol;
          const double a0 = Sq0[w] / s, b0 = k1 / s;
          sd0 += a0 * a0; sd1 += b0 * b0;
        }
      }
      const double d0 = sqrt(SUM(sd0));
      const double d1 = sqrt(SUM(sd1));
      const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
      FOR_E(e)
        put_stage(e, u0[e] - h0 * v0[e], v0[e] + h0 * kv[e], lu0[e] + h0 * klu[e], lv0[e] + h0 * klv[e], -(s0 + h0));
      qc.mode = 7;
      double sd2 = aug_BC(-(s0 + h0), kv + ND, klu + ND, klv + ND);
      n_rhs++;
      FOR_E(e) {
        if (invm[e] == 0.0) continue;
        const double su = atol + fabs(u0[e]) * rtol, sv = atol + fabs(v0[e]) * rtol;
        const double slu = atol + fabs(lu0[e]) * rtol, slv = atol + fabs(lv0[e]) * rtol;
        const double b0 = (-Vs[e] + v0[e]) / su, b1 = (kv[ND + e] - kv[e]) / sv;
        const double b2 = (klu[ND + e] - klu[e]) / slu, b3 = (klv[ND + e] - klv[e]) / slv;
        sd2 += b0 * b0 + b1 * b1 + b2 * b2 + b3 * b3;
      }
      scal_sync(2);  // per-warp probe partials (wk7) complete
      if (crank == 0) for (int w = cwarp; w < NSCAL; w += cnwarp) {
        const double k7 = wtotal(1, w), k1 = wtotal(0, w);
        if (lane == 0) {
          const double s = atol + fabs(Sq0[w]) * rtol;
          const double b0 = (k7 - k1) / s;
          sd2 += b0 * b0;
        }
      }
      const double d2 = sqrt(SUM(sd2)) / h0;
      double h1;
      if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
      else h1 = pow(0.01 / (a.init_step_variant == 0 ? d1 + d2 : fmax(d1, d2)), 0.2);
      h = fmin(100.0 * h0, h1);
    }
    // ---- adaptive steps until s >= s_target --------------------------------------------------------
    double s_cur = s0;
    long long istep = 0;
    bool done = !(s_cur < s_target);
    while (!done) {
      if (!(h > 0.0)) { status |= DFX_STATUS_DT_UNDERFLOW; break; }
      if (istep >= a.max_steps) { status |= DFX_STATUS_MAX_STEPS; break; }
      const double s_new = s_cur + h;
      qc.h = h;
      qc.crossing = !(s_new < s_target);
      qc.x = (s_target - s_cur) / (s_new - s_cur);
      double se = 0.0;
#pragma unroll 1
      for (int st = 0; st < 6; ++st) {
        const double ha = h * tab.alpha[st], h2 = h * h, s_stage = s_cur + ha;
        FOR_E(e) {
          double au = 0.0, av = 0.0, alu = 0.0, alv = 0.0;
          for (int l = 0; l <= st; ++l) {
            const double b = tab.beta[st][l];
            const double k = kv[l * ND + e];
            au = fma(tab.a2[st][l], k, au);
            av = fma(b, k, av);
            alu = fma(b, klu[l * ND + e], alu);
            alv = fma(b, klv[l * ND + e], alv);
          }
          put_stage(e, u0[e] - ha * v0[e] - h2 * au, v0[e] + h * av, lu0[e] + h * alu, lv0[e] + h * alv, -s_stage);
        }
        const int kidx = st + 1;
        qc.mode = kidx;
        qc.cs = tab.c_sol[kidx]; qc.ce = tab.c_err[kidx]; qc.cm = tab.c_mid[kidx];
        qc.cs0 = tab.c_sol[0]; qc.ce0 = tab.c_err[0]; qc.cm0 = tab.c_mid[0];
        se += aug_BC(-s_stage, kv + kidx * ND, klu + kidx * ND, klv + kidx * ND);
      }
      n_rhs += 6;
      // error of the dynamic entries
      FOR_E(e) {
        if (invm[e] == 0.0) continue;
        double eu = 0.0, ev = 0.0, elu = 0.0, elv = 0.0;
#pragma unroll
        for (int l = 0; l < 7; ++l) {
          const double k = kv[l * ND + e];
          eu = fma(tab.e2[l], k, eu);
          ev = fma(tab.c_err[l], k, ev);
          elu = fma(tab.c_err[l], klu[l * ND + e], elu);
          elv = fma(tab.c_err[l], klv[l * ND + e], elv);
        }
        eu = -h * (tab.sum_err * v0[e] + h * eu);
        ev *= h; elu *= h; elv *= h;
        const double r0 = eu / (atol + rtol * fmax(fabs(u0[e]), fabs(Us[e])));
        const double r1 = ev / (atol + rtol * fmax(fabs(v0[e]), fabs(Vs[e])));
        const double r2 = elu / (atol + rtol * fmax(fabs(lu0[e]), fabs(Lus[e])));
        const double r3 = elv / (atol + rtol * fmax(fabs(lv0[e]), fabs(Lvs[e])));
        se += r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3;
      }
      scal_sync(31);  // per-warp scalar partials of the last stage complete
      if (crank == 0) for (int w = cwarp; w < NSCAL; w += cnwarp) {
        const double k1 = wtotal(0, w), k7 = wtotal(1, w);
        const double tsol = wtotal(2, w), terr = wtotal(3, w), tmid = wtotal(4, w);
        if (lane == 0) {
          const double q0 = Sq0[w];
          const double q1 = q0 + h * tsol;
          const double r = h * terr / (atol + rtol * fmax(fabs(q0), fabs(q1)));
          se += r * r;
          Sqnew[w] = qc.crossing ? interp_eval(q0, q1, q0 + h * tmid, h * k1, h * k7, qc.x) : q1;
        }
      }
      const double ratio = sqrt(SUM(se) * inv_n);
      ++n_steps; ++istep;
      if (!isfinite(ratio)) { status |= DFX_STATUS_NONFINITE; break; }
      if (ratio <= 1.0) {
        ++n_acc;
        if (qc.crossing) {
          // interval finished: keep the interpolated cotangents at s_target, add g[i-1]
          const double* gp = gg + (long long)(i - 1) * 2 * nf;
          FOR_E(e) {
            if (invm[e] == 0.0) continue;
            const int j = e / NB, blk = e - j * NB;
            const int f = T.free_of_dof[3 * blk + j];
            double mlu = 0.0, mlv = 0.0;
#pragma unroll
            for (int l = 0; l < 7; ++l) {
              mlu = fma(tab.c_mid[l], klu[l * ND + e], mlu);
              mlv = fma(tab.c_mid[l], klv[l * ND + e], mlv);
            }
            const double nlu = interp_eval(lu0[e], Lus[e], lu0[e] + h * mlu, h * klu[e], h * klu[6 * ND + e], qc.x);
            const double nlv = interp_eval(lv0[e], Lvs[e], lv0[e] + h * mlv, h * klv[e], h * klv[6 * ND + e], qc.x);
            lu0[e] = nlu + gp[f];
            lv0[e] = nlv + gp[nf + f];
          }
          done = true;
        } else {
          FOR_E(e) {
            u0[e] = Us[e]; v0[e] = Vs[e]; lu0[e] = Lus[e]; lv0[e] = Lvs[e];
            kv[e] = kv[6 * ND + e]; klu[e] = klu[6 * ND + e]; klv[e] = klv[6 * ND + e];
          }
          if (lane == 0) for (int w = 0; w < NSCAL; ++w) sc.wk1[w * SCW + cwarp] = sc.wk7[w * SCW + cwarp];
          double* tmp = qc.k1; qc.k1 = qc.k7; qc.k7 = tmp;
        }
        if (crank == 0 && lane == 0) for (int w = cwarp; w < NSCAL; w += cnwarp) Sq0[w] = Sqnew[w];
        double* tmp = qc.q0; qc.q0 = qc.qnew; qc.qnew = tmp;
        s_cur = s_new;
      }
      const double dfactor = ratio < 1.0 ? 1.0 : 0.2;
      const double factor = fmin(10.0, fmax(pow(ratio, -0.2) * 0.9, dfactor));
      h = (ratio == 0.0) ? h * 10.0 : h * factor;
      SYNC();
    }
  }

  // ---- outputs ---------------------------------------------------------------------------------------
  SYNC();
  const double nanv = nan("");
  const bool bad = status != 0;
  FOR_E(e) {
    const int j = e / NB, blk = e - j * NB, dof = 3 * blk + j;
    const int f = T.free_of_dof[dof];
    if (a.grads.damping && has_damp && damp_pd) {
      // entries of the damping leaf that sit on constrained DOFs have zero cotangent (q0 stays 0 there)
      const int ds = T.damp_slot[dof];
      if (ds >= 0) a.grads.damping[(long long)design * T.n_damped * 3 + ds] = bad ? nanv : qc.q0[a.qo_damp + e];
    }
    if (f < 0) continue;
    if (a.y0_bar) {
      a.y0_bar[(long long)design * 2 * nf + f] = bad ? nanv : lu0[e];
      a.y0_bar[(long long)design * 2 * nf + nf + f] = bad ? nanv : lv0[e];
    }
    if (a.grads.inertia) a.grads.inertia[(long long)design * nf + f] = bad ? nanv : qc.q0[a.qo_inertia + e];
  }
  if (a.grads.centroid_node_vectors)
    for (int n = tid; n < NN; n += nthr) {
      a.grads.centroid_node_vectors[((long long)design * NN + n) * 2] = bad ? nanv : qc.q0[a.qo_cnv + n];
      a.grads.centroid_node_vectors[((long long)design * NN + n) * 2 + 1] = bad ? nanv : qc.q0[a.qo_cnv + NN + n];
    }
  if (a.grads.block_centroids && c_dist)
    for (int k = tid; k < NB; k += nthr) {
      a.grads.block_centroids[((long long)design * NB + k) * 2] = bad ? nanv : qc.q0[a.qo_cen + k];
      a.grads.block_centroids[((long long)design * NB + k) * 2 + 1] = bad ? nanv : qc.q0[a.qo_cen + NB + k];
    }
  if (a.grads.reference_vector)
    FOR_B(b) {
      a.grads.reference_vector[((long long)design * NBONDS + b) * 2] = bad ? nanv : qc.q0[a.qo_ref + b];
      a.grads.reference_vector[((long long)design * NBONDS + b) * 2 + 1] = bad ? nanv : qc.q0[a.qo_ref + NBONDS + b];
    }
  {
    double* outs[3] = {a.grads.k_stretch, a.grads.k_shear, a.grads.k_rot};
    const bool pb[3] = {ks_pb, ksh_pb, kr_pb};
    const int qo[3] = {a.qo_ks, a.qo_ksh, a.qo_kr};
    for (int k = 0; k < 3; ++k) {
      if (!outs[k]) continue;
      if (pb[k]) { FOR_B(b) outs[k][(long long)design * NBONDS + b] = bad ? nanv : qc.q0[qo[k] + b]; }
      else if (tid == 0) outs[k][design] = bad ? nanv : Sq0[SC_KS + k];
    }
  }
  if (tid == 0) {
    if (a.grads.damping && has_damp && !damp_pd) a.grads.damping[design] = bad ? nanv : Sq0[SC_DAMP];
    if (a.grads.contact && T.contact) for (int k = 0; k < 3; ++k) a.grads.contact[(long long)design * 3 + k] = bad ? nanv : Sq0[SC_CONTACT + k];
    if (a.grads.drive) for (int k = 0; k < ndp; ++k) a.grads.drive[(long long)design * ndp + k] = bad ? nanv : Sq0[SC_DRIVE + k];
    if (a.ts_bar) a.ts_bar[(long long)design * a.n_t] = bad ? nanv : Sq0[SC_T0];
    if (a.stats) {
      DfxStats st;
      st.steps = n_steps; st.accepted = n_acc; st.rhs_evals = n_rhs; st.status = status; st.reserved = 0; st.last_dt = h;
      a.stats[design] = st;
    }
  }
}

#undef FOR_E
#undef FOR_B

}  // namespace dfx
