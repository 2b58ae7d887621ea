// dfx_forward2.cuh -- the fast forward kernel (same algorithm as dfx_forward.cuh: adaptive Dopri5 of
// jax.experimental.ode with the RHS of dynamics.py:33-55 inlined; see there).  Same structure as the fast adjoint
// (dfx_adjoint2.cuh): thread t owns rigid unit t and bonds t, t+T; the unit's state, its 7-stage derivative
// history and most constants are thread private in tensor memory (tcgen05.ld/st), the rest in shared memory; only
// the stage state and four numbers per bond (the end forces are equal and opposite) cross threads, with two CTA
// barriers per RHS evaluation.  The footprint is small enough for TWO designs per SM (2 x 256 TMEM columns,
// < 113 KB shared, <= 85 registers), which doubles the warps available to hide FP64 latency.
//
// Preconditions (checked by the host, else the generic kernel runs): n_blocks <= T, n_bonds <= 2T, n_npb <= 4.
#pragma once

#include "dfx_adjoint2.cuh"

namespace dfx {

constexpr int F_KV = 0, F_U0 = 21, F_V0 = 24, F_TV = 27;  // history [7][3], state at the step start, parked stage velocity
constexpr int F_INVM = 30, F_CD = 33, F_BOND = 36;        // constants; per bond BC_N values (as in the adjoint)
constexpr int F_KPBC = F_BOND + 2 * BC_N;                 // 56: per-bond stiffness values (per-bond leaves only)
constexpr int F_NSLOT = F_KPBC + 6;                       // 62

struct Fwd2Args {
  FwdArgs a;             // same inputs / outputs as the generic kernel (placement unused)
  const int* node_bond;  // [n_nodes] bond*2+side or -1
  long long tp_scratch_per_design;
  int ns_slots, tmem_cols_per_warp, tmem_cols_alloc;
};

// NS >= 0: the number of thread-private slots in shared memory is a compile-time constant (the tier tests of the store
// disappear); NS < 0: read from the launch arguments.
template <int NT, int TT, int CTAS, int NS = -1>
__global__ void __launch_bounds__(TT, CTAS) forward2_kernel(const __grid_constant__ Fwd2Args A) {
  extern __shared__ double smem[];
  __shared__ uint32_t tmem_base_sh;
  const FwdArgs& a = A.a;
  const DevTopo& T = a.topo;
  const Tableau& tab = a.tab;
  const int design = a.order ? a.order[blockIdx.x] : (int)blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int nthr = TT;
  const int NB = T.n_blocks, NN = T.n_nodes, NBONDS = T.n_bonds, npb = T.n_npb, nf = T.n_free;
  constexpr int NBS = TT, NDS = 2 * TT;  // compile-time strides of the shared arrays

  // ---- shared memory: red[40] | Us[5][T] | SL[4][2T] | drv[4] | nodeb[NN] | tp[ns][T]
  double* red = smem;
  double* Us = red + 40;
  double* SL = Us + 5 * NBS;  // per bond: gdx gdy T1 T2
  double* drv = SL + 4 * NDS;
  int* nodeb = (int*)(drv + 4);
  double* tp_s = (double*)(nodeb + ((NN + 1) & ~1));

  if (NT > 0) {
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       (uint32_t)__cvta_generic_to_shared(&tmem_base_sh)), "r"(A.tmem_cols_alloc));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  TP<NT, NS, TT> tp;
  tp.ns = A.ns_slots;
  tp.s = tp_s + tid;
  tp.g = a.scratch + (long long)design * A.tp_scratch_per_design + tid;
  tp.taddr = NT > 0 ? (tmem_base_sh + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * A.tmem_cols_per_warp)) : 0u;

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
  const double* y0g = a.y0 + (long long)design * a.y0_bstride;
  double* ys = a.ys + (long long)design * a.n_t * 2 * nf;
  const double rtol = a.rtol, atol = a.atol;
  const bool ks_pb = a.p.k_per_bond[0], ksh_pb = a.p.k_per_bond[1], kr_pb = a.p.k_per_bond[2];
  const bool any_pb = ks_pb || ksh_pb || kr_pb;
  const bool damp_pd = a.p.damping_per_dof != 0, has_damp = T.n_damped > 0 && g_damp != nullptr;
  const bool contact = T.contact != 0;
  double cmin = 0, ccut = 0, ckc = 0;
  if (contact) { cmin = g_contact[0]; ccut = g_contact[1]; ckc = g_contact[2]; }
  const double ks_u = g_ks[0], ksh_u = g_ksh[0], kr_u = g_kr[0];

  int fidx[3], cslot[3];
  bool is_free[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int dof = 3 * blk + j;
    fidx[j] = T.free_of_dof[dof];
    is_free[j] = has_blk && fidx[j] >= 0;
    cslot[j] = has_blk ? T.cons_slot[dof] : -1;
  }
  const bool has_cons = cslot[0] >= 0 || cslot[1] >= 0 || cslot[2] >= 0;
  const int2 bb0 = T.bond_blocks[bnd[0]], bb1 = T.bond_blocks[bnd[1]];

  for (int i = tid; i < NN; i += nthr) nodeb[i] = A.node_bond[i];
  for (int i = tid; i < 4 * NDS; i += nthr) SL[i] = 0.0;
  if (tid == 0) { drv[0] = 0.0; drv[1] = 0.0; drv[2] = nan(""); }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int ds = T.damp_slot[3 * blk + j];
    tp.st(F_INVM + j, is_free[j] ? 1.0 / g_inertia[fidx[j]] : 0.0);
    tp.st(F_CD + j, (is_free[j] && ds >= 0 && has_damp) ? (damp_pd ? g_damp[ds] : g_damp[0]) : 0.0);
    const double u = is_free[j] ? y0g[fidx[j]] : 0.0, v = is_free[j] ? y0g[nf + fidx[j]] : 0.0;
    tp.st(F_U0 + j, u); tp.st(F_V0 + j, v);
    if (is_free[j]) { ys[fidx[j]] = u; ys[nf + fidx[j]] = v; }
  }
  auto edge_angle = [&](int n, int dir) {
    const int b = n / npb, l = n - b * npb;
    const int m = b * npb + (dir > 0 ? (l + 1 == npb ? 0 : l + 1) : (l == 0 ? npb - 1 : l - 1));
    return atan2(g_cnv[2 * m + 1] - g_cnv[2 * n + 1], g_cnv[2 * m] - g_cnv[2 * n]);
  };
#pragma unroll 1
  for (int i = 0; i < 2; ++i) {
    const int s0 = F_BOND + i * BC_N, b = bnd[i];
    const int2 nd = T.bond_nodes[b];
    const double rx = g_ref[2 * b], ry = g_ref[2 * b + 1];
    tp.st(s0 + BC_R0X, rx); tp.st(s0 + BC_R0Y, ry);
    tp.st(s0 + BC_L0, sqrt(rx * rx + ry * ry)); tp.st(s0 + BC_IL0, 1.0 / sqrt(rx * rx + ry * ry));
    tp.st(s0 + BC_R1X, g_cnv[2 * nd.x]); tp.st(s0 + BC_R1Y, g_cnv[2 * nd.x + 1]);
    tp.st(s0 + BC_R2X, g_cnv[2 * nd.y]); tp.st(s0 + BC_R2Y, g_cnv[2 * nd.y + 1]);
    double da1 = 0.0, da2 = 0.0;
    if (contact) {
      da1 = edge_angle(nd.x, +1) - edge_angle(nd.y, -1);
      da2 = edge_angle(nd.y, +1) - edge_angle(nd.x, -1);
    }
    tp.st(s0 + BC_DA1, da1); tp.st(s0 + BC_DA2, da2);
    if (any_pb) {
      tp.st(F_KPBC + 3 * i, g_ks[ks_pb ? b : 0]); tp.st(F_KPBC + 3 * i + 1, g_ksh[ksh_pb ? b : 0]);
      tp.st(F_KPBC + 3 * i + 2, g_kr[kr_pb ? b : 0]);
    }
  }
  tp.fence_st();
  __syncthreads();

  // ---- one RHS evaluation (phases B and C); the stage velocity is parked in the store by publish()
  auto rhs_BC = [&](double time, int kidx, double time_next) {
    __syncthreads();
    if (tid == nthr - 1 && T.drive_kind != DFX_DRIVE_ZERO) {  // drive channels of the next evaluation
      DriveEval de;
      drive_eval(T.drive_kind, time_next, g_drive, false, de, T.table);
      drv[0] = de.s[0]; drv[1] = de.s[1]; drv[2] = time_next;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int b = bnd[i];
      const int b1 = i == 0 ? bb0.x : bb1.x, b2 = i == 0 ? bb0.y : bb1.y;
      double c[BC_N];
      tp.template ldn<BC_N>(F_BOND + i * BC_N, 1, c);
      double ks = ks_u, ksh = ksh_u, kr = kr_u;
      if (any_pb) { ks = tp.ld(F_KPBC + 3 * i); ksh = tp.ld(F_KPBC + 3 * i + 1); kr = tp.ld(F_KPBC + 3 * i + 2); }
      BlockState<double> s1, s2;
      make_block(Us[b1], Us[NBS + b1], Us[2 * NBS + b1], Us[3 * NBS + b1], Us[4 * NBS + b1], s1);
      make_block(Us[b2], Us[NBS + b2], Us[2 * NBS + b2], Us[3 * NBS + b2], Us[4 * NBS + b2], s2);
      BondConst bc = {c[BC_R0X], c[BC_R0Y], c[BC_L0], c[BC_IL0]};
      BondOut<double> o;
      bond_gradient<double, false>(T.bond_energy, s1, s2, c[BC_R1X], c[BC_R1Y], c[BC_R2X], c[BC_R2Y], bc, ks, ksh, kr, o);
      if (contact) {
        const double psi1 = wrap_value(s1.th - s2.th + c[BC_DA1]), psi2 = wrap_value(s2.th - s1.th + c[BC_DA2]);
        if ((!(psi1 < cmin) && psi1 < ccut) || (!(psi2 < cmin) && psi2 < ccut)) {
          double e1, e2, d0, d1, d2;
          contact_term<double>(psi1, cmin, ccut, ckc, e1, d0, d1, d2);
          contact_term<double>(psi2, cmin, ccut, ckc, e2, d0, d1, d2);
          o.f1[2] += e1 - e2;
          o.f2[2] += e2 - e1;
        }
      }
      if (has_bnd[i]) { SL[b] = o.f2[0]; SL[NDS + b] = o.f2[1]; SL[2 * NDS + b] = -o.f1[2]; SL[3 * NDS + b] = -o.f2[2]; }
    }
    __syncthreads();
    double F[3] = {0, 0, 0};
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      if (l < npb) {
        const int nb_ = nodeb[blk * npb + l];
        if (nb_ >= 0) {
          const int b = nb_ >> 1;
          const bool second = nb_ & 1;
          const double sg = second ? -1.0 : 1.0;
          F[0] += sg * SL[b]; F[1] += sg * SL[NDS + b]; F[2] += SL[(second ? 3 : 2) * NDS + b];
        }
      }
    }
    double ls = 0.0, lsd = 0.0;
    if (T.load_kind != DFX_LOAD_NONE) load_eval(T.load_kind, time, T.load_consts, ls, lsd);
    double im[3], cdv[3], vst[3];
    tp.template ldn<3>(F_INVM, 1, im);
    tp.template ldn<3>(F_CD, 1, cdv);
    tp.template ldn<3>(F_TV, 1, vst);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double Fj = F[j];
      if (T.load_kind != DFX_LOAD_NONE && is_free[j]) Fj += T.load_mul[3 * blk + j] * ls;
      tp.template st_below<21>(F_KV + kidx * 3 + j, is_free[j] ? (Fj - cdv[j] * vst[j]) * im[j] : 0.0);
    }
    tp.fence_st();
  };

  auto publish = [&](double (&u)[3], double (&v)[3], double time) {
    if (has_cons && T.drive_kind != DFX_DRIVE_ZERO) {
      double s0_, s1_;
      if (drv[2] == time) { s0_ = drv[0]; s1_ = drv[1]; }
      else {
        DriveEval de;
        drive_eval(T.drive_kind, time, g_drive, false, de, T.table);
        s0_ = de.s[0]; s1_ = de.s[1];
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) if (cslot[j] >= 0) u[j] = T.drive_vec0[cslot[j]] * s0_ + T.drive_vec1[cslot[j]] * s1_;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (!is_free[j]) { v[j] = 0.0; if (cslot[j] < 0) u[j] = 0.0; }
      tp.st(F_TV + j, v[j]);
    }
    if (has_blk) {
      double sn, cs;
      sincos_fast(u[2], &sn, &cs);
      Us[blk] = u[0]; Us[NBS + blk] = u[1]; Us[2 * NBS + blk] = u[2]; Us[3 * NBS + blk] = sn; Us[4 * NBS + blk] = cs;
    }
    tp.fence_st();
  };

  long long n_steps = 0, n_acc = 0, n_rhs = 0, istep = 0;
  int status = 0;
  double t = ts[0], dt = 0.0, h0 = 0.0, d0 = 0.0, d1 = 0.0;
  const double inv_n = 1.0 / (2.0 * nf);
  constexpr int EV_INIT = 6, EV_PROBE = 7;
  int ev = EV_INIT, it = 1;
  bool running = a.n_t > 1;
  if (running && !(t < ts[1])) {  // repeated first output time: jax yields 0/0 here
    for (; it < a.n_t && !(t < ts[it]); ++it)
      for (int j = 0; j < 3; ++j) if (is_free[j]) { ys[(long long)it * 2 * nf + fidx[j]] = nan(""); ys[(long long)it * 2 * nf + nf + fidx[j]] = nan(""); }
    running = it < a.n_t;
  }

  while (running) {
    double us[3], vs[3], time;
    int kidx;
    if (ev == EV_INIT) {
      tp.template ldn<3>(F_U0, 1, us);
      tp.template ldn<3>(F_V0, 1, vs);
      time = t; kidx = 0;
    } else if (ev == EV_PROBE) {
      double y0[6], k0[3];
      tp.template ldn<6>(F_U0, 1, y0);
      tp.template ldn<3>(F_KV, 1, k0);
#pragma unroll
      for (int j = 0; j < 3; ++j) { us[j] = y0[j] + h0 * y0[3 + j]; vs[j] = y0[3 + j] + h0 * k0[j]; }
      time = t + h0; kidx = 1;
    } else {
      const double ha = dt * tab.alpha[ev], h2 = dt * dt;
      double au[3] = {0, 0, 0}, av[3] = {0, 0, 0};
      // history stages 0..ev with compile-time tableau indices (constant-bank operands) and one TMEM wait per three stages
      auto acc = [&](auto st_tag, auto l0_tag, auto l1_tag) {
        constexpr int ST = decltype(st_tag)::value, L0 = decltype(l0_tag)::value, L1 = decltype(l1_tag)::value;
        double kv[3 * (L1 - L0 + 1)];
        tp.template ldn_below<21, 3 * (L1 - L0 + 1)>(F_KV + 3 * L0, 1, kv);
#pragma unroll
        for (int l = L0; l <= L1; ++l) {
          const double b = tab.beta[ST][l], b2 = tab.a2[ST][l];
#pragma unroll
          for (int j = 0; j < 3; ++j) { au[j] = fma(b2, kv[3 * (l - L0) + j], au[j]); av[j] = fma(b, kv[3 * (l - L0) + j], av[j]); }
        }
      };
      using std::integral_constant;
#define DFX_F2_ACC(ST, L0, L1) acc(integral_constant<int, ST>{}, integral_constant<int, L0>{}, integral_constant<int, L1>{})
      switch (ev) {
        case 0: DFX_F2_ACC(0, 0, 0); break;
        case 1: DFX_F2_ACC(1, 0, 1); break;
        case 2: DFX_F2_ACC(2, 0, 2); break;
        case 3: DFX_F2_ACC(3, 0, 2); DFX_F2_ACC(3, 3, 3); break;
        case 4: DFX_F2_ACC(4, 0, 2); DFX_F2_ACC(4, 3, 4); break;
        default: DFX_F2_ACC(5, 0, 2); DFX_F2_ACC(5, 3, 5); break;
      }
#undef DFX_F2_ACC
      double y0[6];
      tp.template ldn<6>(F_U0, 1, y0);
#pragma unroll
      for (int j = 0; j < 3; ++j) { us[j] = y0[j] + ha * y0[3 + j] + h2 * au[j]; vs[j] = y0[3 + j] + dt * av[j]; }
      time = t + ha; kidx = ev + 1;
    }
    publish(us, vs, time);
    const double time_next = ev < 5 ? t + dt * tab.alpha[ev + 1] : nan("");
    rhs_BC(time, kidx, time_next);
    n_rhs++;

    if (ev == EV_INIT) {
      double y0[6], k0[3], sd0 = 0, sd1 = 0;
      tp.template ldn<6>(F_U0, 1, y0);
      tp.template ldn<3>(F_KV, 1, k0);
#pragma unroll
      for (int j = 0; j < 3; ++j) if (is_free[j]) {
        const double su = atol + fabs(y0[j]) * rtol, sv = atol + fabs(y0[3 + j]) * rtol;
        const double a0 = y0[j] / su, a1 = y0[3 + j] / sv, b0 = y0[3 + j] / su, b1 = k0[j] / sv;
        sd0 += a0 * a0 + a1 * a1; sd1 += b0 * b0 + b1 * b1;
      }
      d0 = sqrt(block_sum(sd0, red));
      d1 = sqrt(block_sum(sd1, red));
      h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
      ev = EV_PROBE;
    } else if (ev == EV_PROBE) {
      double y0[6], k0[3], k1[3], sd2 = 0;
      tp.template ldn<6>(F_U0, 1, y0);
      tp.template ldn<3>(F_KV, 1, k0);
      tp.template ldn<3>(F_KV + 3, 1, k1);
#pragma unroll
      for (int j = 0; j < 3; ++j) if (is_free[j]) {
        const double su = atol + fabs(y0[j]) * rtol, sv = atol + fabs(y0[3 + j]) * rtol;
        const double b0 = (vs[j] - y0[3 + j]) / su, b1 = (k1[j] - k0[j]) / sv;
        sd2 += b0 * b0 + b1 * b1;
      }
      const double d2 = sqrt(block_sum(sd2, red)) / h0;
      double h1;
      if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
      else h1 = pow(0.01 / (a.init_step_variant == 0 ? d1 + d2 : fmax(d1, d2)), 0.2);
      dt = fmin(100.0 * h0, h1);
      istep = 0;
      if (!(dt > 0.0)) { status |= DFX_STATUS_DT_UNDERFLOW; running = false; }
      ev = 0;
    } else if (ev < 5) {
      ++ev;
    } else {
      // solution, error estimate, step control (jax: mean_error_ratio / optimal_step_size)
      double y1[6], se = 0.0;
      double kv[3][7], y0[6];
      tp.template ldn<6>(F_U0, 1, y0);
#pragma unroll
      for (int j = 0; j < 3; ++j) tp.template ldn<7>(F_KV + j, 3, kv[j]);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double eu = 0.0, evv = 0.0, su = 0.0, sv = 0.0;
#pragma unroll
        for (int l = 0; l < 7; ++l) {
          eu = fma(tab.e2[l], kv[j][l], eu); evv = fma(tab.c_err[l], kv[j][l], evv);
          su = fma(tab.s2[l], kv[j][l], su); sv = fma(tab.c_sol[l], kv[j][l], sv);
        }
        y1[j] = y0[j] + dt * (tab.sum_sol * y0[3 + j] + dt * su);
        y1[3 + j] = y0[3 + j] + dt * sv;
        eu = dt * (tab.sum_err * y0[3 + j] + dt * eu);
        evv *= dt;
        if (is_free[j]) {
          const double ru = eu * rcp_pos(atol + rtol * fmax(fabs(y0[j]), fabs(y1[j])));
          const double rv = evv * rcp_pos(atol + rtol * fmax(fabs(y0[3 + j]), fabs(y1[3 + j])));
          se += ru * ru + rv * rv;
        }
      }
      const double ratio = sqrt(block_sum(se, red) * inv_n);
      ++n_steps; ++istep;
      if (!isfinite(ratio)) { status |= DFX_STATUS_NONFINITE; running = false; }
      else {
        if (ratio <= 1.0) {
          const double t_new = t + dt;
          // dense output for every requested time inside (t, t_new]
          while (it < a.n_t && !(t_new < ts[it])) {
            const double x = (ts[it] - t) / (t_new - t);
            double* out = ys + (long long)it * 2 * nf;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              double mu = 0.0, mv = 0.0;
#pragma unroll
              for (int l = 0; l < 7; ++l) { mu = fma(tab.m2[l], kv[j][l], mu); mv = fma(tab.c_mid[l], kv[j][l], mv); }
              if (is_free[j]) {
                out[fidx[j]] = interp_eval(y0[j], y1[j], y0[j] + dt * (tab.sum_mid * y0[3 + j] + dt * mu), dt * y0[3 + j], dt * y1[3 + j], x);
                out[nf + fidx[j]] = interp_eval(y0[3 + j], y1[3 + j], y0[3 + j] + dt * mv, dt * kv[j][0], dt * kv[j][6], x);
              }
            }
            ++it;
            istep = 0;
          }
#pragma unroll
          for (int j = 0; j < 3; ++j) { tp.st(F_U0 + j, y1[j]); tp.st(F_V0 + j, y1[3 + j]); tp.st(F_KV + j, kv[j][6]); }
          tp.fence_st();
          t = t_new;
          ++n_acc;
          if (it >= a.n_t) running = false;
        }
        const double dfactor = ratio < 1.0 ? 1.0 : 0.2;
        const double factor = fmin(10.0, fmax(inv_fifth_root(ratio) * 0.9, dfactor));
        dt = (ratio == 0.0) ? dt * 10.0 : dt * factor;
        if (running) {
          if (!(dt > 0.0)) { status |= DFX_STATUS_DT_UNDERFLOW; running = false; }
          else if (istep >= a.max_steps) { status |= DFX_STATUS_MAX_STEPS; running = false; }
        }
        ev = 0;
      }
    }
  }
  if (status != 0) {  // integration stopped early: the remaining outputs are not defined
    for (; it < a.n_t; ++it)
      for (int j = 0; j < 3; ++j) if (is_free[j]) { ys[(long long)it * 2 * nf + fidx[j]] = nan(""); ys[(long long)it * 2 * nf + nf + fidx[j]] = nan(""); }
  }
  if (tid == 0 && a.stats) {
    DfxStats st;
    st.steps = n_steps; st.accepted = n_acc; st.rhs_evals = n_rhs; st.status = status; st.reserved = 0; st.last_dt = dt;
    a.stats[design] = st;
  }
  if (NT > 0) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_sh), "r"(A.tmem_cols_alloc));
  }
}

}  // namespace dfx
