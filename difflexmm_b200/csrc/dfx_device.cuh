// dfx_device.cuh -- device-side building blocks shared by the forward and adjoint kernels:
// dual numbers, the analytic per-bond energy gradient (reference energy.py:120-176, :70-117,
// :204-219, :333-361 written out in closed form), drive / load signals
//  and the Dormand-Prince tableau of jax.experimental.ode.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/dfx.h"

namespace dfx {

constexpr double kPi = 3.14159265358979323846;

// ------------------------------------------------------------------------------------------
// Dormand-Prince 5(4) as used by jax.experimental.ode (runge_kutta_step / interp_fit_dopri).
// Because du/dt = v, the displacement stages are expressed through the stored velocity
// derivatives only:  u_s = u0 + h*ALPHA[s]*v0 + h^2 * sum_l A2[s][l]*kv_l   (A2 = beta*beta),
// which removes the seven displacement k-vectors from on-chip storage.
// ------------------------------------------------------------------------------------------
struct Tableau {
  double alpha[6];
  double beta[6][6];
  double c_sol[7], c_err[7], c_mid[7];
  double a2[6][6];                    // u-stage coefficients on kv_l (h^2)
  double s2[7], e2[7], m2[7];         // same for solution / error / midpoint combos
  double sum_sol, sum_err, sum_mid;   // sums of c_* (coefficient of h*v0)
};

__host__ inline Tableau make_tableau() {
  Tableau t = {};
  const double alpha[6] = {1 / 5., 3 / 10., 4 / 5., 8 / 9., 1., 1.};
  const double beta[6][6] = {
      {1 / 5., 0, 0, 0, 0, 0},
      {3 / 40., 9 / 40., 0, 0, 0, 0},
      {44 / 45., -56 / 15., 32 / 9., 0, 0, 0},
      {19372 / 6561., -25360 / 2187., 64448 / 6561., -212 / 729., 0, 0},
      {9017 / 3168., -355 / 33., 46732 / 5247., 49 / 176., -5103 / 18656., 0},
      {35 / 384., 0, 500 / 1113., 125 / 192., -2187 / 6784., 11 / 84.}};
  const double c_sol[7] = {35 / 384., 0, 500 / 1113., 125 / 192., -2187 / 6784., 11 / 84., 0};
  const double c_err[7] = {35 / 384. - 1951 / 21600., 0, 500 / 1113. - 22642 / 50085., 125 / 192. - 451 / 720.,
                           -2187 / 6784. - -12231 / 42400., 11 / 84. - 649 / 6300., -1. / 60.};
  const double c_mid[7] = {6025192743. / 30085553152. / 2, 0, 51252292925. / 65400821598. / 2,
                           -2691868925. / 45128329728. / 2, 187940372067. / 1594534317056. / 2,
                           -1776094331. / 19743644256. / 2, 11237099. / 235043384. / 2};
  for (int i = 0; i < 6; ++i) {
    t.alpha[i] = alpha[i];
    for (int j = 0; j < 6; ++j) t.beta[i][j] = beta[i][j];
  }
  for (int j = 0; j < 7; ++j) { t.c_sol[j] = c_sol[j]; t.c_err[j] = c_err[j]; t.c_mid[j] = c_mid[j]; }
  // ku_1 = v0 ; ku_j = v0 + h * sum_l beta[j-2][l] kv_l   (j = 2..7, 1-based k index)
  // => sum_j w_j ku_j = (sum_j w_j) v0 + h * sum_l (sum_{j>l} w_j beta[j-2][l-1]) kv_l
  auto fold = [&](const double* w, int nw, double* out, double& wsum) {
    long double s = 0;
    for (int j = 0; j < nw; ++j) s += w[j];
    wsum = (double)s;
    for (int l = 0; l < 7; ++l) {
      long double acc = 0;
      for (int j = l + 1; j < nw; ++j) acc += (long double)w[j] * (long double)(l < 6 ? beta[j - 1][l] : 0.0);
      out[l] = (double)acc;
    }
  };
  for (int s = 0; s < 6; ++s) {
    double w[7] = {0, 0, 0, 0, 0, 0, 0}, out[7], ws;
    for (int j = 0; j <= s; ++j) w[j] = beta[s][j];
    fold(w, s + 1, out, ws);
    for (int l = 0; l < 6; ++l) t.a2[s][l] = out[l];
  }
  fold(c_sol, 7, t.s2, t.sum_sol);
  fold(c_err, 7, t.e2, t.sum_err);
  fold(c_mid, 7, t.m2, t.sum_mid);
  return t;
}

// ------------------------------------------------------------------------------------------
// dual numbers: value + one directional derivative (forward mode over the analytic gradient)
// ------------------------------------------------------------------------------------------
struct Dual {
  double v, d;
  __device__ __forceinline__ Dual() {}
  __device__ __forceinline__ Dual(double v_) : v(v_), d(0.0) {}
  __device__ __forceinline__ Dual(double v_, double d_) : v(v_), d(d_) {}
};
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
__device__ __forceinline__ Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, fma(a.v, b.d, a.d * b.v)); }
__device__ __forceinline__ Dual operator*(Dual a, double b) { return Dual(a.v * b, a.d * b); }
__device__ __forceinline__ Dual operator*(double b, Dual a) { return Dual(a.v * b, a.d * b); }
__device__ __forceinline__ Dual operator+(Dual a, double b) { return Dual(a.v + b, a.d); }
__device__ __forceinline__ Dual operator-(Dual a, double b) { return Dual(a.v - b, a.d); }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
  double inv = 1.0 / b.v;
  double q = a.v * inv;
  return Dual(q, (a.d - q * b.d) * inv);
}
__device__ __forceinline__ double val(double a) { return a; }
__device__ __forceinline__ double val(Dual a) { return a.v; }
__device__ __forceinline__ double dot_part(double) { return 0.0; }
__device__ __forceinline__ double dot_part(Dual a) { return a.d; }

// L = sqrt(L2), 1/L2 from r = 1/sqrt(value of L2)
__device__ __forceinline__ double len_from(double L2, double r) { return L2 * r; }
__device__ __forceinline__ Dual len_from(Dual L2, double r) { return Dual(L2.v * r, 0.5 * L2.d * r); }
__device__ __forceinline__ double inv_from(double, double r) { return r * r; }
__device__ __forceinline__ Dual inv_from(Dual L2, double r) { const double i = r * r; return Dual(i, -L2.d * i * i); }
template <class T> __device__ __forceinline__ T make_T(double v, double d);
template <> __device__ __forceinline__ double make_T<double>(double v, double) { return v; }
template <> __device__ __forceinline__ Dual make_T<Dual>(double v, double d) { return Dual(v, d); }

// reciprocal and reciprocal square root helpers
__device__ __forceinline__ double recipT(double a) { return 1.0 / a; }
__device__ __forceinline__ Dual recipT(Dual a) {
  double inv = 1.0 / a.v;
  return Dual(inv, -a.d * inv * inv);
}
__device__ __forceinline__ double sqrtT(double a) { return sqrt(a); }
__device__ __forceinline__ Dual sqrtT(Dual a) {
  double s = sqrt(a.v);
  return Dual(s, 0.5 * a.d / s);
}
// atan2 with the derivative expressed through a supplied 1/(x^2+y^2)
__device__ __forceinline__ double atan2T(double y, double x, double) { return atan2(y, x); }
__device__ __forceinline__ Dual atan2T(Dual y, Dual x, Dual inv_r2) {
  return Dual(atan2(y.v, x.v), (x.v * y.d - y.v * x.d) * inv_r2.v);
}
// jnp.mod(a + pi, 2 pi) - pi on the value; derivative 1
__device__ __forceinline__ double wrap_value(double a) {
  const double two_pi = 2.0 * kPi, inv_two_pi = 1.0 / two_pi;
  double m = a + kPi;
  m = fma(-two_pi, floor(m * inv_two_pi), m);
  return m - kPi;
}
__device__ __forceinline__ double wrapT(double a) { return wrap_value(a); }
__device__ __forceinline__ Dual wrapT(Dual a) { return Dual(wrap_value(a.v), a.d); }

// ------------------------------------------------------------------------------------------
// per-bond gradient
// ------------------------------------------------------------------------------------------
template <class T>
struct BlockState {  // one rigid unit at the current stage: displacement, rotation, sin/cos(theta)
  T x, y, th, s, c;
};
__device__ __forceinline__ void make_block(double x, double y, double th, double sv, double cv, BlockState<double>& b) {
  b.x = x; b.y = y; b.th = th; b.s = sv; b.c = cv;
}

__device__ __forceinline__ void make_block(double x, double y, double th, double sv, double cv, double wx, double wy, double wth,
                                           BlockState<Dual>& b) {
  b.x = Dual(x, wx); b.y = Dual(y, wy); b.th = Dual(th, wth);
  b.s = Dual(sv, cv * wth); b.c = Dual(cv, -sv * wth);
}

struct BondConst {  // per design and bond, precomputed once per launch: reference vector, its length and 1/length
  double r0x, r0y, L0, iL0;
};

// atan(t) for |t| <= tan(pi/8): t * P(t^2), interpolated at Chebyshev nodes, max error 6e-17
__device__ __forceinline__ double atan_small(double t) {
  const double u = t * t;
  double p = -0.01750805688110568;
  p = fma(p, u, 0.03769427981208796);
  p = fma(p, u, -0.0502446052811003);
  p = fma(p, u, 0.05844521903264865);
  p = fma(p, u, -0.06662628649373917);
  p = fma(p, u, 0.07692016940630392);
  p = fma(p, u, -0.0909089521235964);
  p = fma(p, u, 0.11111110689133682);
  p = fma(p, u, -0.14285714278117734);
  p = fma(p, u, 0.19999999999929152);
  p = fma(p, u, -0.33333333333333076);
  p = fma(p, u, 1.0);
  return t * p;
}
// angle in (-pi, pi] of the unit vector (cg, sg); the usual case |angle| <= pi/4 is branch free
__device__ __forceinline__ double angle_of_unit(double sg, double cg) {
  if (cg >= 0.70710678118654757) {
    const double d = 1.0 + cg;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    r = fma(fma(-d, r, 1.0), r, r);
    r = fma(fma(-d, r, 1.0), r, r);
    return 2.0 * atan_small(sg * r);  // tan(angle/2) = sin/(1+cos)
  }
  return atan2(sg, cg);
}
// sin and cos of a moderate angle (block rotations): quadrant reduction with a two-term pi/2 (the products are exact
// inside the fma) and the fdlibm kernel polynomials on [-pi/4, pi/4] (< 1 ulp, checked by dfx_math_selftest); the
// library's sincos spends ~250 instructions on its general range reduction.  |x| >= 1e5 or NaN: library call.
__device__ __forceinline__ void sincos_fast(double x, double* sn, double* cs) {
  if (!(fabs(x) < 1.0e5)) { sincos(x, sn, cs); return; }
  const double q = rint(x * 0.6366197723675814);
  double r = fma(-q, 1.5707963267948966, x);
  r = fma(-q, 6.123233995736766e-17, r);
  const double z = r * r;
  double ps = 1.58969099521155010221e-10;
  ps = fma(ps, z, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  double pc = -1.13596475577881948265e-11;
  pc = fma(pc, z, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  const double s = fma(r * z, ps, r);
  const double c = fma(z * z, pc, fma(-0.5, z, 1.0));
  const int n = (int)q;
  const double ss = (n & 1) ? c : s, cc = (n & 1) ? s : c;
  *sn = (n & 2) ? -ss : ss;
  *cs = ((n + 1) & 2) ? -cc : cc;
}
// ratio^(-1/5) for the step-size controller: float estimate + two Newton steps on y^-5 = r (y <- y (6 - r y^5) / 5), ~1 ulp
__device__ __forceinline__ double inv_fifth_root(double r) {
  if (!(r > 1e-30)) return 1e6;  // the controller clamps the factor to 10
  double y = (double)__powf((float)r, -0.2f);
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const double y2 = y * y, y5 = y2 * y2 * y;
    y = y * fma(-r, y5, 6.0) * 0.2;
  }
  return y;
}

// 1/sqrt(x) for positive normal x: hardware approximation + two Newton steps
__device__ __forceinline__ double rsqrt_pos(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x * r, r, 1.0);
  r = fma(fma(0.375, e, 0.5) * e, r, r);
  e = fma(-x * r, r, 1.0);
  r = fma(0.5 * e, r, r);
  return r;
}

template <class T>
struct BondOut {
  T f1[3], f2[3];        // dE/d(x,y,theta) of block 1 and block 2
  T gr1[2], gr2[2];      // dE/d(centroid_node_vector) of node 1 / node 2
  T gr0[2];              // dE/d(reference_vector)
  T gks, gksh, gkr;      // dE/d(k_stretch, k_shear, k_rot)
};

// contact energy of one void angle psi (reference energy.py:333-361) -- derivative w.r.t. psi and,
// optionally, w.r.t. (min_angle, cutoff_angle, k_contact).  jnp.where semantics: inactive => 0.
template <class T>
__device__ __forceinline__ bool contact_term(T psi, double tmin, double tcut, double kc, T& dpsi, T& dmin, T& dcut, T& dkc) {
  const double pv = val(psi);
  if (!(pv < tmin) && pv < tcut) {
    const double w = tcut - tmin;
    T x = (psi - tcut) * (1.0 / w);
    T ip = recipT(x + 1.0), im = recipT(x - 1.0);
    T h = ip - im - 2.0;
    T hp = im * im - ip * ip;
    dpsi = hp * (kc * 0.25 * w);
    dkc = h * (w * w * 0.25);
    dcut = (h * (2.0 * w) - hp * (x + 1.0) * w) * (kc * 0.25);
    dmin = (hp * x * w - h * (2.0 * w)) * (kc * 0.25);
    return true;
  }
  dpsi = T(0.0); dmin = T(0.0); dcut = T(0.0); dkc = T(0.0);
  return false;
}

// ---- distance-based contact between the void edges of a bond (reference energy.py:222-330, build_contact_energy with
// angle_based=False at :364-407).  The six polygon vertices involved are
//     P[k] = block centroid + block displacement + R(theta) * centroid_node_vector,  k = node, next, previous of block 1, then of block 2,
// the two gaps are edges_distance((P0,P1),(P3,P5)) and edges_distance((P0,P2),(P3,P4)), each the smallest of four
// point-to-segment distances (branches decided on values: jnp.where / jnp.min), and each gap enters the same 1/x-type
// energy as a void angle (contact_term; min / cutoff are lengths here).  Output: dE/d(block DOFs) (whose x, y entries
// are also dE/d(block centroid)), dE/d(centroid_node_vector) of the six vertices and dE/d(contact parameters).
template <class T>
struct DistanceContactOut {
  T f1[3], f2[3];
  T gr[6][2];
  T gmin, gcut, gkc;
};

__device__ __forceinline__ double point_segment_value(double px, double py, double x0, double y0, double x1, double y1) {
  const double ex = x1 - x0, ey = y1 - y0, rx = px - x0, ry = py - y0;
  const double t = (rx * ex + ry * ey) / (ex * ex + ey * ey);
  if (t >= 0.0 && t <= 1.0) return sqrt((rx * rx - (t * ex) * (t * ex)) + (ry * ry - (t * ey) * (t * ey)));
  if (t < 0.0) return sqrt(rx * rx + ry * ry);
  const double sx = px - x1, sy = py - y1;
  return sqrt(sx * sx + sy * sy);
}

// d = distance of P[ip] to the segment (P[i0], P[i1]); adds scale(d) * dd/dP to the vertex gradients G.  On the segment
// d^2 = |r|^2 - (r.e)^2 / |e|^2 with foot parameter t: dd/dp = n, dd/dx1 = -t n, dd/dx0 = (t - 1) n, n = (r - t e) / d.
template <class T, class F>
__device__ __forceinline__ void point_segment_accumulate(const T (*P)[2], int ip, int i0, int i1, T (*G)[2], F&& scale) {
  const T ex = P[i1][0] - P[i0][0], ey = P[i1][1] - P[i0][1], rx = P[ip][0] - P[i0][0], ry = P[ip][1] - P[i0][1];
  const T t = (rx * ex + ry * ey) * recipT(ex * ex + ey * ey);
  const double tv = val(t);
  if (tv >= 0.0 && tv <= 1.0) {
    const T tex = t * ex, tey = t * ey;
    const T d = sqrtT((rx * rx - tex * tex) + (ry * ry - tey * tey));
    const T w = scale(d) * recipT(d);
    const T nx = (rx - tex) * w, ny = (ry - tey) * w;
    G[ip][0] = G[ip][0] + nx; G[ip][1] = G[ip][1] + ny;
    G[i1][0] = G[i1][0] - t * nx; G[i1][1] = G[i1][1] - t * ny;
    const T tm = t - 1.0;
    G[i0][0] = G[i0][0] + tm * nx; G[i0][1] = G[i0][1] + tm * ny;
  } else {
    const int iq = tv < 0.0 ? i0 : i1;  // nearest end of the segment
    const T sx = P[ip][0] - P[iq][0], sy = P[ip][1] - P[iq][1];
    const T d = sqrtT(sx * sx + sy * sy);
    const T w = scale(d) * recipT(d);
    G[ip][0] = G[ip][0] + sx * w; G[ip][1] = G[ip][1] + sy * w;
    G[iq][0] = G[iq][0] - sx * w; G[iq][1] = G[iq][1] - sy * w;
  }
}

// cen: block centroids of block 1, 2; r[k]: centroid_node_vectors of the six vertices
template <class T>
__device__ __noinline__ void distance_contact(const BlockState<T>& b1, const BlockState<T>& b2, const double* cen1, const double* cen2,
                                              const double (*r)[2], double dmin, double dcut, double kc, DistanceContactOut<T>& o) {
  T P[6][2], G[6][2];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const BlockState<T>& b = k < 3 ? b1 : b2;
    const double* cen = k < 3 ? cen1 : cen2;
    P[k][0] = b.x + (b.c * r[k][0] - b.s * r[k][1]) + cen[0];
    P[k][1] = b.y + (b.s * r[k][0] + b.c * r[k][1]) + cen[1];
    G[k][0] = make_T<T>(0.0, 0.0); G[k][1] = make_T<T>(0.0, 0.0);
  }
  o.gmin = make_T<T>(0.0, 0.0); o.gcut = make_T<T>(0.0, 0.0); o.gkc = make_T<T>(0.0, 0.0);
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const int a0 = 0, a1 = side == 0 ? 1 : 2, c0 = 3, c1 = side == 0 ? 5 : 4;
    // candidates in the reference's order: ends of the second edge onto the first, ends of the first onto the second
    const int cand[4][3] = {{c0, a0, a1}, {c1, a0, a1}, {a0, c0, c1}, {a1, c0, c1}};
    int best = 0;
    double dbest = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double dv = point_segment_value(val(P[cand[k][0]][0]), val(P[cand[k][0]][1]), val(P[cand[k][1]][0]), val(P[cand[k][1]][1]),
                                            val(P[cand[k][2]][0]), val(P[cand[k][2]][1]));
      if (k == 0 || dv < dbest) { dbest = dv; best = k; }
    }
    point_segment_accumulate<T>(P, cand[best][0], cand[best][1], cand[best][2], G, [&](T d) {
      T e, m, u, kk;
      contact_term<T>(d, dmin, dcut, kc, e, m, u, kk);
      o.gmin = o.gmin + m; o.gcut = o.gcut + u; o.gkc = o.gkc + kk;
      return e;
    });
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) { o.f1[j] = make_T<T>(0.0, 0.0); o.f2[j] = make_T<T>(0.0, 0.0); }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const BlockState<T>& b = k < 3 ? b1 : b2;
    T* f = k < 3 ? o.f1 : o.f2;
    f[0] = f[0] + G[k][0];
    f[1] = f[1] + G[k][1];
    // d(R r)/d theta = (-s r_x - c r_y, c r_x - s r_y)
    f[2] = f[2] + G[k][1] * (b.c * r[k][0] - b.s * r[k][1]) - G[k][0] * (b.s * r[k][0] + b.c * r[k][1]);
    o.gr[k][0] = b.c * G[k][0] + b.s * G[k][1];
    o.gr[k][1] = b.c * G[k][1] - b.s * G[k][0];
  }
}

// EXT: also compile the bond kinds that only the generic kernels support (DFX_BOND_SPRING)
template <class T, bool PARAMS, bool EXT = false>
__device__ __forceinline__ void bond_gradient(int energy_kind, const BlockState<T>& b1, const BlockState<T>& b2,
                                              double r1x, double r1y, double r2x, double r2y, const BondConst& bc,
                                              double ks, double ksh, double kr, BondOut<T>& o) {
  // node displacements u = u_c + (R(theta) - I) r   (kinematics.py:24-31)
  T c1m = b1.c - 1.0, c2m = b2.c - 1.0;
  T n1x = b1.x + c1m * r1x - b1.s * r1y;
  T n1y = b1.y + b1.s * r1x + c1m * r1y;
  T n2x = b2.x + c2m * r2x - b2.s * r2y;
  T n2y = b2.y + b2.s * r2x + c2m * r2y;
  T dUx = n2x - n1x, dUy = n2y - n1y;
  // d(node)/d(theta) = R'(theta) r
  T t1x = -(b1.s * r1x) - b1.c * r1y, t1y = b1.c * r1x - b1.s * r1y;
  T t2x = -(b2.s * r2x) - b2.c * r2y, t2y = b2.c * r2x - b2.s * r2y;
  T dth = b2.th - b1.th;
  T mean = (b2.th + b1.th) * 0.5;
  const double L0 = bc.L0, L0sq = L0 * L0;
  T gdx, gdy, tq;  // dE/d(dU) and dE/d(mean rotation)
  if (energy_kind == DFX_BOND_LIGAMENT) {
    T dx = dUx + bc.r0x, dy = dUy + bc.r0y;
    T L2 = dx * dx + dy * dy;
    const double rinv = rsqrt_pos(val(L2));
    T iL2 = inv_from(L2, rinv);
    T L = len_from(L2, rinv);
    // shear angle (reference energy.py:139-153): wrap(atan2(current) - atan2(R(mean) r0)) = wrap((phi - phi0) - mean).
    // phi - phi0 is the angle of the unit bond direction rotated by the constant -phi0 of the bond: no atan2
    // range reduction in the usual case, and the mean rotation enters as a plain number exactly as in the reference.
    double gv;
    {
      const double cp = bc.r0x * bc.iL0, sp = bc.r0y * bc.iL0;
      const double ex = val(dx) * rinv, ey = val(dy) * rinv;
      gv = wrap_value(angle_of_unit(cp * ey - sp * ex, cp * ex + sp * ey) - val(mean));
    }
    T gam = make_T<T>(gv, (val(dx) * dot_part(dy) - val(dy) * dot_part(dx)) * val(iL2) - dot_part(mean));
    T ext = L - L0;
    T A = ext * ks * L * iL2;  // ks (L-L0)/L
    T M = gam * (ksh * L0sq);  // dE/dgamma
    T Bc = M * iL2;
    gdx = A * dx - Bc * dy;
    gdy = A * dy + Bc * dx;
    tq = -M;
    if (PARAMS) {
      T dE_dL0 = gam * gam * (ksh * L0) - ext * ks;
      o.gr0[0] = gdx + dE_dL0 * (bc.r0x / L0) + M * (bc.r0y / L0sq);
      o.gr0[1] = gdy + dE_dL0 * (bc.r0y / L0) - M * (bc.r0x / L0sq);
      o.gks = ext * ext * 0.5;
      o.gksh = gam * gam * (0.5 * L0sq);
    }
  } else if (EXT && energy_kind == DFX_BOND_SPRING) {
    // zero-length spring (reference energy.py:49-66): E = ks |dU|^2 / 2 (+ kr dtheta^2 / 2 below)
    gdx = dUx * ks;
    gdy = dUy * ks;
    tq = make_T<T>(0.0, 0.0);
    if (PARAMS) {
      o.gr0[0] = make_T<T>(0.0, 0.0);
      o.gr0[1] = make_T<T>(0.0, 0.0);
      o.gks = (dUx * dUx + dUy * dUy) * 0.5;
      o.gksh = make_T<T>(0.0, 0.0);
    }
  } else {
    const double iL0 = 1.0 / L0;
    T dotp = dUx * bc.r0x + dUy * bc.r0y;
    T crs = dUy * bc.r0x - dUx * bc.r0y;
    T ea = dotp * iL0;
    T es = crs * iL0 - mean * L0;
    T a = ea * ks, s = es * ksh;
    gdx = a * (bc.r0x * iL0) - s * (bc.r0y * iL0);
    gdy = a * (bc.r0y * iL0) + s * (bc.r0x * iL0);
    tq = -(s * L0);
    if (PARAMS) {
      const double iL03 = iL0 * iL0 * iL0;
      T dea0 = dUx * iL0 - dotp * (bc.r0x * iL03), dea1 = dUy * iL0 - dotp * (bc.r0y * iL03);
      T des0 = dUy * iL0 - crs * (bc.r0x * iL03) - mean * (bc.r0x * iL0);
      T des1 = -(dUx * iL0) - crs * (bc.r0y * iL03) - mean * (bc.r0y * iL0);
      o.gr0[0] = a * dea0 + s * des0;
      o.gr0[1] = a * dea1 + s * des1;
      o.gks = ea * ea * 0.5;
      o.gksh = es * es * 0.5;
    }
  }
  T bend = dth * kr;
  o.f1[0] = -gdx; o.f1[1] = -gdy;
  o.f1[2] = tq * 0.5 - (gdx * t1x + gdy * t1y) - bend;
  o.f2[0] = gdx; o.f2[1] = gdy;
  o.f2[2] = tq * 0.5 + (gdx * t2x + gdy * t2y) + bend;
  if (PARAMS) {
    o.gkr = dth * dth * 0.5;
    o.gr1[0] = -(c1m * gdx + b1.s * gdy);
    o.gr1[1] = b1.s * gdx - c1m * gdy;
    o.gr2[0] = c2m * gdx + b2.s * gdy;
    o.gr2[1] = c2m * gdy - b2.s * gdx;
  }
}

// ------------------------------------------------------------------------------------------
// drive / load signals
// ------------------------------------------------------------------------------------------
struct DriveEval {
  double s[2], sdot[2], dsdp[2][DFX_MAX_DRIVE_PARAMS];
};
struct DriveTable {  // tabulated drive (jnp.interp): device arrays
  const double* t;
  const double* v;
  int n;
};

__device__ __forceinline__ void pulse_eval(double tau, double A, double f, bool windowed, bool want_grad, double& s,
                                           double& ds_dtau, double& ds_dA, double& ds_df) {
  const bool on = (tau > 0.0) && (!windowed || tau < 1.0 / f);
  s = ds_dtau = ds_dA = ds_df = 0.0;
  if (on) {
    const double ph = 2.0 * kPi * f * tau;
    double sn, cs;
    sincos_fast(ph, &sn, &cs);
    const double shape = (1.0 - cs) * 0.5;
    s = A * shape;
    if (want_grad) {
      ds_dA = shape;
      ds_dtau = A * kPi * f * sn;
      ds_df = A * kPi * tau * sn;
    }
  }
}

__device__ inline void drive_eval(int kind, double t, const double* p, bool want_grad, DriveEval& e, DriveTable tab = DriveTable{nullptr, nullptr, 0}) {
  e.s[0] = e.s[1] = e.sdot[0] = e.sdot[1] = 0.0;
#pragma unroll
  for (int j = 0; j < DFX_MAX_DRIVE_PARAMS; ++j) e.dsdp[0][j] = e.dsdp[1][j] = 0.0;
  switch (kind) {
    case DFX_DRIVE_PULSE:
    case DFX_DRIVE_HARMONIC: {
      double s, dtau, dA, df;
      pulse_eval(t - p[2], p[0], p[1], kind == DFX_DRIVE_PULSE, want_grad, s, dtau, dA, df);
      e.s[0] = s; e.sdot[0] = dtau;
      e.dsdp[0][0] = dA; e.dsdp[0][1] = df; e.dsdp[0][2] = -dtau;
    } break;
    case DFX_DRIVE_RAMP: {
      if (t < 1.0 / p[1]) { e.s[0] = p[0] * t * p[1]; e.sdot[0] = p[0] * p[1]; e.dsdp[0][0] = t * p[1]; e.dsdp[0][1] = p[0] * t; }
      else { e.s[0] = p[0]; e.dsdp[0][0] = 1.0; }
    } break;
    case DFX_DRIVE_STATIC_PULSE: {
      const double A = p[0], f = p[1], cs = p[2], csr = p[3], delay = p[4];
      double s, dtau, dA, df;
      pulse_eval(t - cs / csr - delay, A, f, true, want_grad, s, dtau, dA, df);
      e.s[0] = s; e.sdot[0] = dtau;
      e.dsdp[0][0] = dA; e.dsdp[0][1] = df;
      e.dsdp[0][2] = -dtau / csr; e.dsdp[0][3] = dtau * cs / (csr * csr); e.dsdp[0][4] = -dtau;
      if (t < cs / csr) { e.s[1] = t * csr; e.sdot[1] = csr; e.dsdp[1][3] = t; }
      else { e.s[1] = cs; e.dsdp[1][2] = 1.0; }
    } break;
    case DFX_DRIVE_TABLE: {
      // jnp.interp(t, xp, fp): i = clip(searchsorted(xp, t, side='right'), 1, n-1); linear on [xp[i-1], xp[i]];
      // constant (and zero slope) outside the table
      if (tab.n <= 0) break;
      if (tab.n == 1 || t < tab.t[0]) { e.s[0] = tab.v[0]; break; }
      if (t > tab.t[tab.n - 1]) { e.s[0] = tab.v[tab.n - 1]; break; }
      int lo = 0, hi = tab.n;  // first index with xp[idx] > t
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (tab.t[mid] <= t) lo = mid + 1; else hi = mid; }
      int i = lo < 1 ? 1 : (lo > tab.n - 1 ? tab.n - 1 : lo);
      const double dx = tab.t[i] - tab.t[i - 1], df = tab.v[i] - tab.v[i - 1];
      if (dx == 0.0) { e.s[0] = tab.v[i]; break; }
      e.s[0] = tab.v[i - 1] + (t - tab.t[i - 1]) / dx * df;
      e.sdot[0] = df / dx;
    } break;
    default: break;
  }
}

__device__ inline void load_eval(int kind, double t, const double* c, double& s, double& sdot) {
  s = sdot = 0.0;
  if (kind == DFX_LOAD_RAMP) {
    if (t < 1.0 / c[1]) { s = c[0] * t * c[1]; sdot = c[0] * c[1]; } else { s = c[0]; }
  } else if (kind == DFX_LOAD_SECH2) {
    const double a = c[0], sh = c[1];
    const double x = t / sh - 3.0;
    const double ch = cosh(x), th = tanh(-x);
    const double pre = 2.0 * a / (sh * sh);
    s = pre / (ch * ch) * th;
    const double dch2 = -2.0 * sinh(x) / (ch * ch * ch);
    const double dth = -(1.0 - th * th);
    sdot = pre * (dch2 * th + dth / (ch * ch)) / sh;
  }
}

// ------------------------------------------------------------------------------------------
// CTA-wide sum (warp shuffles + one shared array), result broadcast to every thread
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// red must hold >= 33 doubles.  Contains two __syncthreads().
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double t = lane < nw ? red[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// ------------------------------------------------------------------------------------------
// thread-block clusters: one lattice spread over the CTAs of a cluster (large lattices, small batches).
// barrier.cluster arrive.release / wait.acquire orders global-memory traffic between the CTAs
// (ptxas adds the L1 invalidate), so stage arrays exchanged through the L2-resident scratch are coherent.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_nctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr int kMaxCluster = 16;   // CTAs of a thread-block cluster (non-portable maximum)
constexpr int kMaxGroup = 160;    // CTAs cooperating on one design in either multi-CTA mode (>= SM count)

// Multi-CTA mode of the generic kernels.  MODE 1: thread-block cluster (hardware barrier, <= 16 CTAs).
// MODE 2: a group of co-resident CTAs of a cooperative launch with a software barrier (monotonic arrival counter in
// global memory; thread 0 of every CTA arrives, spins and fences, the rest wait at the CTA barrier -- the structure of
// cooperative_groups' grid sync, but per group of CTAs so that several designs can share the GPU).
struct GroupCtx {
  int mode, rank, ncta;
  unsigned long long* counter;  // MODE 2: arrival counter of this group (zeroed before the launch)
  unsigned long long epoch;     // number of barriers passed
  double* cred;                 // [2][kMaxGroup] partial sums
  int parity;
};

__device__ __forceinline__ void group_sync(GroupCtx& g) {
  if (g.mode == 1) { cluster_sync_all(); return; }
  __syncthreads();
  if (threadIdx.x == 0) {
    // arrive with release semantics (the CTA's writes, ordered before this by the CTA barrier, become visible first),
    // then poll with acquire loads (which also drop this SM's stale L1 lines for the threads released below)
    const unsigned long long target = (g.epoch + 1) * (unsigned long long)g.ncta;
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(g.counter), "l"(1ULL) : "memory");
    unsigned long long seen, spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(g.counter) : "memory");
      if (++spins > (1ULL << 28)) __trap();  // a CTA of the group never arrived: fail the launch instead of hanging
    } while (seen < target);
  }
  g.epoch++;
  __syncthreads();
}

// Sum over every thread of the group; identical (bitwise) on all CTAs, so control flow stays uniform.
__device__ __forceinline__ double group_sum(double v, double* red, GroupCtx& g) {
  const double part = block_sum(v, red);
  if (threadIdx.x == 0) g.cred[g.parity * kMaxGroup + g.rank] = part;
  group_sync(g);
  double s = 0.0;
  for (int r = threadIdx.x & 31; r < g.ncta; r += 32) s += g.cred[g.parity * kMaxGroup + r];
  g.parity ^= 1;
  return warp_sum(s);
}

}  // namespace dfx
