// dfx_geometry.cuh -- design -> solver parameters on the device, and the vector-Jacobian product back
// (the step immediately before the solver).
//
// Reference: QuadGeometry / KagomeGeometry `reference_node_vectors`, `centroid_node_vectors`, `block_centroids`
// (geometry.py:607-952) and compute_inertia / polygon_area / polygon_centroid / polygon_polar_moment
// (geometry.py:71-160).  Every polygon vertex is `base_node + design[node_design]` (one design 2-vector per vertex,
// or none); a block's outputs depend on its own vertices only:
//     c      = polygon centroid (shoelace, |area|)        -> block_centroids = reference point + c
//     cnv_l  = vertex_l - c
//     inertia = rho * (A, A, J),  J = polar moment about the centroid
// Forward: one thread per (design, block).  VJP: forward-mode duals over the 2*n_npb vertex coordinates of the block
// (the function is tiny), then a gather over the design->vertex table: no atomics.
#pragma once

#include "dfx_device.cuh"

namespace dfx {

constexpr int kMaxPolygon = 8;

struct DevGeometry {
  int n_blocks, n_npb, n_nodes, n_design;
  const double* base_nodes;  // [n_nodes][2]
  const int* node_design;    // [n_nodes] design 2-vector index or -1
  const int* design_off;     // [n_design + 1] CSR offsets into design_nodes
  const int* design_nodes;   // vertices fed by each design 2-vector
};

__device__ __forceinline__ double absT(double a) { return fabs(a); }
__device__ __forceinline__ Dual absT(Dual a) { return a.v < 0.0 ? Dual(-a.v, -a.d) : a; }

// polygon of npb vertices -> centroid (cx, cy), area A, polar moment J about the centroid
template <class T>
__device__ __forceinline__ void polygon_props(const T* vx, const T* vy, int npb, T& cx, T& cy, T& A, T& J) {
  T sc = T(0.0), nx = T(0.0), ny = T(0.0);
  for (int l = 0; l < npb; ++l) {
    const int p = l == 0 ? npb - 1 : l - 1;
    const T cr = vx[p] * vy[l] - vy[p] * vx[l];
    sc = sc + cr;
    nx = nx + (vx[p] + vx[l]) * cr;
    ny = ny + (vy[p] + vy[l]) * cr;
  }
  A = absT(sc * 0.5);
  const T i6A = recipT(A * 6.0);
  cx = nx * i6A;
  cy = ny * i6A;
  T sj = T(0.0);
  for (int l = 0; l < npb; ++l) {
    const int p = l == 0 ? npb - 1 : l - 1;
    const T x1 = vx[p] - cx, y1 = vy[p] - cy, x2 = vx[l] - cx, y2 = vy[l] - cy;
    const T cr = x1 * y2 - y1 * x2;
    const T quad = x1 * x1 + x1 * x2 + x2 * x2 + y1 * y1 + y1 * y2 + y2 * y2;
    sj = sj + cr * quad;
  }
  J = absT(sj * (1.0 / 12.0));
}

__device__ __forceinline__ void load_vertices(const DevGeometry& G, const double* design, int blk, double* vx, double* vy) {
  for (int l = 0; l < G.n_npb; ++l) {
    const int n = blk * G.n_npb + l;
    const int d = G.node_design[n];
    vx[l] = G.base_nodes[2 * n] + (d >= 0 ? design[2 * d] : 0.0);
    vy[l] = G.base_nodes[2 * n + 1] + (d >= 0 ? design[2 * d + 1] : 0.0);
  }
}

// grid (ceil(n_blocks / blockDim), batch)
__global__ void geometry_forward_kernel(DevGeometry G, const double* design, const double* density, long long density_bstride,
                                        double* cnv, double* centroid_shift, double* inertia) {
  const int b = blockIdx.y, blk = blockIdx.x * blockDim.x + threadIdx.x;
  if (blk >= G.n_blocks) return;
  const double* dsg = design + (long long)b * G.n_design * 2;
  double vx[kMaxPolygon], vy[kMaxPolygon];
  load_vertices(G, dsg, blk, vx, vy);
  double cx, cy, A, J;
  polygon_props<double>(vx, vy, G.n_npb, cx, cy, A, J);
  for (int l = 0; l < G.n_npb; ++l) {
    const long long n = (long long)b * G.n_nodes + blk * G.n_npb + l;
    cnv[2 * n] = vx[l] - cx;
    cnv[2 * n + 1] = vy[l] - cy;
  }
  if (centroid_shift) {
    centroid_shift[((long long)b * G.n_blocks + blk) * 2] = cx;
    centroid_shift[((long long)b * G.n_blocks + blk) * 2 + 1] = cy;
  }
  if (inertia) {
    const double rho = density[(long long)b * density_bstride];
    double* o = inertia + ((long long)b * G.n_blocks + blk) * 3;
    o[0] = rho * A; o[1] = rho * A; o[2] = rho * J;
  }
}

// one CTA per design.  node_bar: [batch][n_nodes][2] scratch.
__global__ void geometry_vjp_kernel(DevGeometry G, const double* design, const double* density, long long density_bstride,
                                    const double* cnv_bar, const double* centroid_bar, const double* inertia_bar,
                                    double* node_bar, double* design_bar, double* density_bar) {
  __shared__ double red[40];
  const int b = blockIdx.x;
  const double* dsg = design + (long long)b * G.n_design * 2;
  const double rho = density ? density[(long long)b * density_bstride] : 0.0;
  double* nb = node_bar + (long long)b * G.n_nodes * 2;
  double rho_bar = 0.0;
  for (int blk = threadIdx.x; blk < G.n_blocks; blk += blockDim.x) {
    double vx[kMaxPolygon], vy[kMaxPolygon];
    load_vertices(G, dsg, blk, vx, vy);
    double gx[kMaxPolygon], gy[kMaxPolygon], gc[2] = {0.0, 0.0}, gi[3] = {0.0, 0.0, 0.0};
    for (int l = 0; l < G.n_npb; ++l) {
      const long long n = (long long)b * G.n_nodes + blk * G.n_npb + l;
      gx[l] = cnv_bar ? cnv_bar[2 * n] : 0.0;
      gy[l] = cnv_bar ? cnv_bar[2 * n + 1] : 0.0;
    }
    if (centroid_bar) { gc[0] = centroid_bar[((long long)b * G.n_blocks + blk) * 2]; gc[1] = centroid_bar[((long long)b * G.n_blocks + blk) * 2 + 1]; }
    if (inertia_bar) for (int k = 0; k < 3; ++k) gi[k] = inertia_bar[((long long)b * G.n_blocks + blk) * 3 + k];
    // S = <cnv_bar, cnv> + <centroid_bar, c> + <inertia_bar, rho (A, A, J)>;  dS/d(vertex) by one dual pass per coordinate
    Dual dx[kMaxPolygon], dy[kMaxPolygon];
    for (int l = 0; l < G.n_npb; ++l) { dx[l] = Dual(vx[l]); dy[l] = Dual(vy[l]); }
    for (int k = 0; k < 2 * G.n_npb; ++k) {
      const int l = k >> 1;
      if (k & 1) dy[l].d = 1.0; else dx[l].d = 1.0;
      Dual cx, cy, A, J;
      polygon_props<Dual>(dx, dy, G.n_npb, cx, cy, A, J);
      double s = gc[0] * cx.d + gc[1] * cy.d + rho * ((gi[0] + gi[1]) * A.d + gi[2] * J.d);
      double sg = 0.0;
      for (int m = 0; m < G.n_npb; ++m) sg += gx[m] * cx.d + gy[m] * cy.d;
      s += ((k & 1) ? gy[l] : gx[l]) - sg;  // cnv_m = v_m - c
      nb[2 * (blk * G.n_npb + l) + (k & 1)] = s;
      if (k & 1) dy[l].d = 0.0; else dx[l].d = 0.0;
    }
    if (density_bar) {
      double cx, cy, A, J;
      polygon_props<double>(vx, vy, G.n_npb, cx, cy, A, J);
      rho_bar += (gi[0] + gi[1]) * A + gi[2] * J;
    }
  }
  __syncthreads();  // node_bar of this design was written by this CTA
  for (int d = threadIdx.x; d < G.n_design; d += blockDim.x) {
    double sx = 0.0, sy = 0.0;
    for (int q = G.design_off[d]; q < G.design_off[d + 1]; ++q) {
      const int n = G.design_nodes[q];
      sx += nb[2 * n];
      sy += nb[2 * n + 1];
    }
    design_bar[((long long)b * G.n_design + d) * 2] = sx;
    design_bar[((long long)b * G.n_design + d) * 2 + 1] = sy;
  }
  if (density_bar) {
    const double t = block_sum(rho_bar, red);
    if (threadIdx.x == 0) density_bar[b] = t;
  }
}

}  // namespace dfx

// ---- inequality constraints of the design and their sparse Jacobian -------------------------------
// Reference: OptimizationProblem.setup_angle_constraints / setup_edge_length_constraints
// (problems/quads_focusing.py:473-544) and the jit(jacobian(...)) the nlopt callbacks evaluate (:585-588, :613-616).
// Every row is a function of two polygon edges a = v[p1] - v[p0], b = v[q1] - v[q0] of the undeformed lattice
// (vertex = base + design[node_design]; the polygon centroid cancels in an edge):
//     angle row:  c = -(mod(atan2(a x b, a . b), 2 pi) - min)      (void angles, block angles, boundary block angles)
//     edge row:   c = -(|a| - min)
// so a Jacobian row has at most 4 design 2-vectors: fixed-width rows [4][2] with a shared column table.
namespace dfx {

struct ConstraintRow {
  int p0, p1, q0, q1;  // vertex ids; q0 < 0: edge-length row
  int slot[4];         // Jacobian slot that the gradient w.r.t. (p0, p1, q0, q1) is added to, or -1 (vertex has no design variable)
  double minimum;
};

// grid (ceil(m / blockDim), batch)
__global__ void constraints_kernel(DevGeometry G, const ConstraintRow* rows, int m, const double* design, double* values, double* jac) {
  const int b = blockIdx.y, r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  const ConstraintRow R = rows[r];
  const double* dsg = design + (long long)b * G.n_design * 2;
  auto vertex = [&](int n, double& x, double& y) {
    const int d = G.node_design[n];
    x = G.base_nodes[2 * n] + (d >= 0 ? dsg[2 * d] : 0.0);
    y = G.base_nodes[2 * n + 1] + (d >= 0 ? dsg[2 * d + 1] : 0.0);
  };
  double x0, y0, x1, y1;
  vertex(R.p0, x0, y0);
  vertex(R.p1, x1, y1);
  const double ax = x1 - x0, ay = y1 - y0;
  double gax, gay, gbx = 0.0, gby = 0.0, c;  // dc/da, dc/db
  if (R.q0 < 0) {
    const double len = sqrt(ax * ax + ay * ay);
    c = -(len - R.minimum);
    gax = -ax / len;
    gay = -ay / len;
  } else {
    double u0, v0, u1, v1;
    vertex(R.q0, u0, v0);
    vertex(R.q1, u1, v1);
    const double bx = u1 - u0, by = v1 - v0;
    double phi = atan2(ax * by - ay * bx, ax * bx + ay * by);
    const double two_pi = 6.283185307179586476925286766559;
    phi = phi - two_pi * floor(phi / two_pi);  // jnp.mod(phi, 2 pi)
    c = -(phi - R.minimum);
    const double ia = 1.0 / (ax * ax + ay * ay), ib = 1.0 / (bx * bx + by * by);
    gax = -ay * ia;  // -d phi / d a,  d phi / d a = (a_y, -a_x) / |a|^2
    gay = ax * ia;
    gbx = by * ib;   // -d phi / d b,  d phi / d b = (-b_y, b_x) / |b|^2
    gby = -bx * ib;
  }
  values[(long long)b * m + r] = c;
  if (jac) {
    double out[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
    const double gx[4] = {-gax, gax, -gbx, gbx}, gy[4] = {-gay, gay, -gby, gby};
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int t = R.slot[s];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (t == k) { out[k][0] += gx[s]; out[k][1] += gy[s]; }
    }
    double* o = jac + ((long long)b * m + r) * 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) { o[2 * k] = out[k][0]; o[2 * k + 1] = out[k][1]; }
  }
}

}  // namespace dfx

// ---- RotatedSquareGeometry: one design angle per lattice (reference geometry.py:354-443) --------------------------
// Block k = (i1, i2) of the n1 x n2 grid (i1 fastest) is a square whose half-diagonal direction is rotated by
// parity * angle, parity = (-1)^(i1 + i2):  vertex_l = R(l * pi/2) * half * (1, tan(parity * angle)),  l = 0..3
// (the polygon centroid is the block centre, so these are the centroid_node_vectors), inertia = density * (A, A, J).
namespace dfx {

template <class T>
__device__ __forceinline__ void rotated_square_vertices(T a, double half, T* vx, T* vy) {
  T sn, cs;
  if constexpr (sizeof(T) == sizeof(double)) { double s_, c_; sincos(val(a), &s_, &c_); sn = make_T<T>(s_, 0.0); cs = make_T<T>(c_, 0.0); }
  else { double s_, c_; sincos(val(a), &s_, &c_); sn = make_T<T>(s_, c_ * dot_part(a)); cs = make_T<T>(c_, -s_ * dot_part(a)); }
  const T x = make_T<T>(half, 0.0), y = sn * recipT(cs) * half;
  vx[0] = x; vy[0] = y;          // quarter turns: (x, y) -> (-y, x) -> (-x, -y) -> (y, -x)
  vx[1] = -y; vy[1] = x;
  vx[2] = -x; vy[2] = -y;
  vx[3] = y; vy[3] = -x;
}

// grid (ceil(n_blocks / blockDim), batch)
__global__ void rotated_square_forward_kernel(int n1, int n_blocks, double half, const double* angle, const double* density,
                                              long long density_bstride, double* cnv, double* inertia) {
  const int b = blockIdx.y, k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_blocks) return;
  const double parity = ((k % n1 + k / n1) & 1) ? -1.0 : 1.0;
  double vx[4], vy[4];
  rotated_square_vertices<double>(parity * angle[b], half, vx, vy);
  double* o = cnv + ((long long)b * n_blocks + k) * 8;
#pragma unroll
  for (int l = 0; l < 4; ++l) { o[2 * l] = vx[l]; o[2 * l + 1] = vy[l]; }
  if (inertia) {
    double cx, cy, A, J;
    polygon_props<double>(vx, vy, 4, cx, cy, A, J);
    const double rho = density[(long long)b * density_bstride];
    double* q = inertia + ((long long)b * n_blocks + k) * 3;
    q[0] = rho * A; q[1] = rho * A; q[2] = rho * J;
  }
}

// one CTA per design: angle_bar = sum_blocks parity * d/da (<cnv_bar, cnv> + <inertia_bar, rho (A, A, J)>), one dual pass per block
__global__ void rotated_square_vjp_kernel(int n1, int n_blocks, double half, const double* angle, const double* density,
                                          long long density_bstride, const double* cnv_bar, const double* inertia_bar,
                                          double* angle_bar, double* density_bar) {
  __shared__ double red[40];
  const int b = blockIdx.x;
  const double rho = density ? density[(long long)b * density_bstride] : 0.0;
  double ga = 0.0, grho = 0.0;
  for (int k = threadIdx.x; k < n_blocks; k += blockDim.x) {
    const double parity = ((k % n1 + k / n1) & 1) ? -1.0 : 1.0;
    Dual vx[4], vy[4];
    rotated_square_vertices<Dual>(Dual(parity * angle[b], parity), half, vx, vy);
    double s = 0.0;
    if (cnv_bar) {
      const double* g = cnv_bar + ((long long)b * n_blocks + k) * 8;
#pragma unroll
      for (int l = 0; l < 4; ++l) s += g[2 * l] * vx[l].d + g[2 * l + 1] * vy[l].d;
    }
    if (inertia_bar) {
      Dual cx, cy, A, J;
      polygon_props<Dual>(vx, vy, 4, cx, cy, A, J);
      const double* g = inertia_bar + ((long long)b * n_blocks + k) * 3;
      s += rho * ((g[0] + g[1]) * A.d + g[2] * J.d);
      grho += (g[0] + g[1]) * A.v + g[2] * J.v;
    }
    ga += s;
  }
  const double ta = block_sum(ga, red);
  if (threadIdx.x == 0) angle_bar[b] = ta;
  if (density_bar) {
    const double tr = block_sum(grho, red);
    if (threadIdx.x == 0) density_bar[b] = tr;
  }
}

}  // namespace dfx
