// bond_phase_microbench.cu -- how much does phase B of the adjoint (dual-number ligament + contact gradient of 768 bonds
// per design) gain from more warps with fewer registers?  Variant A mirrors the fast adjoint kernel (384 threads x 2 bonds,
// <= 168 registers, 12 warps/SM); variant B gives every bond its own thread (768 threads, <= 85 registers, 24 warps/SM).
// Both run one CTA per SM with the block states, bond constants and result slots in shared memory and a CTA barrier
// between "phases", like the real kernel.  Diagnostic for the round-2 redesign (profiles/r01_perf_log.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/bond_mb tools/bond_phase_microbench.cu && /tmp/bond_mb
#include <cstdio>
#include <vector>

#include "../difflexmm_b200/csrc/dfx_device.cuh"

using namespace dfx;

constexpr int NBLK = 384, NBND = 768;

template <int NBT, int TT>
__global__ void __launch_bounds__(TT, 1) bond_phase(const double* in, double* out, int iters, double cmin, double ccut, double ckc) {
  extern __shared__ double sm[];
  double* Us = sm;                 // [5][NBLK]
  double* Ws = Us + 5 * NBLK;      // [3][NBLK]
  double* BC = Ws + 3 * NBLK;      // [10][NBND] bond constants
  double* SL = BC + 10 * NBND;     // [14][NBND] results
  const int tid = threadIdx.x;
  for (int i = tid; i < 8 * NBLK + 10 * NBND; i += TT) sm[i] = in[i];
  __syncthreads();
  double acc = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NBT; ++i) {
      const int b = tid + i * TT;
      const int b1 = b % NBLK, b2 = (b * 7 + 3) % NBLK;
      BlockState<Dual> s1, s2;
      make_block(Us[b1], Us[NBLK + b1], Us[2 * NBLK + b1], Us[3 * NBLK + b1], Us[4 * NBLK + b1], Ws[b1], Ws[NBLK + b1], Ws[2 * NBLK + b1], s1);
      make_block(Us[b2], Us[NBLK + b2], Us[2 * NBLK + b2], Us[3 * NBLK + b2], Us[4 * NBLK + b2], Ws[b2], Ws[NBLK + b2], Ws[2 * NBLK + b2], s2);
      double c[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) c[k] = BC[k * NBND + b];
      BondConst bc = {c[0], c[1], c[2], c[3]};
      BondOut<Dual> o;
      bond_gradient<Dual, true>(DFX_BOND_LIGAMENT, s1, s2, c[4], c[5], c[6], c[7], bc, 120.0, 1.19, 1.5, o);
      Dual psi1 = wrapT(s1.th - s2.th + c[8]);
      Dual psi2 = wrapT(s2.th - s1.th + c[9]);
      double a1 = 0.0, a2 = 0.0;
      const bool act1 = !(psi1.v < cmin) && psi1.v < ccut, act2 = !(psi2.v < cmin) && psi2.v < ccut;
      if (act1 || act2) {
        Dual e1, e2, m1, m2, c1, c2, k1, k2;
        contact_term<Dual>(psi1, cmin, ccut, ckc, e1, m1, c1, k1);
        contact_term<Dual>(psi2, cmin, ccut, ckc, e2, m2, c2, k2);
        o.f1[2] = o.f1[2] + e1 - e2;
        o.f2[2] = o.f2[2] + e2 - e1;
        a1 = e1.d; a2 = e2.d;
        acc -= m1.d + m2.d + c1.d + c2.d + k1.d + k2.d;
      }
      SL[b] = o.f2[0].v; SL[NBND + b] = o.f2[1].v; SL[2 * NBND + b] = -o.f1[2].v; SL[3 * NBND + b] = -o.f2[2].v;
      SL[4 * NBND + b] = o.f2[0].d; SL[5 * NBND + b] = o.f2[1].d; SL[6 * NBND + b] = o.f1[2].d; SL[7 * NBND + b] = o.f2[2].d;
      SL[8 * NBND + b] = -o.gr1[0].d; SL[9 * NBND + b] = -o.gr1[1].d; SL[10 * NBND + b] = -o.gr2[0].d; SL[11 * NBND + b] = -o.gr2[1].d;
      SL[12 * NBND + b] = a1; SL[13 * NBND + b] = a2;
      acc += o.gr0[0].d + o.gr0[1].d + o.gks.d + o.gksh.d + o.gkr.d;
    }
    __syncthreads();
    if (tid < NBLK) {  // stand-in for the unit phase: feed a little of the result back so that nothing is hoisted
      const double f = SL[tid] + SL[NBND + tid + NBLK] + SL[4 * NBND + tid];
      Us[tid] += 1e-12 * f;
      Ws[tid] += 1e-12 * SL[5 * NBND + tid];
    }
    __syncthreads();
  }
  out[blockIdx.x * TT + tid] = acc + Us[tid % NBLK];
}

int main() {
  const int n_in = 8 * NBLK + 10 * NBND;
  std::vector<double> h(n_in);
  unsigned s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / (double)(1 << 24) - 0.5; };
  for (int i = 0; i < NBLK; ++i) {
    const double th = 0.2 * rnd();
    h[i] = 0.5 * rnd(); h[NBLK + i] = 0.5 * rnd(); h[2 * NBLK + i] = th; h[3 * NBLK + i] = sin(th); h[4 * NBLK + i] = cos(th);
    h[5 * NBLK + i] = rnd(); h[6 * NBLK + i] = rnd(); h[7 * NBLK + i] = rnd();
  }
  double* bcst = h.data() + 8 * NBLK;
  for (int b = 0; b < NBND; ++b) {
    const double rx = 2.25, ry = 0.1 * rnd(), L0 = sqrt(rx * rx + ry * ry);
    bcst[b] = rx; bcst[NBND + b] = ry; bcst[2 * NBND + b] = L0; bcst[3 * NBND + b] = 1.0 / L0;
    bcst[4 * NBND + b] = 6.0 + rnd(); bcst[5 * NBND + b] = rnd(); bcst[6 * NBND + b] = -6.0 + rnd(); bcst[7 * NBND + b] = rnd();
    bcst[8 * NBND + b] = 0.7 + 0.1 * rnd(); bcst[9 * NBND + b] = 2.4 + 0.1 * rnd();
  }
  double *din, *dout;
  cudaMalloc(&din, n_in * sizeof(double));
  cudaMalloc(&dout, 148 * 768 * sizeof(double));
  cudaMemcpy(din, h.data(), n_in * sizeof(double), cudaMemcpyHostToDevice);
  const size_t smem = (size_t)(8 * NBLK + 24 * NBND) * sizeof(double);
  cudaFuncSetAttribute(bond_phase<2, 384>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(bond_phase<1, 768>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 20000;
  for (int contact_on = 0; contact_on < 2; ++contact_on) {
    // contact window: inactive (as for most bonds of the regular design) or active for every bond
    const double cmin = contact_on ? 0.0 : -0.26, ccut = contact_on ? 3.0 : -0.17;
    for (int variant = 0; variant < 2; ++variant) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (variant == 0) bond_phase<2, 384><<<148, 384, smem>>>(din, dout, iters, cmin, ccut, 1.5);
        else bond_phase<1, 768><<<148, 768, smem>>>(din, dout, iters, cmin, ccut, 1.5);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      cudaError_t err = cudaGetLastError();
      printf("{\"contact\": \"%s\", \"variant\": \"%s\", \"us_per_phase\": %.3f, \"error\": \"%s\"}\n", contact_on ? "active" : "inactive",
             variant == 0 ? "384 threads x 2 bonds (<=168 regs, 12 warps)" : "768 threads x 1 bond (<=85 regs, 24 warps)",
             1e3 * ms / iters, cudaGetErrorString(err));
    }
  }
  return 0;
}
