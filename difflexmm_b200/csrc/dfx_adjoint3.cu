// dfx_adjoint3.cu -- instantiations and launcher of the 24-warp adjoint kernel (dfx_adjoint3.cuh); its own translation
// unit so that the library builds in parallel.
#include "dfx_adjoint3.cuh"

namespace dfx {

size_t adjoint3_smem_bytes(bool contact) {
  return (size_t)(contact ? k3::Lay<14>::END : k3::Lay<12>::END) * sizeof(double);
}
long long adjoint3_scratch_doubles() { return k3::G_TOTAL; }

template <int NPB, bool CONTACT, int DAMP>
static cudaError_t launch_one(const Adj3Args& A, int batch, cudaStream_t stream) {
  const size_t smem = adjoint3_smem_bytes(CONTACT);
  cudaError_t e = cudaFuncSetAttribute(adjoint3_kernel<NPB, CONTACT, DAMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  adjoint3_kernel<NPB, CONTACT, DAMP><<<batch, k3::TT, smem, stream>>>(A);
  return cudaGetLastError();
}

bool adjoint3_supported(int npb, bool contact, int damp) {
#ifdef DFX_A3_MAIN_ONLY
  return npb == 4 && contact && damp == 2;
#else
  (void)contact;
  return (npb == 4 || npb == 3) && damp >= 0 && damp <= 2;
#endif
}

cudaError_t launch_adjoint3(const Adj3Args& A, int npb, bool contact, int damp, int batch, cudaStream_t stream) {
#define DFX_A3_CASE(N, C, D) if (npb == N && contact == C && damp == D) return launch_one<N, C, D>(A, batch, stream)
  DFX_A3_CASE(4, true, 2);
#ifndef DFX_A3_MAIN_ONLY  // experiments build the bench instance only
  DFX_A3_CASE(4, false, 2);
  DFX_A3_CASE(4, true, 1); DFX_A3_CASE(4, false, 1);
  DFX_A3_CASE(4, true, 0); DFX_A3_CASE(4, false, 0);
  DFX_A3_CASE(3, true, 2); DFX_A3_CASE(3, false, 2);
  DFX_A3_CASE(3, true, 1); DFX_A3_CASE(3, false, 1);
  DFX_A3_CASE(3, true, 0); DFX_A3_CASE(3, false, 0);
#endif
#undef DFX_A3_CASE
  return cudaErrorInvalidValue;
}

}  // namespace dfx
