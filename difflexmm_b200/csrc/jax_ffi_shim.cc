// jax_ffi_shim.cc -- XLA FFI handlers that expose libdfx's C ABI (include/dfx.h) to JAX as custom calls.
//
// NOT built in this image: it needs the XLA FFI headers shipped with a modern jaxlib (`jax.ffi.include_dir()`,
// jax >= 0.4.31) and neither jax nor its headers are installed here.  It IS type-checked by the CPU
// test suite (tests/test_abi.py) against tests/stubs/xla/ffi/api/ffi.h, a stand-in for the FFI surface used here whose
// handler macro verifies that every implementation matches the argument list of its binding.
// Handlers: DfxForward, DfxAdjoint (the odeint call and its custom_vjp backward), DfxObjectiveValue, DfxAdjointObjective
// (objective fused with the adjoint), DfxGeometryForward, DfxGeometryVjp (design maps).
// It is kept in-tree, next to the kernels it binds, so that a maintainer with a JAX install can build it:
//
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c "import jax; print(jax.ffi.include_dir())") \
//       -I include -o libdfx_jax.so difflexmm_b200/csrc/jax_ffi_shim.cc -L difflexmm_b200 -ldfx
//
// and register it from Python:
//
//   jax.ffi.register_ffi_target("dfx_forward", jax.ffi.pycapsule(lib.DfxForward), platform="CUDA")
//   jax.ffi.register_ffi_target("dfx_adjoint", jax.ffi.pycapsule(lib.DfxAdjoint), platform="CUDA")
//
// The two handlers replace the `odeint(rhs, _state0, timepoints, control_params, _inertia, rtol, atol)` call at
// /root/reference/difflexmm/dynamics.py:166 and the backward of its custom_vjp.  Buffers are XLA-owned device
// memory; the topology handle (created once per setup_dynamic_solver call) travels as an int64 attribute.

#include <cstdint>

#include <cuda_runtime_api.h>  // cudaStream_t

#include "xla/ffi/api/ffi.h"

#include "dfx.h"

namespace ffi = xla::ffi;

namespace {

DfxLeaf Leaf(const ffi::Buffer<ffi::F64>& b, int64_t batch) {
  DfxLeaf l;
  l.ptr = b.typed_data();
  // a leaf whose leading dimension equals the batch carries one slice per design, otherwise it is shared
  l.bstride = (b.dimensions().size() > 0 && b.dimensions()[0] == batch) ? (int64_t)(b.element_count() / batch) : 0;
  return l;
}

ffi::Error ForwardImpl(cudaStream_t stream, int64_t topo, int64_t k_per_bond_mask, int64_t damping_per_dof, double rtol,
                       double atol, int64_t init_step_variant, ffi::Buffer<ffi::F64> y0, ffi::Buffer<ffi::F64> ts,
                       ffi::Buffer<ffi::F64> cnv, ffi::Buffer<ffi::F64> ref, ffi::Buffer<ffi::F64> ks,
                       ffi::Buffer<ffi::F64> ksh, ffi::Buffer<ffi::F64> kr, ffi::Buffer<ffi::F64> damping,
                       ffi::Buffer<ffi::F64> inertia, ffi::Buffer<ffi::F64> contact, ffi::Buffer<ffi::F64> drive,
                       ffi::ResultBuffer<ffi::F64> ys, ffi::ResultBuffer<ffi::U8> stats) {
  const int64_t batch = ys->dimensions()[0], n_t = ys->dimensions()[1];
  DfxParams p = {};
  p.centroid_node_vectors = Leaf(cnv, batch); p.reference_vector = Leaf(ref, batch);
  p.k_stretch = Leaf(ks, batch); p.k_shear = Leaf(ksh, batch); p.k_rot = Leaf(kr, batch);
  for (int i = 0; i < 3; ++i) p.k_per_bond[i] = (k_per_bond_mask >> i) & 1;
  p.damping = Leaf(damping, batch); p.damping_per_dof = (int32_t)damping_per_dof;
  p.inertia = Leaf(inertia, batch); p.contact = Leaf(contact, batch); p.drive = Leaf(drive, batch);
  DfxOptions opt = {(int32_t)init_step_variant, 0, 0};
  const int64_t N = y0.dimensions().back();
  int rc = dfx_forward(reinterpret_cast<const DfxTopology*>(topo), &p, (int)batch, y0.typed_data(),
                       y0.dimensions().size() == 2 ? N : 0, ts.typed_data(), ts.dimensions().size() == 2 ? n_t : 0,
                       (int)n_t, rtol, atol, &opt, ys->typed_data(), reinterpret_cast<DfxStats*>(stats->typed_data()),
                       /*workspace=*/nullptr, 0, stream);
  if (rc != DFX_OK) return ffi::Error(ffi::ErrorCode::kInternal, dfx_last_error());
  return ffi::Error::Success();
}

ffi::Error AdjointImpl(cudaStream_t stream, int64_t topo, int64_t k_per_bond_mask, int64_t damping_per_dof, double rtol,
                       double atol, int64_t aug_size, int64_t init_step_variant, ffi::Buffer<ffi::F64> ys,
                       ffi::Buffer<ffi::F64> ts, ffi::Buffer<ffi::F64> g, ffi::Buffer<ffi::F64> cnv,
                       ffi::Buffer<ffi::F64> ref, ffi::Buffer<ffi::F64> ks, ffi::Buffer<ffi::F64> ksh,
                       ffi::Buffer<ffi::F64> kr, ffi::Buffer<ffi::F64> damping, ffi::Buffer<ffi::F64> inertia,
                       ffi::Buffer<ffi::F64> contact, ffi::Buffer<ffi::F64> drive, ffi::ResultBuffer<ffi::F64> y0_bar,
                       ffi::ResultBuffer<ffi::F64> ts_bar, ffi::ResultBuffer<ffi::F64> cnv_bar,
                       ffi::ResultBuffer<ffi::F64> ref_bar, ffi::ResultBuffer<ffi::F64> ks_bar,
                       ffi::ResultBuffer<ffi::F64> ksh_bar, ffi::ResultBuffer<ffi::F64> kr_bar,
                       ffi::ResultBuffer<ffi::F64> damping_bar, ffi::ResultBuffer<ffi::F64> inertia_bar,
                       ffi::ResultBuffer<ffi::F64> contact_bar, ffi::ResultBuffer<ffi::F64> drive_bar,
                       ffi::ResultBuffer<ffi::U8> stats) {
  const int64_t batch = ys.dimensions()[0], n_t = ys.dimensions()[1];
  DfxParams p = {};
  p.centroid_node_vectors = Leaf(cnv, batch); p.reference_vector = Leaf(ref, batch);
  p.k_stretch = Leaf(ks, batch); p.k_shear = Leaf(ksh, batch); p.k_rot = Leaf(kr, batch);
  for (int i = 0; i < 3; ++i) p.k_per_bond[i] = (k_per_bond_mask >> i) & 1;
  p.damping = Leaf(damping, batch); p.damping_per_dof = (int32_t)damping_per_dof;
  p.inertia = Leaf(inertia, batch); p.contact = Leaf(contact, batch); p.drive = Leaf(drive, batch);
  DfxParamGrads gr = {cnv_bar->typed_data(), ref_bar->typed_data(), ks_bar->typed_data(), ksh_bar->typed_data(),
                      kr_bar->typed_data(), damping_bar->typed_data(), inertia_bar->typed_data(),
                      contact_bar->typed_data(), drive_bar->typed_data()};
  DfxOptions opt = {(int32_t)init_step_variant, 0, 0};
  int rc = dfx_adjoint(reinterpret_cast<const DfxTopology*>(topo), &p, (int)batch, ys.typed_data(), ts.typed_data(),
                       ts.dimensions().size() == 2 ? n_t : 0, (int)n_t, g.typed_data(), rtol, atol, aug_size, &opt,
                       y0_bar->typed_data(), ts_bar->typed_data(), &gr, reinterpret_cast<DfxStats*>(stats->typed_data()),
                       /*workspace=*/nullptr, 0, stream);
  if (rc != DFX_OK) return ffi::Error(ffi::ErrorCode::kInternal, dfx_last_error());
  return ffi::Error::Success();
}

// objective value J[b] (+ explicit dJ/d inertia, dJ/d arm) from a trajectory: dfx_objective
ffi::Error ObjectiveImpl(cudaStream_t stream, int64_t topo, int64_t kind, ffi::Buffer<ffi::F64> ys, ffi::Buffer<ffi::F64> inertia,
                         ffi::Buffer<ffi::S32> target_free_ids, ffi::Buffer<ffi::F64> arm, ffi::ResultBuffer<ffi::F64> value,
                         ffi::ResultBuffer<ffi::F64> inertia_bar, ffi::ResultBuffer<ffi::F64> arm_bar) {
  const int64_t batch = ys.dimensions()[0], n_t = ys.dimensions()[1];
  DfxParams p = {};
  p.inertia = Leaf(inertia, batch);
  DfxObjective obj = {};
  obj.kind = (int32_t)kind; obj.n_target = (int32_t)target_free_ids.element_count();
  obj.target_free_ids = target_free_ids.typed_data();
  obj.arm = arm.typed_data();
  obj.arm_bstride = (arm.dimensions().size() == 3 && arm.dimensions()[0] == batch) ? (int64_t)(arm.element_count() / batch) : 0;
  int rc = dfx_objective(reinterpret_cast<const DfxTopology*>(topo), &p, (int)batch, ys.typed_data(), (int)n_t, &obj,
                         value->typed_data(), inertia_bar->typed_data(), kind == DFX_OBJ_ANGULAR ? arm_bar->typed_data() : nullptr,
                         stream);
  if (rc != DFX_OK) return ffi::Error(ffi::ErrorCode::kInternal, dfx_last_error());
  return ffi::Error::Success();
}

// adjoint with the objective's cotangent formed in the kernel (weights = dL/dJ): dfx_adjoint_objective
ffi::Error AdjointObjectiveImpl(cudaStream_t stream, int64_t topo, int64_t kind, int64_t k_per_bond_mask, int64_t damping_per_dof,
                                double rtol, double atol, int64_t aug_size, int64_t init_step_variant, ffi::Buffer<ffi::F64> ys,
                                ffi::Buffer<ffi::F64> ts, ffi::Buffer<ffi::F64> weights, ffi::Buffer<ffi::S32> target_free_ids,
                                ffi::Buffer<ffi::F64> arm, ffi::Buffer<ffi::F64> cnv, ffi::Buffer<ffi::F64> ref,
                                ffi::Buffer<ffi::F64> ks, ffi::Buffer<ffi::F64> ksh, ffi::Buffer<ffi::F64> kr,
                                ffi::Buffer<ffi::F64> damping, ffi::Buffer<ffi::F64> inertia, ffi::Buffer<ffi::F64> contact,
                                ffi::Buffer<ffi::F64> drive, ffi::ResultBuffer<ffi::F64> y0_bar, ffi::ResultBuffer<ffi::F64> ts_bar,
                                ffi::ResultBuffer<ffi::F64> cnv_bar, ffi::ResultBuffer<ffi::F64> ref_bar,
                                ffi::ResultBuffer<ffi::F64> ks_bar, ffi::ResultBuffer<ffi::F64> ksh_bar,
                                ffi::ResultBuffer<ffi::F64> kr_bar, ffi::ResultBuffer<ffi::F64> damping_bar,
                                ffi::ResultBuffer<ffi::F64> inertia_bar, ffi::ResultBuffer<ffi::F64> contact_bar,
                                ffi::ResultBuffer<ffi::F64> drive_bar, ffi::ResultBuffer<ffi::U8> stats) {
  const int64_t batch = ys.dimensions()[0], n_t = ys.dimensions()[1];
  DfxParams p = {};
  p.centroid_node_vectors = Leaf(cnv, batch); p.reference_vector = Leaf(ref, batch);
  p.k_stretch = Leaf(ks, batch); p.k_shear = Leaf(ksh, batch); p.k_rot = Leaf(kr, batch);
  for (int i = 0; i < 3; ++i) p.k_per_bond[i] = (k_per_bond_mask >> i) & 1;
  p.damping = Leaf(damping, batch); p.damping_per_dof = (int32_t)damping_per_dof;
  p.inertia = Leaf(inertia, batch); p.contact = Leaf(contact, batch); p.drive = Leaf(drive, batch);
  DfxParamGrads gr = {cnv_bar->typed_data(), ref_bar->typed_data(), ks_bar->typed_data(), ksh_bar->typed_data(),
                      kr_bar->typed_data(), damping_bar->typed_data(), inertia_bar->typed_data(),
                      contact_bar->typed_data(), drive_bar->typed_data()};
  DfxObjective obj = {};
  obj.kind = (int32_t)kind; obj.n_target = (int32_t)target_free_ids.element_count();
  obj.target_free_ids = target_free_ids.typed_data(); obj.weights = weights.typed_data();
  obj.arm = arm.typed_data();
  obj.arm_bstride = (arm.dimensions().size() == 3 && arm.dimensions()[0] == batch) ? (int64_t)(arm.element_count() / batch) : 0;
  DfxOptions opt = {(int32_t)init_step_variant, 0, 0};
  int rc = dfx_adjoint_objective(reinterpret_cast<const DfxTopology*>(topo), &p, (int)batch, ys.typed_data(), ts.typed_data(),
                                 ts.dimensions().size() == 2 ? n_t : 0, (int)n_t, &obj, rtol, atol, aug_size, &opt,
                                 y0_bar->typed_data(), ts_bar->typed_data(), &gr, reinterpret_cast<DfxStats*>(stats->typed_data()),
                                 /*workspace=*/nullptr, 0, stream);
  if (rc != DFX_OK) return ffi::Error(ffi::ErrorCode::kInternal, dfx_last_error());
  return ffi::Error::Success();
}

// design -> (centroid_node_vectors, centroid shift, inertia) and its VJP: dfx_geometry_forward / dfx_geometry_vjp
ffi::Error GeometryForwardImpl(cudaStream_t stream, int64_t geo, ffi::Buffer<ffi::F64> design, ffi::Buffer<ffi::F64> density,
                               ffi::ResultBuffer<ffi::F64> cnv, ffi::ResultBuffer<ffi::F64> centroid_shift,
                               ffi::ResultBuffer<ffi::F64> inertia) {
  const int64_t batch = design.dimensions()[0];
  int rc = dfx_geometry_forward(reinterpret_cast<const DfxGeometry*>(geo), (int)batch, design.typed_data(), density.typed_data(),
                                density.element_count() == (size_t)batch && batch > 1 ? 1 : 0, cnv->typed_data(),
                                centroid_shift->typed_data(), inertia->typed_data(), stream);
  if (rc != DFX_OK) return ffi::Error(ffi::ErrorCode::kInternal, dfx_last_error());
  return ffi::Error::Success();
}

ffi::Error GeometryVjpImpl(cudaStream_t stream, int64_t geo, ffi::Buffer<ffi::F64> design, ffi::Buffer<ffi::F64> density,
                           ffi::Buffer<ffi::F64> cnv_bar, ffi::Buffer<ffi::F64> centroid_bar, ffi::Buffer<ffi::F64> inertia_bar,
                           ffi::ResultBuffer<ffi::F64> design_bar, ffi::ResultBuffer<ffi::F64> density_bar) {
  const int64_t batch = design.dimensions()[0];
  int rc = dfx_geometry_vjp(reinterpret_cast<const DfxGeometry*>(geo), (int)batch, design.typed_data(), density.typed_data(),
                            density.element_count() == (size_t)batch && batch > 1 ? 1 : 0, cnv_bar.typed_data(),
                            centroid_bar.typed_data(), inertia_bar.typed_data(), design_bar->typed_data(),
                            density_bar->typed_data(), stream);
  if (rc != DFX_OK) return ffi::Error(ffi::ErrorCode::kInternal, dfx_last_error());
  return ffi::Error::Success();
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    DfxObjectiveValue, ObjectiveImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<int64_t>("topology").Attr<int64_t>("kind")
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::F64>>()  // ys, inertia, ids, arm
        .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>());                          // J, inertia_bar, arm_bar

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    DfxAdjointObjective, AdjointObjectiveImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<int64_t>("topology").Attr<int64_t>("kind").Attr<int64_t>("k_per_bond_mask").Attr<int64_t>("damping_per_dof")
        .Attr<double>("rtol").Attr<double>("atol").Attr<int64_t>("aug_size").Attr<int64_t>("init_step_variant")
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()     // ys, ts, weights
        .Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::F64>>()                                  // target ids, arm
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()     // cnv, ref, ks
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()     // ksh, kr, damping
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()     // inertia, contact, drive
        .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()                                  // y0_bar, ts_bar
        .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()     // cnv_bar, ref_bar, ks_bar
        .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()     // ksh_bar, kr_bar, damping_bar
        .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()     // inertia_bar, contact_bar, drive_bar
        .Ret<ffi::Buffer<ffi::U8>>());                                                              // stats

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    DfxGeometryForward, GeometryForwardImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<int64_t>("geometry")
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()                                  // design (B, n_design, 2), density
        .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>());   // cnv, centroid shift, inertia

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    DfxGeometryVjp, GeometryVjpImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<int64_t>("geometry")
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()                                  // design, density
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()     // cnv_bar, centroid_bar, inertia_bar
        .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>());                                // design_bar, density_bar

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    DfxForward, ForwardImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<int64_t>("topology").Attr<int64_t>("k_per_bond_mask").Attr<int64_t>("damping_per_dof")
        .Attr<double>("rtol").Attr<double>("atol").Attr<int64_t>("init_step_variant")
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()                                  // y0, ts
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()     // cnv, ref, ks
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()     // ksh, kr, damping
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()     // inertia, contact, drive
        .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::U8>>());                                 // ys, stats

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    DfxAdjoint, AdjointImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<int64_t>("topology").Attr<int64_t>("k_per_bond_mask").Attr<int64_t>("damping_per_dof")
        .Attr<double>("rtol").Attr<double>("atol").Attr<int64_t>("aug_size").Attr<int64_t>("init_step_variant")
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()     // ys, ts, g
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()     // cnv, ref, ks
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()     // ksh, kr, damping
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()     // inertia, contact, drive
        .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()                                  // y0_bar, ts_bar
        .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()     // cnv_bar, ref_bar, ks_bar
        .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()     // ksh_bar, kr_bar, damping_bar
        .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()     // inertia_bar, contact_bar, drive_bar
        .Ret<ffi::Buffer<ffi::U8>>());                                                              // stats
