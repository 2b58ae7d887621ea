/* dfx.h -- C ABI of the B200-native DifFlexMM dynamic solver (libdfx.so).
 *
 * Drop-in boundary: these entry points replace the call
 *     odeint(rhs, _state0, timepoints, control_params, _inertia, rtol, atol)
 * at /root/reference/difflexmm/dynamics.py:166 and what jax.grad derives for it
 * (jax.experimental.ode._odeint_rev, jax 0.4.8), i.e. reference functions a3-a12 of
 * SURVEY.md section 8(a):
 *   - dfx_forward  <- odeint forward solve              (dynamics.py:166, rhs at :33-55)
 *   - dfx_adjoint  <- the custom_vjp backward of odeint (continuous adjoint, restarted
 *                     per output interval; cotangents for y0, ts, every ControlParams
 *                     leaf and the reduced inertia)
 *   - dfx_expand_fields <- the history reconstruction    (dynamics.py:129-136,169-182)
 *   - dfx_topology_create <- what setup_dynamic_solver closes over (dynamics.py:60-136):
 *                     bond list, constrained DOF pairs and their drive signal, loaded
 *                     DOFs and their load signal, damped blocks, energy vocabulary.
 *
 * Plain C: pointers + sizes, no C++/torch types.  All array arguments of
 * dfx_forward / dfx_adjoint / dfx_expand_fields are DEVICE pointers (float64 unless
 * noted); the DfxTopologyDesc passed to dfx_topology_create holds HOST pointers and is
 * copied.  The caller owns every buffer.  Calls are stream-ordered and re-entrant; a
 * topology handle is immutable and may be shared between threads and streams of the
 * device it was created on.
 *
 * Layouts (row-major, batch index b outermost):
 *   state vectors   [2*n_free]   = free displacements then free velocities, i.e. the
 *                                  raveled (2, n_free) array the reference hands to odeint
 *   ys, g           [B][n_t][2*n_free]
 *   every parameter leaf has a batch stride in elements; stride 0 = shared by all designs
 */
#ifndef DFX_H
#define DFX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- vocabulary ------------------------------------------ */
enum { DFX_BOND_LIGAMENT = 0,    /* energy.py:158-176 ligament_energy            */
       DFX_BOND_LINEARIZED = 1,  /* energy.py:99-117  ligament_energy_linearized */
       DFX_BOND_SPRING = 2 };    /* energy.py:49-66   stretching_torsional_spring_energy: zero-length spring between
                                    coincident nodes, k_stretch |dU|^2 / 2 + k_rot dtheta^2 / 2.  The k_shear and
                                    reference_vector leaves are ignored (they must still be valid pointers; their
                                    cotangents are zero).  Generic kernels only. */

/* constrained-DOF drive u_c(t) = vec0[c]*s0(t) + vec1[c]*s1(t); parameter order fixed */
enum {
  DFX_DRIVE_ZERO = 0,     /* dynamics.py:66 default `lambda t: 0`; 0 params                  */
  DFX_DRIVE_PULSE = 1,    /* problems/quads_focusing.py:211-222; (amplitude, loading_rate,
                             input_delay): A*(1-cos(2 pi f tau))/2 on 0<tau<1/f, tau=t-delay  */
  DFX_DRIVE_HARMONIC = 2, /* problems/quads_spin.py:210-221; same params, on tau>0          */
  DFX_DRIVE_RAMP = 3,     /* problems/hinge_characterization.py:134-139; (amplitude,
                             loading_rate): A*(t<1/f ? t*f : 1)                              */
  DFX_DRIVE_TABLE = 5,    /* `excited_blocks_fn = jnp.interp(t, table_t, table_v)` of the experiment notebooks
                             (exp/.../experiment_vs_simulation.ipynb cell 12; problems/quads_focusing.py:223-227):
                             s0 = interp(t), constant outside the table, 0 params                           */
  DFX_DRIVE_STATIC_PULSE = 4 /* problems/quads_kinetic_energy_static_tuning.py:176-196;
                             (amplitude, loading_rate, compressive_strain,
                             compressive_strain_rate, input_delay):
                             s0 = pulse(t - cs/csr - delay; A, f)  [vec0 = dynamic vector]
                             s1 = (t < cs/csr ? t*csr : cs)        [vec1 = static vector,
                                                                    geometric scale folded in] */
};
#define DFX_MAX_DRIVE_PARAMS 5

/* external load on loaded DOFs: load_vec[l] * s(t); constants only (the reference's
 * loading closures capture their constants: tests/test_difflexmm.py:85-86,
 * scripts/pulse_RS.py:49-50) */
enum {
  DFX_LOAD_NONE = 0,
  DFX_LOAD_RAMP = 1, /* c0 * (t < 1/c1 ? t*c1 : 1)                                   */
  DFX_LOAD_SECH2 = 2 /* 2*c0/c1^2 * cosh(t/c1-3)^-2 * tanh(3-t/c1)                   */
};
#define DFX_MAX_LOAD_CONSTS 4

enum { DFX_CONTACT_NONE = 0, DFX_CONTACT_ANGLE = 1, DFX_CONTACT_DISTANCE = 2 };

/* ---- static topology --------------------------------------------------------------- */
typedef struct DfxTopologyDesc {
  int32_t n_blocks;          /* rigid units                                            */
  int32_t n_npb;             /* nodes (polygon vertices) per block                      */
  int32_t n_bonds;
  const int32_t* bond_nodes; /* [n_bonds][2] global node ids (block*n_npb + local)      */
  int32_t n_constrained;
  const int32_t* constrained_dofs; /* [n_constrained] global DOF ids block*3+dof, in the
                                      order of constrained_block_DOF_pairs               */
  int32_t bond_energy;       /* DFX_BOND_*                                              */
  int32_t contact;           /* DFX_CONTACT_*: 0 none, 1 angle-based (energy.py:204-219, 364-407),
                                2 distance-based between the void edges (energy.py:222-330; generic kernels) */
  int32_t drive_kind;        /* DFX_DRIVE_*                                             */
  const double* drive_vec0;  /* [n_constrained] or NULL (= zeros)                       */
  const double* drive_vec1;  /* [n_constrained] or NULL                                 */
  int32_t drive_table_len;   /* DFX_DRIVE_TABLE: samples of the tabulated signal        */
  const double* drive_table_t; /* [drive_table_len] increasing times                    */
  const double* drive_table_v; /* [drive_table_len] values                              */
  int32_t load_kind;         /* DFX_LOAD_*                                              */
  int32_t n_loaded;
  const int32_t* loaded_dofs;/* [n_loaded] global DOF ids                               */
  const double* load_vec;    /* [n_loaded] multiplier per loaded DOF, or NULL (= ones)  */
  double load_consts[DFX_MAX_LOAD_CONSTS];
  int32_t n_damped;          /* 0 = no damping term                                     */
  const int32_t* damped_blocks; /* [n_damped] block ids (loading.py:88-89)              */
} DfxTopologyDesc;

typedef struct DfxTopology DfxTopology; /* opaque */

/* ---- runtime parameters: the differentiable leaves of (control_params, _inertia) ---- */
typedef struct DfxLeaf {
  const double* ptr; /* device */
  int64_t bstride;   /* elements between consecutive designs; 0 = shared               */
} DfxLeaf;

typedef struct DfxParams {
  DfxLeaf centroid_node_vectors; /* [n_blocks*n_npb*2]                                  */
  DfxLeaf reference_vector;      /* [n_bonds*2]                                         */
  DfxLeaf k_stretch, k_shear, k_rot; /* [1] or [n_bonds], see k_per_bond               */
  int32_t k_per_bond[3];         /* 0 = scalar leaf, 1 = (n_bonds,) leaf                */
  DfxLeaf damping;               /* [1] or [n_damped*3]                                 */
  int32_t damping_per_dof;       /* 0 = scalar leaf, 1 = (n_damped,3) leaf              */
  DfxLeaf inertia;               /* [n_free]  (reduced to the free DOFs, dynamics.py:157-163) */
  DfxLeaf contact;               /* [3] = (min_angle, cutoff_angle, k_contact); NULL if contact==0 */
  DfxLeaf drive;                 /* [n_drive_params(kind)]                              */
  DfxLeaf block_centroids;       /* [n_blocks*2]; read by DFX_CONTACT_DISTANCE only (the other energies depend on
                                    displacements, not positions), NULL otherwise          */
} DfxParams;

/* cotangent outputs of dfx_adjoint, one slice per design (no sharing), same leaf shapes.
 * A NULL pointer skips the store (the quadrature is still integrated: it is part of the
 * error norm of the reference's augmented system). */
typedef struct DfxParamGrads {
  double* centroid_node_vectors; /* [B][n_blocks*n_npb*2] */
  double* reference_vector;      /* [B][n_bonds*2]        */
  double* k_stretch;             /* [B][1 or n_bonds]     */
  double* k_shear;
  double* k_rot;
  double* damping;               /* [B][1 or n_damped*3]  */
  double* inertia;               /* [B][n_free]           */
  double* contact;               /* [B][3]                */
  double* drive;                 /* [B][n_drive_params]   */
  double* block_centroids;       /* [B][n_blocks*2]; written with DFX_CONTACT_DISTANCE only (zero otherwise: leave NULL) */
} DfxParamGrads;

typedef struct DfxOptions {
  int32_t init_step_variant; /* 0: h1 = (0.01/(d1+d2))^(1/5)  (jax 0.4.8, the reference's pin)
                                1: h1 = (0.01/max(d1,d2))^(1/5) (later jax releases)     */
  int32_t threads;           /* CTA size override, 0 = choose                           */
  int64_t max_steps;         /* attempted-step cap per odeint call, 0 = 1<<40 (jax: inf)*/
  const int32_t* design_order; /* device array [batch] or NULL: CTA (group) k of the launch works on design
                                design_order[k].  A permutation of 0..batch-1 that puts the designs expected to take
                                longest first (e.g. by the step counts of the forward solve) shortens the tail of a
                                launch with more designs than SMs; the results do not depend on it.            */
} DfxOptions;

enum { DFX_OK = 0, DFX_ERR_INVALID = 1, DFX_ERR_CUDA = 2, DFX_ERR_UNSUPPORTED = 3 };
enum { DFX_STATUS_OK = 0, DFX_STATUS_MAX_STEPS = 1, DFX_STATUS_DT_UNDERFLOW = 2,
       DFX_STATUS_NONFINITE = 4 };

typedef struct DfxStats { /* one per design, device memory */
  int64_t steps;     /* attempted RK steps            */
  int64_t accepted;  /* accepted RK steps             */
  int64_t rhs_evals; /* (augmented) RHS evaluations   */
  int32_t status;    /* DFX_STATUS_* bit-or           */
  int32_t reserved;
  double last_dt;
} DfxStats;

/* ---- entry points ------------------------------------------------------------------ */
int dfx_topology_create(const DfxTopologyDesc* desc, int device, DfxTopology** out);
void dfx_topology_destroy(DfxTopology* topo);
int dfx_topology_n_free(const DfxTopology* topo);
int dfx_drive_n_params(int drive_kind);

/* bytes of device scratch dfx_adjoint / dfx_forward need for `batch` designs */
size_t dfx_forward_workspace_bytes(const DfxTopology* topo, int batch);
size_t dfx_adjoint_workspace_bytes(const DfxTopology* topo, int batch);

/* forward solve: ys[b][0] = y0[b]; ys[b][i] = state at ts[i].  stream = cudaStream_t. */
int dfx_forward(const DfxTopology* topo, const DfxParams* params, int batch,
                const double* y0, int64_t y0_bstride,
                const double* ts, int64_t ts_bstride, int n_t,
                double rtol, double atol, const DfxOptions* opt,
                double* ys, DfxStats* stats,
                void* workspace, size_t workspace_bytes, void* stream);

/* continuous adjoint.  aug_size = number of entries the reference's augmented state has,
 * 2*(2 n_free) + 1 + sum(size of every leaf of (control_params, _inertia)); it is the
 * denominator of the RMS error norm.  Pass 0 to count only the leaves listed in DfxParams. */
int dfx_adjoint(const DfxTopology* topo, const DfxParams* params, int batch,
                const double* ys, const double* ts, int64_t ts_bstride, int n_t,
                const double* g, double rtol, double atol, int64_t aug_size,
                const DfxOptions* opt,
                double* y0_bar, double* ts_bar, const DfxParamGrads* grads, DfxStats* stats,
                void* workspace, size_t workspace_bytes, void* stream);

/* ---- objectives on the device ------------------------------------------------------------------
 * DFX_OBJ_KINETIC   J_b = w_b sum_t sum_{f in target} 1/2 m_f v_f(t)^2
 *                   (target kinetic energy: problems/quads_focusing.py:453-467, energy.py:494-499)
 * DFX_OBJ_ANGULAR   J_b = w_b sum_t sum_{k in target blocks} (arm_k + u_k(t)) x (m v_k(t))_xy + I_k omega_k(t)
 *                   (angular momentum about a spin centre: problems/quads_spin.py:395-428, energy.py:502-519;
 *                    arm_k = reference centroid of block k minus the spin centre)
 * `target_free_ids` are indices into the free-DOF vector: the three DOFs (x, y, theta) of every target block, block
 * after block; they must be free.  dfx_objective evaluates J (w = 1) and, optionally, its explicit derivatives w.r.t.
 * the reduced inertia and the arms; dfx_adjoint_objective is dfx_adjoint with the cotangent g = dJ/dys generated
 * inside the kernel (g is never materialised: no 3.5 MB per design of cotangent traffic), weights[b] = dL/dJ_b. */
enum { DFX_OBJ_KINETIC = 0, DFX_OBJ_ANGULAR = 1 };
typedef struct DfxObjective {
  int32_t kind;
  int32_t n_target;               /* number of target DOFs (3 per target block) */
  const int32_t* target_free_ids; /* device, [n_target] */
  const double* weights;          /* device, [B], or NULL (= 1) */
  const double* arm;              /* device, [B or 1][n_target / 3][2]; DFX_OBJ_ANGULAR only */
  int64_t arm_bstride;            /* doubles between designs (0 = shared) */
} DfxObjective;

int dfx_objective(const DfxTopology* topo, const DfxParams* params, int batch, const double* ys, int n_t,
                  const DfxObjective* obj, double* value /*[B]*/, double* inertia_bar /*[B][n_free] or NULL*/,
                  double* arm_bar /*[B][n_target / 3][2] or NULL*/, void* stream);

int dfx_adjoint_objective(const DfxTopology* topo, const DfxParams* params, int batch,
                          const double* ys, const double* ts, int64_t ts_bstride, int n_t,
                          const DfxObjective* obj, double rtol, double atol, int64_t aug_size,
                          const DfxOptions* opt,
                          double* y0_bar, double* ts_bar, const DfxParamGrads* grads, DfxStats* stats,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- design -> solver parameters on the device ------------------------------------------------
 * Replaces the reference's design maps and their JAX-derived VJPs for lattices whose every polygon vertex is
 * `base_node + design[node_design]`: QuadGeometry / KagomeGeometry reference_node_vectors, centroid_node_vectors,
 * block_centroids (geometry.py:607-952) and compute_inertia (geometry.py:71-160).
 *   cnv[b][n][:]            = vertex - polygon centroid
 *   centroid_shift[b][k][:] = polygon centroid (block_centroids = lattice reference point + centroid_shift)
 *   inertia[b][k][:]        = density * (area, area, polar moment about the centroid)
 * dfx_geometry_vjp returns design_bar = J^T (cnv_bar, centroid_bar, inertia_bar) and optionally density_bar. */
typedef struct DfxGeometryDesc {
  int32_t n_blocks, n_npb;
  int32_t n_design;            /* number of design 2-vectors */
  const double* base_nodes;    /* host, [n_blocks * n_npb][2] */
  const int32_t* node_design;  /* host, [n_blocks * n_npb]: design 2-vector added to the vertex, or -1 */
} DfxGeometryDesc;
typedef struct DfxGeometry DfxGeometry;

int dfx_geometry_create(const DfxGeometryDesc* desc, int device, DfxGeometry** out);
void dfx_geometry_destroy(DfxGeometry* geo);
int dfx_geometry_forward(const DfxGeometry* geo, int batch, const double* design /*[B][n_design][2]*/,
                         const double* density, int64_t density_bstride /* 0 = shared */,
                         double* cnv, double* centroid_shift /* or NULL */, double* inertia /* or NULL */, void* stream);
int dfx_geometry_vjp(const DfxGeometry* geo, int batch, const double* design, const double* density, int64_t density_bstride,
                     const double* cnv_bar /* or NULL */, const double* centroid_bar /* or NULL */,
                     const double* inertia_bar /* or NULL */, double* design_bar /*[B][n_design][2]*/,
                     double* density_bar /*[B] or NULL*/, void* stream);

/* RotatedSquareGeometry (geometry.py:354-443): the design is ONE angle per lattice.  Block (i1, i2) of the
 * n1_blocks x n2_blocks grid (i1 fastest) has vertices R(l pi/2) * half_side * (1, tan(+-angle)), sign (-1)^(i1+i2),
 * half_side = (spacing - bond_length) / 2; inertia = density * (area, area, polar moment) (geometry.py:71-160).
 * The VJP returns angle_bar[b] = <cnv_bar, d cnv/d angle> + <inertia_bar, d inertia/d angle> (and density_bar). */
int dfx_rotated_square_forward(int n1_blocks, int n2_blocks, double half_side, int batch, const double* angle /*[B]*/,
                               const double* density, int64_t density_bstride /* 0 = shared */,
                               double* cnv /*[B][n_blocks][4][2]*/, double* inertia /*[B][n_blocks][3] or NULL*/, void* stream);
int dfx_rotated_square_vjp(int n1_blocks, int n2_blocks, double half_side, int batch, const double* angle, const double* density,
                           int64_t density_bstride, const double* cnv_bar /* or NULL */, const double* inertia_bar /* or NULL */,
                           double* angle_bar /*[B]*/, double* density_bar /*[B] or NULL*/, void* stream);

/* ---- inequality constraints of the design and their Jacobian on the device ----------------------
 * Replaces OptimizationProblem.setup_angle_constraints / setup_edge_length_constraints
 * (problems/quads_focusing.py:473-544) and the jit(jacobian(...)) evaluated inside the nlopt callbacks (:585-588,
 * :613-616) for the lattices a DfxGeometry describes.  Rows (all <= 0 when satisfied), in the reference's order:
 *   angles: -(void_angle_1 - min_void) per bond, -(void_angle_2 - min_void) per bond, -(block_angle_1 - min_block),
 *           -(block_angle_2 - min_block), then -(block angle - min_block) of each boundary node; angles mod 2 pi
 *   edges:  -(edge length - min_edge_length) per polygon vertex (edge to the previous vertex), block-major
 * A row depends on at most 4 design 2-vectors: jac[b][row][slot][xy] with the shared column table
 * columns[row][slot] = design 2-vector index or -1 (empty slot). */
typedef struct DfxConstraintDesc {
  int32_t n_bonds;
  const int32_t* bonds;           /* host, [n_bonds][2] node ids */
  int32_t n_boundary;
  const int32_t* boundary_nodes;  /* host, [n_boundary] or NULL */
  int32_t angles, edges;          /* which row families to build */
  double min_void_angle, min_block_angle, min_edge_length;
} DfxConstraintDesc;
typedef struct DfxConstraints DfxConstraints;

int dfx_constraints_create(const DfxGeometry* geo, const DfxConstraintDesc* desc, DfxConstraints** out); /* geo must outlive it */
void dfx_constraints_destroy(DfxConstraints* con);
int dfx_constraints_rows(const DfxConstraints* con, int32_t* n_angle_rows /* or NULL */); /* -> number of rows */
int dfx_constraints_columns(const DfxConstraints* con, int32_t* columns /* host, [rows][4] */);
int dfx_constraints_eval(const DfxConstraints* con, int batch, const double* design /*[B][n_design][2]*/,
                         double* values /*[B][rows]*/, double* jac /*[B][rows][4][2] or NULL*/, void* stream);

/* fields[b][i][0][blk][dof] = displacement, fields[b][i][1][blk][dof] = velocity of every
 * block DOF (constrained DOFs follow the drive and its time derivative). */
int dfx_expand_fields(const DfxTopology* topo, const DfxParams* params, int batch,
                      const double* ys, const double* ts, int64_t ts_bstride, int n_t,
                      double* fields, void* stream);

/* diagnostic: name of the adjoint kernel that dfx_adjoint / dfx_adjoint_objective would launch for this topology, these
 * leaf forms and this batch size (thread-local string, valid until the next call) */
const char* dfx_adjoint_plan(const DfxTopology* topology, const DfxParams* params, int batch);

/* measurement helper: FP64 FMA throughput of the current device in TFLOP/s (roofline denominator) */
double dfx_fp64_peak(void* stream);

/* test helper: evaluates the device math primitives (1/sqrt(x), angle of the unit vector along (x,y), 1/x) */
int dfx_math_selftest(const double* x, const double* y, double* out_rsqrt, double* out_angle, double* out_rcp, int n, void* stream);
/* test helper: the kernels' sin / cos of a block rotation (quadrant reduction + fdlibm kernel polynomials) */
int dfx_sincos_selftest(const double* x, double* out_sin, double* out_cos, int n, void* stream);

const char* dfx_last_error(void);
const char* dfx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DFX_H */
