// dfx_api.cu -- C ABI of libdfx.so (see include/dfx.h): topology handles, shared-memory
// placement planning and kernel launches.  No torch / C++ types cross the boundary.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include <cstdlib>
#include <type_traits>

#include "dfx_forward2.cuh"
#include "dfx_geometry.cuh"
#include "dfx_adjoint3.cuh"

namespace dfx {  // dfx_adjoint3.cu
size_t adjoint3_smem_bytes(bool contact);
long long adjoint3_scratch_doubles();
bool adjoint3_supported(int npb, bool contact, int damp);
cudaError_t launch_adjoint3(const Adj3Args& A, int npb, bool contact, int damp, int batch, cudaStream_t stream);
}

using namespace dfx;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(x)                                                                                   \
  do {                                                                                                \
    cudaError_t e_ = (x);                                                                             \
    if (e_ != cudaSuccess) return fail(DFX_ERR_CUDA, "%s failed: %s", #x, cudaGetErrorString(e_));    \
  } while (0)

// The stream-ordered pool of the device keeps what it has allocated (release threshold = unlimited).  With the default
// threshold of 0 every synchronisation hands the scratch of dfx_geometry_vjp (and of workspace-less solver calls) back to the
// OS and the next call maps it again: that re-mapping stalled one end-to-end step in ten by 200 - 800 ms (tools/e2e_jitter.py).
void keep_pool_memory(int device) {
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t threshold = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  }
  cudaGetLastError();
}

constexpr size_t kSmemBytes = 232448;  // 227 KB: the opt-in maximum of one CTA on sm_100
constexpr int kRedDoubles = 40;
static_assert(kRedDoubles == kRedDoublesDev, "shared-memory reduction area");
constexpr int kGroupMinBonds = 8192;  // lattices with at least this many bonds default to the cooperative group mode
constexpr int kGroupDefault = kMaxGroup;  // CTAs per design in that mode: as many as fit (100x100: 64 CTAs 127 ms, 148 CTAs 110 ms adjoint)
constexpr int kScratchSlots = 256;  // >= %nsmid of any sm_100 part: SM-indexed scratch of the fast adjoint kernel

template <class T>
cudaError_t upload(const std::vector<T>& h, T** d) {
  *d = nullptr;
  if (h.empty()) return cudaSuccess;
  cudaError_t e = cudaMalloc((void**)d, h.size() * sizeof(T));
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

}  // namespace

struct DfxTopology {
  int device;
  DevTopo dev;
  const int* node_bond;  // [n_nodes] bond*2+side or -1 (fast adjoint kernel)
  int n_cons_units;      // rigid units with at least one constrained DOF
  std::vector<void*> allocs;
  int sm_count;
};

namespace {

int n_drive_params_of(int kind) {
  switch (kind) {
    case DFX_DRIVE_PULSE:
    case DFX_DRIVE_HARMONIC: return 3;
    case DFX_DRIVE_RAMP: return 2;
    case DFX_DRIVE_STATIC_PULSE: return 5;
    default: return 0;
  }
}

// greedy placement: arrays in priority order go to shared memory while they fit
// cluster launches (cluster > 1): everything goes to the global scratch, whose first 2 * kMaxCluster doubles hold
// the cluster-sum partials
template <int N>
void plan(const long long (&sizes)[N], long long* off, size_t* smem_bytes, long long* scratch_doubles, int cluster = 1) {
  long long s = kRedDoubles, g = cluster > 1 ? kClusterReserve : 0;
  const long long cap = cluster > 1 ? s : (long long)(kSmemBytes / sizeof(double));
  for (int i = 0; i < N; ++i) {
    const long long n = (sizes[i] + 1) & ~1LL;  // keep 16-byte alignment
    if (n == 0) { off[i] = 0; continue; }
    if (s + n <= cap) { off[i] = s; s += n; }
    else { off[i] = -(g + 1); g += n; }
  }
  *smem_bytes = (size_t)s * sizeof(double);
  *scratch_doubles = g;
}

void forward_sizes(const DevTopo& T, long long (&sz)[FA_COUNT]) {
  const long long NB = T.n_blocks, NN = T.n_nodes, NBONDS = T.n_bonds;
  sz[FA_US] = 5 * NB; sz[FA_VS] = 3 * NB; sz[FA_FS] = 3 * NN; sz[FA_U0] = 3 * NB; sz[FA_V0] = 3 * NB;
  sz[FA_KV] = 21 * NB; sz[FA_INVM] = 3 * NB; sz[FA_CD] = 3 * NB; sz[FA_BONDC] = 4 * NBONDS; sz[FA_CNV] = 2 * NN;
  sz[FA_ALPHA] = T.contact == DFX_CONTACT_ANGLE ? 2 * NN : 0;
}

struct QuadLayout { int qo_cnv, qo_ref, qo_ks, qo_ksh, qo_kr, qo_damp, qo_inertia, qo_cen, nq; };

QuadLayout quad_layout(const DevTopo& T, const DfxParams& p) {
  QuadLayout q;
  int o = 0;
  q.qo_cnv = o; o += 2 * T.n_nodes;
  q.qo_ref = o; o += 2 * T.n_bonds;
  q.qo_ks = o; o += p.k_per_bond[0] ? T.n_bonds : 0;
  q.qo_ksh = o; o += p.k_per_bond[1] ? T.n_bonds : 0;
  q.qo_kr = o; o += p.k_per_bond[2] ? T.n_bonds : 0;
  q.qo_damp = o; o += (T.n_damped > 0 && p.damping_per_dof) ? 3 * T.n_blocks : 0;
  q.qo_inertia = o; o += 3 * T.n_blocks;
  q.qo_cen = o; o += T.contact == DFX_CONTACT_DISTANCE ? 2 * T.n_blocks : 0;  // block_centroids leaf [2][NB]
  q.nq = o;
  return q;
}

void adjoint_sizes(const DevTopo& T, const QuadLayout& q, long long (&sz)[AA_COUNT], int cluster = 1) {
  const long long NB = T.n_blocks, NN = T.n_nodes, NBONDS = T.n_bonds;
  sz[AA_US] = 5 * NB; sz[AA_WS] = 3 * NB; sz[AA_VS] = 3 * NB; sz[AA_LUS] = 3 * NB; sz[AA_LVS] = 3 * NB;
  sz[AA_FS] = 3 * NN; sz[AA_HS] = 3 * NN; sz[AA_GS] = 2 * NN;
  // angle contact: d/d(alpha_next, alpha_prev) per node; distance contact: cotangents of the next / previous vertex (x, y each)
  // and the contact part of the force pair [2][NN] for the block_centroids leaf
  sz[AA_GA] = T.contact == DFX_CONTACT_ANGLE ? 2 * NN : (T.contact == DFX_CONTACT_DISTANCE ? 6 * NN : 0);
  sz[AA_SC] = cluster > 1 ? 0 : kScalDoubles;  // cluster mode keeps it in each CTA's shared memory
  sz[AA_INVM] = 3 * NB; sz[AA_CD] = 3 * NB;
  sz[AA_U0] = 3 * NB; sz[AA_V0] = 3 * NB; sz[AA_LU0] = 3 * NB; sz[AA_LV0] = 3 * NB;
  sz[AA_KV] = 21 * NB; sz[AA_KLU] = 21 * NB; sz[AA_KLV] = 21 * NB;
  sz[AA_BONDC] = 4 * NBONDS; sz[AA_CNV] = 2 * NN; sz[AA_ALPHA] = T.contact == DFX_CONTACT_ANGLE ? 2 * NN : 0;
  sz[AA_EDGED] = T.contact == DFX_CONTACT_ANGLE ? 4 * NN : 0;
  for (int i = AA_QK3; i <= AA_QNEW; ++i) sz[i] = q.nq;
}

int pick_threads(const DevTopo& T, int requested) {
  int t = requested;
  if (t <= 0) {
    // bonds dominate the cost: choose the CTA size (multiple of 32, <= 512) that wastes the fewest
    // lanes in the bond loop, breaking ties towards more threads
    double best = 1e30;
    t = 128;
    for (int c = 128; c <= 512; c += 32) {
      const int rounds_b = (T.n_bonds + c - 1) / c, rounds_d = (3 * T.n_blocks + c - 1) / c;
      const double cost = 3.0 * rounds_b + 1.0 * rounds_d;  // relative phase weights
      if (cost <= best) { best = cost; t = c; }
    }
  }
  t = (t + 31) / 32 * 32;
  if (t < 32) t = 32;
  if (t > 512) t = 512;
  return t;
}

// Cluster size of the generic kernels: lattices with many bonds per thread are spread over a thread-block cluster
// when the batch alone cannot fill the GPU.  DFX_CLUSTER=1|2|4|8|16 overrides.
int pick_cluster(const DevTopo& T, int batch, int sm_count) {
  int cl = 1;
  if (const char* e = std::getenv("DFX_CLUSTER")) {
    cl = std::atoi(e);
  } else {
    int by_size = 1;
    while (by_size < kMaxCluster && (long long)by_size * 1024 < T.n_bonds) by_size *= 2;
    int by_batch = 1;
    while (by_batch * 2 * (long long)batch <= sm_count && by_batch < kMaxCluster) by_batch *= 2;
    cl = by_size < by_batch ? by_size : by_batch;
  }
  if (cl < 1) cl = 1;
  if (cl > kMaxCluster) cl = kMaxCluster;
  while (cl & (cl - 1)) cl &= cl - 1;  // power of two
  return cl;
}

// Multi-CTA plan of the generic kernels: mode 0 = one CTA per design, 1 = thread-block cluster, 2 = group of
// co-resident CTAs with a software barrier (cooperative launch).  DFX_GROUP=<n> forces mode 2 with n CTAs per design.
struct MultiCta { int mode, ncta; };
MultiCta pick_multi_cta(const DevTopo& T, int batch, int sm_count) {
  if (const char* e = std::getenv("DFX_GROUP")) {
    int n = std::atoi(e);
    if (n > kMaxGroup) n = kMaxGroup;
    if ((long long)n * batch > sm_count) n = sm_count / batch;
    if (n > 1) return {2, n};
    return {0, 1};
  }
  if (!std::getenv("DFX_CLUSTER") && T.n_bonds >= kGroupMinBonds && 8LL * batch <= sm_count) {
    // a lattice this large keeps more SMs busy than a cluster can span; for a small batch of them a group of
    // SMs / batch CTAs also beats 16-CTA clusters, of which only a few fit the GPCs at a time (8 lattices of 100 x 100:
    // 2.1 lattices/s with groups of 18 against 1.26 with clusters of 16, profiles/r02_lattice_timing.jsonl)
    int n = kGroupDefault;
    if ((long long)n * batch > sm_count) n = sm_count / batch;
    if (n >= 8) return {2, n};
  }
  const int cl = pick_cluster(T, batch, sm_count);
  return {cl > 1 ? 1 : 0, cl};
}

template <class Kernel, class Args>
cudaError_t launch_group(Kernel kernel, int batch, int ncta, size_t smem, cudaStream_t stream, Args& args, double* scratch,
                         long long scratch_per_design) {
  // zero every design's barrier counter (first 8 bytes of its scratch slice)
  cudaError_t e = cudaMemset2DAsync(scratch, (size_t)scratch_per_design * sizeof(double), 0, sizeof(unsigned long long), batch, stream);
  if (e != cudaSuccess) return e;
  void* params[] = {(void*)&args};
  return cudaLaunchCooperativeKernel((const void*)kernel, dim3((unsigned)batch * ncta), dim3(512), params, smem, stream);
}

template <class Kernel, class Args>
cudaError_t launch_cluster(Kernel kernel, int batch, int cluster, int threads, size_t smem, cudaStream_t stream, const Args& args) {
  if (cluster > 8) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)batch * cluster, 1, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args);
}

// launch plan of the fast adjoint kernel (dfx_adjoint2.cuh); ok=false -> use the generic kernel
struct FastPlan { bool ok; int threads, nt, ns, cols_per_warp; size_t smem; long long scratch; };

FastPlan plan_fast_adjoint(const DevTopo& T) {
  FastPlan f = {};
  const char* mode = getenv("DFX_ADJOINT_KERNEL");  // "generic" | "notmem" | unset (fast + TMEM)
  if (mode && !strcmp(mode, "generic")) return f;
  if (T.bond_energy == DFX_BOND_SPRING || T.contact == DFX_CONTACT_DISTANCE) return f;  // generic kernels only
  int t = T.n_blocks > (T.n_bonds + 1) / 2 ? T.n_blocks : (T.n_bonds + 1) / 2;
  t = t <= 384 ? 384 : 512;  // the CTA size is a compile-time constant of the kernel (addresses become immediates)
  if (const char* e = getenv("DFX_ADJOINT_THREADS")) { if (atoi(e) == 512) t = 512; }  // experiment: 16 warps at <= 128 registers
  if (T.n_npb > 4 || T.n_blocks > t || T.n_bonds > 2 * t) return f;
  const int nwarp = t / 32, groups = (nwarp + 3) / 4;
  f.cols_per_warp = (512 / groups) & ~1;
  f.nt = f.cols_per_warp >= 168 ? 84 : (f.cols_per_warp >= 128 ? 64 : 0);
  if (mode && !strcmp(mode, "notmem")) f.nt = 0;
  const size_t cap = kSmemBytes - 1024;
  auto fixed_bytes = [&](long long nbs, long long nds) {
    return (size_t)(40 + 8 * nbs + 14 * nds + (2 * NSCAL + 5 * NSCAL * SCW) + 32) * 8 + (size_t)((T.n_nodes + 1) & ~1) * 4;
  };
  size_t fixed = fixed_bytes(T.n_blocks, T.n_bonds);
  if (fixed + 8 * (size_t)t > cap) return f;
  int ns = (int)((cap - fixed) / (8 * (size_t)t));
  if (ns > S_NCONST - f.nt) ns = S_NCONST - f.nt;
  if (f.nt == 84 && t == 384) {
    // specialised variant: all constants in shared memory (32 slots) and compile-time array strides (T, 2T)
    const size_t fixed_s = fixed_bytes(t, 2 * t);
    if (fixed_s + 32 * 8 * (size_t)t <= cap) { fixed = fixed_s; ns = 32; }
    else if (ns == 32) ns = 31;  // cannot use the specialised variant: keep the run-time one
  }
  f.ns = ns;
  f.smem = fixed + (size_t)ns * t * 8;
  const int n_over = S_NCONST - f.nt - ns > 0 ? S_NCONST - f.nt - ns : 0;  // constants that spill to global
  f.scratch = (long long)(n_over + (NQA + 1) * NE) * t + 32;
  f.threads = t;
  f.ok = true;
  return f;
}

// The 24-warp adjoint kernel (dfx_adjoint3.cuh) serves lattices of the cfg1 / cfg3 size with the common vocabulary;
// DFX_ADJOINT_KERNEL=v2 keeps them on dfx_adjoint2.cuh (A/B comparisons).
struct Fast3Plan { bool ok; int npb, damp; bool contact; };

Fast3Plan plan_adjoint3(const DevTopo& T, const DfxParams& p, int n_cons_units) {
  Fast3Plan f = {};
  const char* mode = getenv("DFX_ADJOINT_KERNEL");
  if (mode && (!strcmp(mode, "generic") || !strcmp(mode, "notmem") || !strcmp(mode, "v2"))) return f;
  if (getenv("DFX_ADJOINT_THREADS")) return f;
  if (T.bond_energy != DFX_BOND_LIGAMENT || T.load_kind != DFX_LOAD_NONE || T.contact == DFX_CONTACT_DISTANCE) return f;
  if (p.k_per_bond[0] || p.k_per_bond[1] || p.k_per_bond[2]) return f;
  if (T.n_blocks > k3::TU || T.n_bonds > k3::TT || T.n_blocks < 1 || T.n_bonds < 1) return f;
  if (n_cons_units > k3::NCU) return f;  // the drive vectors of the constrained units are cached in shared memory
  const bool has_damp = T.n_damped > 0 && p.damping.ptr != nullptr;
  f.npb = T.n_npb; f.contact = T.contact != 0; f.damp = !has_damp ? 0 : (p.damping_per_dof ? 2 : 1);
  f.ok = adjoint3_supported(f.npb, f.contact, f.damp);
  return f;
}

// launch plan of the fast forward kernel (dfx_forward2.cuh)
struct FastFwdPlan { bool ok; int threads, nt, ns, cols_per_warp, cols_alloc, ctas; size_t smem; long long scratch; };

FastFwdPlan plan_fast_forward(const DevTopo& T) {
  FastFwdPlan f = {};
  const char* mode = getenv("DFX_FORWARD_KERNEL");  // "generic" | "notmem" | unset (fast + TMEM)
  if (mode && !strcmp(mode, "generic")) return f;
  if (T.bond_energy == DFX_BOND_SPRING || T.contact == DFX_CONTACT_DISTANCE) return f;  // generic kernels only
  int t = T.n_blocks > (T.n_bonds + 1) / 2 ? T.n_blocks : (T.n_bonds + 1) / 2;
  t = t <= 384 ? 384 : 512;
  if (T.n_npb > 4 || T.n_blocks > t || T.n_bonds > 2 * t) return f;
  f.threads = t;
  f.ctas = t == 384 ? 2 : 1;
  f.nt = 42;
  f.cols_per_warp = 2 * f.nt;
  f.cols_alloc = t == 384 ? 256 : 512;  // 3 (resp. 4) warps share a lane quarter
  if (mode && !strcmp(mode, "notmem")) { f.nt = 0; f.cols_alloc = 0; }
  const size_t fixed = (size_t)(40 + 5LL * t + 8LL * t + 4) * 8 + (size_t)((T.n_nodes + 1) & ~1) * 4;
  const size_t cap = kSmemBytes / f.ctas - 1024;
  if (fixed + 8 * (size_t)t > cap) return f;
  int ns = (int)((cap - fixed) / (8 * (size_t)t));
  if (ns > F_NSLOT - f.nt) ns = F_NSLOT - f.nt;
  f.ns = ns;
  f.smem = fixed + (size_t)ns * t * 8;
  const int n_over = F_NSLOT - f.nt - ns > 0 ? F_NSLOT - f.nt - ns : 0;
  f.scratch = (long long)n_over * t + 32;
  f.ok = true;
  return f;
}

int check_params(const DevTopo& T, const DfxParams* p) {
  if (!p) return fail(DFX_ERR_INVALID, "params is NULL");
  if (!p->centroid_node_vectors.ptr || !p->reference_vector.ptr || !p->k_stretch.ptr || !p->k_shear.ptr || !p->k_rot.ptr ||
      !p->inertia.ptr)
    return fail(DFX_ERR_INVALID, "a required parameter leaf is NULL");
  if (T.contact && !p->contact.ptr) return fail(DFX_ERR_INVALID, "topology has contact but params.contact is NULL");
  if (T.contact == DFX_CONTACT_DISTANCE && !p->block_centroids.ptr)
    return fail(DFX_ERR_INVALID, "distance-based contact needs params.block_centroids");
  if (T.n_drive_params > 0 && !p->drive.ptr) return fail(DFX_ERR_INVALID, "drive signal needs params.drive");
  return DFX_OK;
}

}  // namespace

extern "C" {

const char* dfx_last_error(void) { return g_err; }
const char* dfx_version(void) { return "difflexmm_b200 libdfx 0.1 (sm_100a)"; }
int dfx_drive_n_params(int kind) { return n_drive_params_of(kind); }
int dfx_topology_n_free(const DfxTopology* t) { return t ? t->dev.n_free : -1; }

int dfx_topology_create(const DfxTopologyDesc* d, int device, DfxTopology** out) {
  if (device >= 0) keep_pool_memory(device);
  if (!d || !out) return fail(DFX_ERR_INVALID, "NULL argument");
  if (d->n_blocks <= 0 || d->n_npb < 2 || d->n_bonds < 0) return fail(DFX_ERR_INVALID, "bad sizes");
  if (d->bond_energy != DFX_BOND_LIGAMENT && d->bond_energy != DFX_BOND_LINEARIZED && d->bond_energy != DFX_BOND_SPRING)
    return fail(DFX_ERR_UNSUPPORTED, "unknown bond energy %d", d->bond_energy);
  if (d->drive_kind < DFX_DRIVE_ZERO || d->drive_kind > DFX_DRIVE_TABLE)
    return fail(DFX_ERR_UNSUPPORTED, "unknown drive kind %d", d->drive_kind);
  if (d->load_kind < DFX_LOAD_NONE || d->load_kind > DFX_LOAD_SECH2)
    return fail(DFX_ERR_UNSUPPORTED, "unknown load kind %d", d->load_kind);
  const int n_dof = 3 * d->n_blocks, n_nodes = d->n_blocks * d->n_npb;
  std::vector<int> cons_slot(n_dof, -1), free_of_dof(n_dof, -1), damp_slot(n_dof, -1), free_dofs, node_used(n_nodes, 0);
  std::vector<double> v0(d->n_constrained > 0 ? d->n_constrained : 1, 0.0), v1(v0), load_mul(n_dof, 0.0);
  for (int c = 0; c < d->n_constrained; ++c) {
    const int dof = d->constrained_dofs[c];
    if (dof < 0 || dof >= n_dof) return fail(DFX_ERR_INVALID, "constrained DOF %d out of range", dof);
    if (cons_slot[dof] >= 0) return fail(DFX_ERR_INVALID, "constrained DOF %d listed twice", dof);
    cons_slot[dof] = c;
    if (d->drive_vec0) v0[c] = d->drive_vec0[c];
    if (d->drive_vec1) v1[c] = d->drive_vec1[c];
  }
  for (int i = 0; i < n_dof; ++i)
    if (cons_slot[i] < 0) { free_of_dof[i] = (int)free_dofs.size(); free_dofs.push_back(i); }
  for (int k = 0; k < d->n_damped; ++k) {
    const int blk = d->damped_blocks[k];
    if (blk < 0 || blk >= d->n_blocks) return fail(DFX_ERR_INVALID, "damped block %d out of range", blk);
    for (int j = 0; j < 3; ++j) damp_slot[3 * blk + j] = 3 * k + j;
  }
  if (d->load_kind != DFX_LOAD_NONE)
    for (int l = 0; l < d->n_loaded; ++l) {
      const int dof = d->loaded_dofs[l];
      if (dof < 0 || dof >= n_dof) return fail(DFX_ERR_INVALID, "loaded DOF %d out of range", dof);
      load_mul[dof] = d->load_vec ? d->load_vec[l] : 1.0;
    }
  std::vector<int2> bn(d->n_bonds), bb(d->n_bonds);
  std::vector<int> node_bond(n_nodes, -1);
  for (int b = 0; b < d->n_bonds; ++b) {
    const int na = d->bond_nodes[2 * b], nb = d->bond_nodes[2 * b + 1];
    if (na < 0 || na >= n_nodes || nb < 0 || nb >= n_nodes) return fail(DFX_ERR_INVALID, "bond %d refers to a node out of range", b);
    if (node_used[na]++ || node_used[nb]++)
      return fail(DFX_ERR_UNSUPPORTED,
                  "a polygon vertex belongs to more than one bond (bond %d); the slot-based force assembly "
                  "requires at most one bond per vertex, as in every geometry class of the reference", b);
    bn[b] = make_int2(na, nb);
    bb[b] = make_int2(na / d->n_npb, nb / d->n_npb);
    node_bond[na] = 2 * b;
    node_bond[nb] = 2 * b + 1;
  }
  int cur = 0;
  CUDA_TRY(cudaGetDevice(&cur));
  CUDA_TRY(cudaSetDevice(device));
  DfxTopology* t = new (std::nothrow) DfxTopology();
  if (!t) return fail(DFX_ERR_INVALID, "out of host memory");
  t->device = device;
  DevTopo& D = t->dev;
  std::memset(&D, 0, sizeof(D));
  D.n_blocks = d->n_blocks; D.n_npb = d->n_npb; D.n_bonds = d->n_bonds; D.n_nodes = n_nodes; D.n_dof = n_dof;
  D.n_free = (int)free_dofs.size(); D.n_cons = d->n_constrained;
  D.bond_energy = d->bond_energy; D.contact = d->contact; D.drive_kind = d->drive_kind;
  D.load_kind = d->load_kind; D.n_drive_params = n_drive_params_of(d->drive_kind); D.n_damped = d->n_damped;
  for (int i = 0; i < DFX_MAX_LOAD_CONSTS; ++i) D.load_consts[i] = d->load_consts[i];
  int2 *dbn, *dbb; int *dfo, *dcs, *dds, *dfd, *dnb; double *dv0, *dv1, *dlm, *dtt = nullptr, *dtv = nullptr;
  std::vector<double> tab_t, tab_v;
  if (d->drive_kind == DFX_DRIVE_TABLE) {
    if (d->drive_table_len < 1 || !d->drive_table_t || !d->drive_table_v) { delete t; cudaSetDevice(cur); return fail(DFX_ERR_INVALID, "tabulated drive needs drive_table_t / drive_table_v"); }
    tab_t.assign(d->drive_table_t, d->drive_table_t + d->drive_table_len);
    tab_v.assign(d->drive_table_v, d->drive_table_v + d->drive_table_len);
  }
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = upload(bn, &dbn);
  if (e == cudaSuccess) e = upload(bb, &dbb);
  if (e == cudaSuccess) e = upload(free_of_dof, &dfo);
  if (e == cudaSuccess) e = upload(cons_slot, &dcs);
  if (e == cudaSuccess) e = upload(damp_slot, &dds);
  if (e == cudaSuccess) e = upload(free_dofs, &dfd);
  if (e == cudaSuccess) e = upload(v0, &dv0);
  if (e == cudaSuccess) e = upload(v1, &dv1);
  if (e == cudaSuccess) e = upload(load_mul, &dlm);
  if (e == cudaSuccess) e = upload(node_bond, &dnb);
  if (e == cudaSuccess) e = upload(tab_t, &dtt);
  if (e == cudaSuccess) e = upload(tab_v, &dtv);
  if (e != cudaSuccess) { delete t; cudaSetDevice(cur); return fail(DFX_ERR_CUDA, "topology upload failed: %s", cudaGetErrorString(e)); }
  D.bond_nodes = dbn; D.bond_blocks = dbb; D.free_of_dof = dfo; D.cons_slot = dcs; D.damp_slot = dds; D.free_dofs = dfd;
  D.drive_vec0 = dv0; D.drive_vec1 = dv1; D.load_mul = dlm;
  t->node_bond = dnb;
  t->n_cons_units = 0;
  for (int b = 0; b < D.n_blocks; ++b)
    if (cons_slot[3 * b] >= 0 || cons_slot[3 * b + 1] >= 0 || cons_slot[3 * b + 2] >= 0) t->n_cons_units++;
  D.table.t = dtt; D.table.v = dtv; D.table.n = (int)tab_t.size();
  t->allocs = {dbn, dbb, dfo, dcs, dds, dfd, dv0, dv1, dlm, dnb, dtt, dtv};
  cudaDeviceGetAttribute(&t->sm_count, cudaDevAttrMultiProcessorCount, device);
  cudaFuncSetAttribute(forward_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  cudaFuncSetAttribute(forward2_kernel<42, 384, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes / 2 - 1024);
  cudaFuncSetAttribute(forward2_kernel<0, 384, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes / 2 - 1024);
  cudaFuncSetAttribute(forward2_kernel<42, 512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes - 1024);
  cudaFuncSetAttribute(forward2_kernel<0, 512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes - 1024);
  cudaFuncSetAttribute(forward2_kernel<42, 384, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(forward2_kernel<42, 384, 2, F_NSLOT - 42>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes / 2 - 1024);
  cudaFuncSetAttribute(forward2_kernel<42, 384, 2, F_NSLOT - 42>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(forward2_kernel<42, 512, 1, F_NSLOT - 42>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes - 1024);
  cudaFuncSetAttribute(forward2_kernel<0, 384, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(adjoint_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  cudaFuncSetAttribute(adjoint2_kernel<84, 32, 384>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes - 1024);
  cudaFuncSetAttribute(adjoint2_kernel<84, -1, 384>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes - 1024);
  cudaFuncSetAttribute(adjoint2_kernel<64, -1, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes - 1024);
  cudaFuncSetAttribute(adjoint2_kernel<0, -1, 384>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes - 1024);
  cudaFuncSetAttribute(adjoint2_kernel<0, -1, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes - 1024);
  cudaSetDevice(cur);
  *out = t;
  return DFX_OK;
}

void dfx_topology_destroy(DfxTopology* t) {
  if (!t) return;
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(t->device);
  for (void* p : t->allocs) if (p) cudaFree(p);
  cudaSetDevice(cur);
  delete t;
}

size_t dfx_forward_workspace_bytes(const DfxTopology* t, int batch) {
  if (!t) return 0;
  long long sz[FA_COUNT], off[FA_COUNT], g;
  size_t smem;
  FastFwdPlan f = plan_fast_forward(t->dev);
  forward_sizes(t->dev, sz);
  plan(sz, off, &smem, &g, f.ok ? 1 : pick_multi_cta(t->dev, batch, t->sm_count).ncta);
  const long long fast = f.ok ? (long long)F_NSLOT * f.threads + 32 : 0;
  if (fast > g) g = fast;
  return (size_t)g * sizeof(double) * (size_t)batch;
}

size_t dfx_adjoint_workspace_bytes(const DfxTopology* t, int batch) {
  if (!t) return 0;
  // worst case over the leaf forms (per-bond stiffnesses, per-DOF damping)
  DfxParams p;
  std::memset(&p, 0, sizeof(p));
  p.k_per_bond[0] = p.k_per_bond[1] = p.k_per_bond[2] = 1;
  p.damping_per_dof = 1;
  QuadLayout q = quad_layout(t->dev, p);
  long long sz[AA_COUNT], off[AA_COUNT], g;
  size_t smem;
  FastPlan f = plan_fast_adjoint(t->dev);
  const int cluster = f.ok ? 1 : pick_multi_cta(t->dev, batch, t->sm_count).ncta;
  adjoint_sizes(t->dev, q, sz, cluster);
  plan(sz, off, &smem, &g, cluster);
  // worst case of the two kernels (the fast one assumes S_TOTAL slots with no TMEM)
  long long fast = f.ok ? (long long)(S_NCONST - f.ns + (NQA + 1) * NE) * f.threads + 32 : 0;
  if (fast > g) g = fast;
  if (f.ok && adjoint3_scratch_doubles() > g) g = adjoint3_scratch_doubles();
  const int units = f.ok && batch < kScratchSlots ? kScratchSlots : batch;  // the fast kernel indexes scratch by SM id
  return (size_t)g * sizeof(double) * (size_t)units;
}

int dfx_forward(const DfxTopology* t, const DfxParams* params, int batch, const double* y0, int64_t y0_bstride,
                const double* ts, int64_t ts_bstride, int n_t, double rtol, double atol, const DfxOptions* opt,
                double* ys, DfxStats* stats, void* workspace, size_t workspace_bytes, void* stream_) {
  if (!t || !y0 || !ts || !ys) return fail(DFX_ERR_INVALID, "NULL argument");
  if (batch <= 0 || n_t < 1) return fail(DFX_ERR_INVALID, "batch and n_t must be positive");
  if (int rc = check_params(t->dev, params)) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  FwdArgs a;
  std::memset(&a, 0, sizeof(a));
  a.topo = t->dev; a.p = *params; a.tab = make_tableau();
  long long sz[FA_COUNT], g;
  size_t smem;
  const FastFwdPlan fp = plan_fast_forward(t->dev);
  const MultiCta mc = fp.ok ? MultiCta{0, 1} : pick_multi_cta(t->dev, batch, t->sm_count);
  const int cluster = mc.ncta;
  forward_sizes(t->dev, sz);
  plan(sz, a.place.off, &smem, &g, cluster);
  a.y0 = y0; a.y0_bstride = y0_bstride; a.ts = ts; a.ts_bstride = ts_bstride; a.n_t = n_t;
  a.rtol = rtol; a.atol = atol;
  a.init_step_variant = opt ? opt->init_step_variant : 0;
  a.max_steps = (opt && opt->max_steps > 0) ? opt->max_steps : (1LL << 40);
  a.order = opt ? opt->design_order : nullptr;
  a.ys = ys; a.stats = stats;
  if (fp.ok) g = fp.scratch;
  a.scratch_per_design = g;
  bool own_ws = false;
  if (g > 0) {
    const size_t need = (size_t)g * sizeof(double) * (size_t)batch;
    if (workspace) {
      if (workspace_bytes < need) return fail(DFX_ERR_INVALID, "workspace too small: %zu < %zu", workspace_bytes, need);
      a.scratch = (double*)workspace;
    } else {
      CUDA_TRY(cudaMallocAsync((void**)&a.scratch, need, stream));
      own_ws = true;
    }
  }
  if (fp.ok) {
    Fwd2Args A2;
    A2.a = a;
    A2.node_bond = t->node_bond;
    A2.tp_scratch_per_design = fp.scratch;
    A2.ns_slots = fp.ns;
    A2.tmem_cols_per_warp = fp.cols_per_warp;
    A2.tmem_cols_alloc = fp.cols_alloc;
    if (fp.threads == 384 && fp.nt == 42 && fp.ns == F_NSLOT - 42) forward2_kernel<42, 384, 2, F_NSLOT - 42><<<batch, 384, fp.smem, stream>>>(A2);
    else if (fp.threads == 384 && fp.nt == 42) forward2_kernel<42, 384, 2><<<batch, 384, fp.smem, stream>>>(A2);
    else if (fp.threads == 384) forward2_kernel<0, 384, 2><<<batch, 384, fp.smem, stream>>>(A2);
    else if (fp.nt == 42 && fp.ns == F_NSLOT - 42) forward2_kernel<42, 512, 1, F_NSLOT - 42><<<batch, 512, fp.smem, stream>>>(A2);
    else if (fp.nt == 42) forward2_kernel<42, 512, 1><<<batch, 512, fp.smem, stream>>>(A2);
    else forward2_kernel<0, 512, 1><<<batch, 512, fp.smem, stream>>>(A2);
  } else {
    const int threads = pick_threads(t->dev, opt ? opt->threads : 0);
    if (cluster > 1) {
      a.group = cluster;
      cudaError_t le = mc.mode == 2 ? launch_group(forward_kernel<2>, batch, cluster, smem, stream, a, a.scratch, a.scratch_per_design)
                                    : launch_cluster(forward_kernel<1>, batch, cluster, 512, smem, stream, a);
      if (le != cudaSuccess) { if (own_ws) cudaFreeAsync(a.scratch, stream); return fail(DFX_ERR_CUDA, "forward_kernel multi-CTA launch (%d CTAs per design) failed: %s", cluster, cudaGetErrorString(le)); }
    } else {
      forward_kernel<0><<<batch, threads, smem, stream>>>(a);
    }
  }
  cudaError_t e = cudaGetLastError();
  if (own_ws) cudaFreeAsync(a.scratch, stream);
  if (e != cudaSuccess) return fail(DFX_ERR_CUDA, "forward_kernel launch failed: %s", cudaGetErrorString(e));
  return DFX_OK;
}

}  // extern "C"

namespace {
__global__ void objective_cotangent_kernel(AdjArgs a, double* g) {
  // materialised cotangent of the device objective for the generic adjoint kernels: g[b][i][:]
  const int b = blockIdx.y, i = blockIdx.x, nf = a.topo.n_free;
  double* gi = g + ((long long)b * a.n_t + i) * 2 * nf;
  for (int k = threadIdx.x; k < 2 * nf; k += blockDim.x) gi[k] = 0.0;
  __syncthreads();
  for (int k = threadIdx.x; k < 2 * a.obj_n; k += blockDim.x) {
    const bool is_v = k >= a.obj_n;
    const int f = a.obj_ids[is_v ? k - a.obj_n : k];
    gi[(is_v ? nf : 0) + f] = objective_cotangent(a, b, i, f, is_v);
  }
}

__global__ void objective_kernel(DevTopo T, DfxLeaf inertia, const double* ys, int n_t, DfxObjective obj, double* value,
                                 double* inertia_bar, double* arm_bar) {
  // one CTA per design: J (weights not applied) and its explicit derivatives w.r.t. the inertia and the arms
  __shared__ double red[40];
  const int b = blockIdx.x, nf = T.n_free;
  const double* m = inertia.ptr + (long long)b * inertia.bstride;
  const int* ids = obj.target_free_ids;
  if (inertia_bar) for (int k = threadIdx.x; k < nf; k += blockDim.x) inertia_bar[(long long)b * nf + k] = 0.0;
  __syncthreads();
  double acc = 0.0;
  if (obj.kind == DFX_OBJ_KINETIC) {
    for (int k = threadIdx.x; k < obj.n_target; k += blockDim.x) {
      const int f = ids[k];
      double s2 = 0.0;
      for (int i = 0; i < n_t; ++i) { const double v = ys[((long long)b * n_t + i) * 2 * nf + nf + f]; s2 = fma(v, v, s2); }
      acc += 0.5 * m[f] * s2;
      if (inertia_bar) inertia_bar[(long long)b * nf + f] = 0.5 * s2;
    }
  } else {
    for (int kb = threadIdx.x; kb < obj.n_target / 3; kb += blockDim.x) {
      const int fx = ids[3 * kb], fy = ids[3 * kb + 1], ft = ids[3 * kb + 2];
      const double* arm = obj.arm + (long long)b * obj.arm_bstride + 2 * kb;
      double s_pxvy = 0.0, s_pyvx = 0.0, s_vy = 0.0, s_vx = 0.0, s_w = 0.0;
      for (int i = 0; i < n_t; ++i) {
        const double* y = ys + ((long long)b * n_t + i) * 2 * nf;
        const double px = arm[0] + y[fx], py = arm[1] + y[fy], vx = y[nf + fx], vy = y[nf + fy];
        s_pxvy = fma(px, vy, s_pxvy); s_pyvx = fma(py, vx, s_pyvx); s_vx += vx; s_vy += vy; s_w += y[nf + ft];
      }
      acc += m[fy] * s_pxvy - m[fx] * s_pyvx + m[ft] * s_w;
      if (inertia_bar) {
        inertia_bar[(long long)b * nf + fx] = -s_pyvx;
        inertia_bar[(long long)b * nf + fy] = s_pxvy;
        inertia_bar[(long long)b * nf + ft] = s_w;
      }
      if (arm_bar) {
        arm_bar[((long long)b * (obj.n_target / 3) + kb) * 2] = m[fy] * s_vy;
        arm_bar[((long long)b * (obj.n_target / 3) + kb) * 2 + 1] = -m[fx] * s_vx;
      }
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) value[b] = acc;
}

int check_objective(const DfxObjective* obj) {
  if (obj->kind != DFX_OBJ_KINETIC && obj->kind != DFX_OBJ_ANGULAR) return fail(DFX_ERR_UNSUPPORTED, "unknown objective kind %d", obj->kind);
  if (obj->n_target < 0 || (obj->n_target > 0 && !obj->target_free_ids)) return fail(DFX_ERR_INVALID, "bad objective");
  if (obj->kind == DFX_OBJ_ANGULAR && (obj->n_target % 3 != 0 || !obj->arm))
    return fail(DFX_ERR_INVALID, "the angular-momentum objective needs (x, y, theta) triples of target DOFs and the arms");
  return DFX_OK;
}

int adjoint_impl(const DfxTopology* t, const DfxParams* params, int batch, const double* ys, const double* ts,
                 int64_t ts_bstride, int n_t, const double* g_, const DfxObjective* obj, double rtol, double atol,
                 int64_t aug_size, const DfxOptions* opt, double* y0_bar, double* ts_bar, const DfxParamGrads* grads,
                 DfxStats* stats, void* workspace, size_t workspace_bytes, void* stream_) {
  if (!t || !ys || !ts || (!g_ && !obj)) return fail(DFX_ERR_INVALID, "NULL argument");
  if (obj) { if (int rc = check_objective(obj)) return rc; }
  if (batch <= 0 || n_t < 1) return fail(DFX_ERR_INVALID, "batch and n_t must be positive");
  if (int rc = check_params(t->dev, params)) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  const DevTopo& T = t->dev;
  AdjArgs a;
  std::memset(&a, 0, sizeof(a));
  a.topo = T; a.p = *params; a.tab = make_tableau();
  QuadLayout q = quad_layout(T, *params);
  a.qo_cnv = q.qo_cnv; a.qo_ref = q.qo_ref; a.qo_ks = q.qo_ks; a.qo_ksh = q.qo_ksh; a.qo_kr = q.qo_kr;
  a.qo_damp = q.qo_damp; a.qo_inertia = q.qo_inertia; a.qo_cen = q.qo_cen; a.nq = q.nq;
  long long sz[AA_COUNT], g;
  size_t smem;
  const FastPlan fp = plan_fast_adjoint(T);
  const Fast3Plan f3 = fp.ok ? plan_adjoint3(T, *params, t->n_cons_units) : Fast3Plan{};
  const MultiCta mc = fp.ok ? MultiCta{0, 1} : pick_multi_cta(T, batch, t->sm_count);
  const int cluster = mc.ncta;
  adjoint_sizes(T, q, sz, cluster);
  plan(sz, a.place.off, &smem, &g, cluster);
  a.ys = ys; a.ts = ts; a.ts_bstride = ts_bstride; a.n_t = n_t; a.g = g_;
  if (obj) {
    a.obj_kind = obj->kind; a.obj_ids = obj->target_free_ids; a.obj_n = obj->n_target; a.obj_w = obj->weights;
    a.obj_arm = obj->arm; a.obj_arm_bstride = obj->arm_bstride;
  }
  a.rtol = rtol; a.atol = atol;
  if (aug_size <= 0) {
    // count the leaves listed in DfxParams: y, y_bar, t0_bar, then every leaf
    long long n = 4LL * T.n_free + 1;
    n += 2LL * T.n_nodes + 2LL * T.n_bonds + T.n_free;
    for (int k = 0; k < 3; ++k) n += params->k_per_bond[k] ? T.n_bonds : 1;
    if (T.n_damped > 0 && params->damping.ptr) n += params->damping_per_dof ? 3LL * T.n_damped : 1;
    if (T.contact) n += 3;
    if (T.contact == DFX_CONTACT_DISTANCE) n += 2LL * T.n_blocks;
    n += T.n_drive_params;
    aug_size = n;
  }
  a.aug_size = aug_size;
  a.init_step_variant = opt ? opt->init_step_variant : 0;
  a.max_steps = (opt && opt->max_steps > 0) ? opt->max_steps : (1LL << 40);
  a.order = opt ? opt->design_order : nullptr;
  a.y0_bar = y0_bar; a.ts_bar = ts_bar;
  if (grads) a.grads = *grads;
  a.stats = stats;
  if (fp.ok) g = fp.scratch;
  if (f3.ok) g = adjoint3_scratch_doubles();
  a.scratch_per_design = g;
  bool own_ws = false;
  // fast kernels: scratch indexed by SM id when that is smaller than one slice per design (always for the 24-warp kernel)
  const int scratch_slots = f3.ok ? kScratchSlots : ((fp.ok && batch > kScratchSlots) ? kScratchSlots : 0);
  if (g > 0) {
    const size_t need = (size_t)g * sizeof(double) * (size_t)(scratch_slots ? scratch_slots : batch);
    if (workspace) {
      if (workspace_bytes < need) return fail(DFX_ERR_INVALID, "workspace too small: %zu < %zu", workspace_bytes, need);
      a.scratch = (double*)workspace;
    } else {
      CUDA_TRY(cudaMallocAsync((void**)&a.scratch, need, stream));
      own_ws = true;
    }
  }
  if (f3.ok) {
    Adj3Args A3;
    A3.a = a;
    A3.node_bond = t->node_bond;
    A3.scratch_per_slot = g;
    A3.scratch_slots = scratch_slots;
    cudaError_t le = launch_adjoint3(A3, f3.npb, f3.contact, f3.damp, batch, stream);
    if (le != cudaSuccess) { if (own_ws) cudaFreeAsync(a.scratch, stream); return fail(DFX_ERR_CUDA, "adjoint3_kernel launch failed: %s", cudaGetErrorString(le)); }
  } else if (fp.ok) {
    Adj2Args A2;
    A2.a = a;
    A2.node_bond = t->node_bond;
    A2.tp_scratch_per_design = fp.scratch;
    A2.ns_slots = fp.ns;
    A2.scratch_slots = scratch_slots;
    A2.tmem_cols_per_warp = fp.cols_per_warp;
    if (fp.nt == 84 && fp.ns == 32) adjoint2_kernel<84, 32, 384><<<batch, 384, fp.smem, stream>>>(A2);
    else if (fp.nt == 84) adjoint2_kernel<84, -1, 384><<<batch, 384, fp.smem, stream>>>(A2);
    else if (fp.nt == 64) adjoint2_kernel<64, -1, 512><<<batch, 512, fp.smem, stream>>>(A2);
    else if (fp.threads == 384) adjoint2_kernel<0, -1, 384><<<batch, 384, fp.smem, stream>>>(A2);
    else adjoint2_kernel<0, -1, 512><<<batch, 512, fp.smem, stream>>>(A2);
  } else {
    // generic kernel: it reads a materialised cotangent
    double* gtmp = nullptr;
    if (!g_) {
      CUDA_TRY(cudaMallocAsync((void**)&gtmp, (size_t)batch * n_t * 2 * T.n_free * sizeof(double), stream));
      objective_cotangent_kernel<<<dim3(n_t, batch), 256, 0, stream>>>(a, gtmp);
      a.g = gtmp;
    }
    const int threads = pick_threads(T, opt ? opt->threads : 0);
    cudaError_t le = cudaSuccess;
    a.group = cluster;
    const size_t smem_multi = (size_t)(kRedDoubles + kScalDoubles) * sizeof(double);
    if (mc.mode == 2) le = launch_group(adjoint_kernel<2>, batch, cluster, smem_multi, stream, a, a.scratch, a.scratch_per_design);
    else if (cluster > 1) le = launch_cluster(adjoint_kernel<1>, batch, cluster, 512, smem_multi, stream, a);
    else adjoint_kernel<0><<<batch, threads, smem, stream>>>(a);
    if (gtmp) cudaFreeAsync(gtmp, stream);
    if (le != cudaSuccess) { if (own_ws) cudaFreeAsync(a.scratch, stream); return fail(DFX_ERR_CUDA, "adjoint_kernel multi-CTA launch (%d CTAs per design) failed: %s", cluster, cudaGetErrorString(le)); }
  }
  cudaError_t e = cudaGetLastError();
  if (own_ws) cudaFreeAsync(a.scratch, stream);
  if (e != cudaSuccess) return fail(DFX_ERR_CUDA, "adjoint_kernel launch failed: %s", cudaGetErrorString(e));
  return DFX_OK;
}
}  // namespace

extern "C" {

int dfx_adjoint(const DfxTopology* t, const DfxParams* params, int batch, const double* ys, const double* ts,
                int64_t ts_bstride, int n_t, const double* g_, double rtol, double atol, int64_t aug_size,
                const DfxOptions* opt, double* y0_bar, double* ts_bar, const DfxParamGrads* grads, DfxStats* stats,
                void* workspace, size_t workspace_bytes, void* stream_) {
  if (!g_) return fail(DFX_ERR_INVALID, "NULL argument");
  return adjoint_impl(t, params, batch, ys, ts, ts_bstride, n_t, g_, nullptr, rtol, atol, aug_size, opt, y0_bar, ts_bar, grads,
                      stats, workspace, workspace_bytes, stream_);
}

int dfx_adjoint_objective(const DfxTopology* t, const DfxParams* params, int batch, const double* ys, const double* ts,
                          int64_t ts_bstride, int n_t, const DfxObjective* obj, double rtol, double atol,
                          int64_t aug_size, const DfxOptions* opt, double* y0_bar, double* ts_bar, const DfxParamGrads* grads,
                          DfxStats* stats, void* workspace, size_t workspace_bytes, void* stream_) {
  if (!obj) return fail(DFX_ERR_INVALID, "NULL argument");
  return adjoint_impl(t, params, batch, ys, ts, ts_bstride, n_t, nullptr, obj, rtol, atol, aug_size, opt, y0_bar, ts_bar, grads,
                      stats, workspace, workspace_bytes, stream_);
}

int dfx_objective(const DfxTopology* t, const DfxParams* params, int batch, const double* ys, int n_t, const DfxObjective* obj,
                  double* value, double* inertia_bar, double* arm_bar, void* stream_) {
  if (!t || !params || !ys || !obj || !value || !params->inertia.ptr) return fail(DFX_ERR_INVALID, "NULL argument");
  if (int rc = check_objective(obj)) return rc;
  objective_kernel<<<batch, 128, 0, (cudaStream_t)stream_>>>(t->dev, params->inertia, ys, n_t, *obj, value, inertia_bar, arm_bar);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(DFX_ERR_CUDA, "objective launch failed: %s", cudaGetErrorString(e));
  return DFX_OK;
}

}  // extern "C"

// ---- design -> parameters (dfx_geometry.cuh) --------------------------------------------------------------------
struct DfxGeometry {
  int device;
  DevGeometry dev;
  std::vector<void*> allocs;
};

extern "C" {

int dfx_geometry_create(const DfxGeometryDesc* d, int device, DfxGeometry** out) {
  if (!d || !out || !d->base_nodes || !d->node_design) return fail(DFX_ERR_INVALID, "NULL argument");
  keep_pool_memory(device);
  if (d->n_blocks <= 0 || d->n_npb < 3 || d->n_npb > kMaxPolygon || d->n_design < 0)
    return fail(DFX_ERR_INVALID, "bad geometry sizes (polygons have 3..%d vertices)", kMaxPolygon);
  const int nn = d->n_blocks * d->n_npb;
  std::vector<int> off(d->n_design + 1, 0), nodes;
  for (int n = 0; n < nn; ++n) {
    const int k = d->node_design[n];
    if (k >= d->n_design) return fail(DFX_ERR_INVALID, "node_design[%d] = %d out of range", n, k);
    if (k >= 0) off[k + 1]++;
  }
  for (int k = 0; k < d->n_design; ++k) off[k + 1] += off[k];
  nodes.resize(off[d->n_design]);
  {
    std::vector<int> fill(off.begin(), off.end() - 1);
    for (int n = 0; n < nn; ++n) if (d->node_design[n] >= 0) nodes[fill[d->node_design[n]]++] = n;
  }
  int cur = 0;
  cudaGetDevice(&cur);
  if (cudaSetDevice(device) != cudaSuccess) return fail(DFX_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  DfxGeometry* g = new (std::nothrow) DfxGeometry();
  if (!g) { cudaSetDevice(cur); return fail(DFX_ERR_INVALID, "out of host memory"); }
  g->device = device;
  std::vector<double> base(d->base_nodes, d->base_nodes + 2 * (size_t)nn);
  std::vector<int> nd(d->node_design, d->node_design + nn);
  double* dbase = nullptr; int *dnd = nullptr, *doff = nullptr, *dnodes = nullptr;
  cudaError_t e = upload(base, &dbase);
  if (e == cudaSuccess) e = upload(nd, &dnd);
  if (e == cudaSuccess) e = upload(off, &doff);
  if (e == cudaSuccess) e = upload(nodes, &dnodes);
  g->allocs = {dbase, dnd, doff, dnodes};
  cudaSetDevice(cur);
  if (e != cudaSuccess) { dfx_geometry_destroy(g); return fail(DFX_ERR_CUDA, "geometry upload failed: %s", cudaGetErrorString(e)); }
  g->dev.n_blocks = d->n_blocks; g->dev.n_npb = d->n_npb; g->dev.n_nodes = nn; g->dev.n_design = d->n_design;
  g->dev.base_nodes = dbase; g->dev.node_design = dnd; g->dev.design_off = doff; g->dev.design_nodes = dnodes;
  *out = g;
  return DFX_OK;
}

void dfx_geometry_destroy(DfxGeometry* g) {
  if (!g) return;
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(g->device);
  for (void* p : g->allocs) if (p) cudaFree(p);
  cudaSetDevice(cur);
  delete g;
}

int dfx_geometry_forward(const DfxGeometry* g, int batch, const double* design, const double* density, int64_t density_bstride,
                         double* cnv, double* centroid_shift, double* inertia, void* stream_) {
  if (!g || !design || !cnv || (inertia && !density)) return fail(DFX_ERR_INVALID, "NULL argument");
  if (batch <= 0) return fail(DFX_ERR_INVALID, "batch must be positive");
  const int threads = 128;
  dim3 grid((g->dev.n_blocks + threads - 1) / threads, batch);
  geometry_forward_kernel<<<grid, threads, 0, (cudaStream_t)stream_>>>(g->dev, design, density, density_bstride, cnv,
                                                                      centroid_shift, inertia);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(DFX_ERR_CUDA, "geometry_forward launch failed: %s", cudaGetErrorString(e));
  return DFX_OK;
}

int dfx_geometry_vjp(const DfxGeometry* g, int batch, const double* design, const double* density, int64_t density_bstride,
                     const double* cnv_bar, const double* centroid_bar, const double* inertia_bar, double* design_bar,
                     double* density_bar, void* stream_) {
  if (!g || !design || !design_bar || (inertia_bar && !density)) return fail(DFX_ERR_INVALID, "NULL argument");
  if (batch <= 0) return fail(DFX_ERR_INVALID, "batch must be positive");
  cudaStream_t stream = (cudaStream_t)stream_;
  double* node_bar = nullptr;
  CUDA_TRY(cudaMallocAsync((void**)&node_bar, (size_t)batch * g->dev.n_nodes * 2 * sizeof(double), stream));
  geometry_vjp_kernel<<<batch, 256, 0, stream>>>(g->dev, design, density, density_bstride, cnv_bar, centroid_bar, inertia_bar,
                                                 node_bar, design_bar, density_bar);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(node_bar, stream);
  if (e != cudaSuccess) return fail(DFX_ERR_CUDA, "geometry_vjp launch failed: %s", cudaGetErrorString(e));
  return DFX_OK;
}

}  // extern "C"

// ---- RotatedSquareGeometry design map (dfx_geometry.cuh) -----------------------------------------------------------
extern "C" {

int dfx_rotated_square_forward(int n1_blocks, int n2_blocks, double half_side, int batch, const double* angle, const double* density,
                               int64_t density_bstride, double* cnv, double* inertia, void* stream_) {
  if (!angle || !cnv || (inertia && !density)) return fail(DFX_ERR_INVALID, "NULL argument");
  if (n1_blocks <= 0 || n2_blocks <= 0 || batch <= 0 || !(half_side > 0.0)) return fail(DFX_ERR_INVALID, "bad rotated-square sizes");
  const int nb = n1_blocks * n2_blocks, threads = 128;
  dim3 grid((nb + threads - 1) / threads, batch);
  rotated_square_forward_kernel<<<grid, threads, 0, (cudaStream_t)stream_>>>(n1_blocks, nb, half_side, angle, density, density_bstride, cnv, inertia);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(DFX_ERR_CUDA, "rotated_square_forward launch failed: %s", cudaGetErrorString(e));
  return DFX_OK;
}

int dfx_rotated_square_vjp(int n1_blocks, int n2_blocks, double half_side, int batch, const double* angle, const double* density,
                           int64_t density_bstride, const double* cnv_bar, const double* inertia_bar, double* angle_bar,
                           double* density_bar, void* stream_) {
  if (!angle || !angle_bar || (inertia_bar && !density)) return fail(DFX_ERR_INVALID, "NULL argument");
  if (n1_blocks <= 0 || n2_blocks <= 0 || batch <= 0 || !(half_side > 0.0)) return fail(DFX_ERR_INVALID, "bad rotated-square sizes");
  rotated_square_vjp_kernel<<<batch, 256, 0, (cudaStream_t)stream_>>>(n1_blocks, n1_blocks * n2_blocks, half_side, angle, density,
                                                                     density_bstride, cnv_bar, inertia_bar, angle_bar, density_bar);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(DFX_ERR_CUDA, "rotated_square_vjp launch failed: %s", cudaGetErrorString(e));
  return DFX_OK;
}

}  // extern "C"

// ---- design constraints (dfx_geometry.cuh) -----------------------------------------------------------------------
struct DfxConstraints {
  int device;       // copies: the handle must stay destroyable after its geometry (finalisers run in any order)
  DevGeometry dev;  // device tables owned by the geometry, which must be alive for dfx_constraints_eval
  int m, m_angle;
  ConstraintRow* rows;         // device
  std::vector<int32_t> cols;   // host, [m][4]
};

extern "C" {

int dfx_constraints_create(const DfxGeometry* g, const DfxConstraintDesc* d, DfxConstraints** out) {
  if (!g || !d || !out) return fail(DFX_ERR_INVALID, "NULL argument");
  if (d->n_bonds < 0 || d->n_boundary < 0 || (d->n_bonds > 0 && !d->bonds) || (d->n_boundary > 0 && !d->boundary_nodes))
    return fail(DFX_ERR_INVALID, "bad constraint tables");
  const int npb = g->dev.n_npb, nn = g->dev.n_nodes;
  std::vector<int> nd(nn);
  int cur = 0;
  cudaGetDevice(&cur);
  if (cudaSetDevice(g->device) != cudaSuccess) return fail(DFX_ERR_CUDA, "cudaSetDevice(%d) failed", g->device);
  cudaError_t e = cudaMemcpy(nd.data(), g->dev.node_design, nn * sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { cudaSetDevice(cur); return fail(DFX_ERR_CUDA, "constraint table download failed: %s", cudaGetErrorString(e)); }
  auto next = [&](int n) { return (n / npb) * npb + (n % npb + 1) % npb; };
  auto prev = [&](int n) { return (n / npb) * npb + (n % npb + npb - 1) % npb; };
  std::vector<ConstraintRow> rows;
  std::vector<int32_t> cols;
  auto add = [&](int p0, int p1, int q0, int q1, double minimum) {
    ConstraintRow R;
    R.p0 = p0; R.p1 = p1; R.q0 = q0; R.q1 = q1; R.minimum = minimum;
    const int v[4] = {p0, p1, q0, q1};
    int32_t c[4] = {-1, -1, -1, -1};
    int used = 0;
    for (int s = 0; s < 4; ++s) {
      const int col = v[s] >= 0 ? nd[v[s]] : -1;
      R.slot[s] = -1;
      if (col < 0) continue;
      int t = 0;
      while (t < used && c[t] != col) ++t;  // vertices fed by the same design 2-vector share one Jacobian slot
      if (t == used) c[used++] = col;
      R.slot[s] = t;
    }
    rows.push_back(R);
    cols.insert(cols.end(), c, c + 4);
  };
  for (int b = 0; b < d->n_bonds; ++b)
    for (int k = 0; k < 2; ++k)
      if (d->bonds[2 * b + k] < 0 || d->bonds[2 * b + k] >= nn) { cudaSetDevice(cur); return fail(DFX_ERR_INVALID, "bond %d: node id out of range", b); }
  for (int i = 0; i < d->n_boundary; ++i)
    if (d->boundary_nodes[i] < 0 || d->boundary_nodes[i] >= nn) { cudaSetDevice(cur); return fail(DFX_ERR_INVALID, "boundary node %d out of range", i); }
  if (d->angles) {
    // compute_edge_angles (geometry.py:234-253): e1 = next - here, e2 = previous - here of each bond node
    for (int b = 0; b < d->n_bonds; ++b) { const int n1 = d->bonds[2 * b], n2 = d->bonds[2 * b + 1]; add(n2, prev(n2), n1, next(n1), d->min_void_angle); }
    for (int b = 0; b < d->n_bonds; ++b) { const int n1 = d->bonds[2 * b], n2 = d->bonds[2 * b + 1]; add(n1, prev(n1), n2, next(n2), d->min_void_angle); }
    for (int b = 0; b < d->n_bonds; ++b) { const int n1 = d->bonds[2 * b]; add(n1, next(n1), n1, prev(n1), d->min_block_angle); }
    for (int b = 0; b < d->n_bonds; ++b) { const int n2 = d->bonds[2 * b + 1]; add(n2, next(n2), n2, prev(n2), d->min_block_angle); }
    for (int i = 0; i < d->n_boundary; ++i) { const int n = d->boundary_nodes[i]; add(n, next(n), n, prev(n), d->min_block_angle); }
  }
  const int m_angle = (int)rows.size();
  if (d->edges)  // compute_edge_lengths (geometry.py:205-218): |vertex[l-1] - vertex[l]|
    for (int n = 0; n < nn; ++n) add(n, prev(n), -1, -1, d->min_edge_length);
  if (rows.empty()) { cudaSetDevice(cur); return fail(DFX_ERR_INVALID, "no constraint rows requested"); }
  DfxConstraints* h = new (std::nothrow) DfxConstraints();
  if (!h) { cudaSetDevice(cur); return fail(DFX_ERR_INVALID, "out of host memory"); }
  h->device = g->device; h->dev = g->dev; h->m = (int)rows.size(); h->m_angle = m_angle; h->rows = nullptr; h->cols = std::move(cols);
  e = upload(rows, &h->rows);
  cudaSetDevice(cur);
  if (e != cudaSuccess) { delete h; return fail(DFX_ERR_CUDA, "constraint table upload failed: %s", cudaGetErrorString(e)); }
  *out = h;
  return DFX_OK;
}

void dfx_constraints_destroy(DfxConstraints* h) {
  if (!h) return;
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(h->device);
  if (h->rows) cudaFree(h->rows);
  cudaSetDevice(cur);
  delete h;
}

int dfx_constraints_rows(const DfxConstraints* h, int32_t* n_angle_rows) {
  if (!h) return -1;
  if (n_angle_rows) *n_angle_rows = h->m_angle;
  return h->m;
}

int dfx_constraints_columns(const DfxConstraints* h, int32_t* cols) {
  if (!h || !cols) return fail(DFX_ERR_INVALID, "NULL argument");
  std::copy(h->cols.begin(), h->cols.end(), cols);
  return DFX_OK;
}

int dfx_constraints_eval(const DfxConstraints* h, int batch, const double* design, double* values, double* jac, void* stream_) {
  if (!h || !design || !values) return fail(DFX_ERR_INVALID, "NULL argument");
  if (batch <= 0) return fail(DFX_ERR_INVALID, "batch must be positive");
  const int threads = 128;
  dim3 grid((h->m + threads - 1) / threads, batch);
  constraints_kernel<<<grid, threads, 0, (cudaStream_t)stream_>>>(h->dev, h->rows, h->m, design, values, jac);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(DFX_ERR_CUDA, "constraints launch failed: %s", cudaGetErrorString(e));
  return DFX_OK;
}

}  // extern "C"

// ---- field reconstruction (dynamics.py:129-136, 169-182 without the dense Jacobian) ------------
namespace {
__global__ void expand_fields_kernel(DevTopo T, DfxLeaf drive, const double* ys, const double* ts, long long ts_bstride, int n_t,
                                     double* fields) {
  const int design = blockIdx.y, i = blockIdx.x;
  const int nf = T.n_free, nd = T.n_dof;
  const double* y = ys + ((long long)design * n_t + i) * 2 * nf;
  double* o = fields + ((long long)design * n_t + i) * 2 * nd;
  const double t = ts[(long long)design * ts_bstride + i];
  const double* dp = drive.ptr ? drive.ptr + (long long)design * drive.bstride : nullptr;
  DriveEval de;
  drive_eval(T.drive_kind, t, dp, true, de, T.table);
  for (int dof = threadIdx.x; dof < nd; dof += blockDim.x) {
    const int f = T.free_of_dof[dof];
    double u, v;
    if (f >= 0) { u = y[f]; v = y[nf + f]; }
    else {
      const int c = T.cons_slot[dof];
      u = T.drive_vec0[c] * de.s[0] + T.drive_vec1[c] * de.s[1];
      v = T.drive_vec0[c] * de.sdot[0] + T.drive_vec1[c] * de.sdot[1];
    }
    o[dof] = u;
    o[nd + dof] = v;
  }
}
}  // namespace

extern "C" int dfx_expand_fields(const DfxTopology* t, const DfxParams* params, int batch, const double* ys, const double* ts,
                                 int64_t ts_bstride, int n_t, double* fields, void* stream_) {
  if (!t || !params || !ys || !ts || !fields) return fail(DFX_ERR_INVALID, "NULL argument");
  if (t->dev.n_drive_params > 0 && !params->drive.ptr) return fail(DFX_ERR_INVALID, "drive signal needs params.drive");
  dim3 grid(n_t, batch);
  expand_fields_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(t->dev, params->drive, ys, ts, ts_bstride, n_t, fields);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(DFX_ERR_CUDA, "expand_fields launch failed: %s", cudaGetErrorString(e));
  return DFX_OK;
}

// ---- FP64 FMA peak micro-benchmark: the roofline denominator of this path (MEASURED_PEAKS.json holds
// HBM and bf16 only).  Register-resident chains of dependent DFMAs, 8 independent chains per thread.
namespace {
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
}  // namespace

// returns the measured FP64 throughput in TFLOP/s (2 flops per FMA), or a negative value on error
extern "C" double dfx_fp64_peak(void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1.0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int threads = 256, blocks = sms * 8, iters = 4096;
  double* out = nullptr;
  if (cudaMalloc((void**)&out, sizeof(double) * threads * blocks) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = -1.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, stream);
    fp64_peak_kernel<<<blocks, threads, 0, stream>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, stream);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 64.0 * iters * (double)threads * blocks;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}

// ---- self-test of the device math primitives (tests/test_gpu_parity.py::test_device_math) ------------
namespace {
__global__ void math_selftest_kernel(const double* x, const double* y, double* o_rsqrt, double* o_angle, double* o_rcp, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  o_rsqrt[i] = rsqrt_pos(x[i]);
  o_rcp[i] = rcp_pos(x[i]);
  const double r = rsqrt_pos(x[i] * x[i] + y[i] * y[i]);
  o_angle[i] = angle_of_unit(y[i] * r, x[i] * r);
}
__global__ void sincos_selftest_kernel(const double* x, double* o_sin, double* o_cos, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sincos_fast(x[i], &o_sin[i], &o_cos[i]);
}
}  // namespace

/* which adjoint kernel dfx_adjoint / dfx_adjoint_objective would launch for this topology, these leaf forms and this batch
 * (diagnostic: tests and bench.py report it).  The string lives in thread-local storage until the next call. */
extern "C" const char* dfx_adjoint_plan(const DfxTopology* t, const DfxParams* params, int batch) {
  static thread_local char buf[96];
  if (!t || !params) return "invalid";
  const FastPlan fp = plan_fast_adjoint(t->dev);
  const Fast3Plan f3 = fp.ok ? plan_adjoint3(t->dev, *params, t->n_cons_units) : Fast3Plan{};
  if (f3.ok) std::snprintf(buf, sizeof(buf), "adjoint3_kernel<%d,%d,%d> 768 threads", f3.npb, (int)f3.contact, f3.damp);
  else if (fp.ok) std::snprintf(buf, sizeof(buf), "adjoint2_kernel<%d,%d,%d>", fp.nt, fp.nt == 84 && fp.ns == 32 ? 32 : -1, fp.threads);
  else {
    const MultiCta mc = pick_multi_cta(t->dev, batch, t->sm_count);
    std::snprintf(buf, sizeof(buf), "adjoint_kernel<%d> %d CTA(s) per design", mc.mode == 2 ? 2 : (mc.ncta > 1 ? 1 : 0), mc.ncta);
  }
  return buf;
}

extern "C" int dfx_sincos_selftest(const double* x, double* o_sin, double* o_cos, int n, void* stream_) {
  sincos_selftest_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(x, o_sin, o_cos, n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(DFX_ERR_CUDA, "sincos_selftest launch failed: %s", cudaGetErrorString(e));
  return DFX_OK;
}

extern "C" int dfx_math_selftest(const double* x, const double* y, double* o_rsqrt, double* o_angle, double* o_rcp, int n, void* stream_) {
  math_selftest_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(x, y, o_rsqrt, o_angle, o_rcp, n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(DFX_ERR_CUDA, "math_selftest launch failed: %s", cudaGetErrorString(e));
  return DFX_OK;
}
